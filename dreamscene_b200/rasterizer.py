"""Host side of the drop-in `diff_gaussian_rasterization` replacement.

Mirrors the Python surface of DreamScene's un-vendored extension exactly as DreamScene uses it
(/root/reference/scene_gaussian.py:586-601,637-646 [score], :737-752,861-870 [scene],
:951-966,1012-1021 [object]):

    GaussianRasterizationSettings(image_height, image_width, tanfovx, tanfovy, bg,
        scale_modifier, viewmatrix, projmatrix, sh_degree, campos, prefiltered, score_flag)
    GaussianRasterizer(raster_settings)(means3D, means2D, opacities, shs, colors_precomp,
        scales, rotations, cov3D_precomp)
      -> (color[3,H,W], radii[P] int32, depth_alpha[2,H,W])                    score_flag False
      -> (important_score[P], color, radii, depth_alpha)                        score_flag True

All compute happens in libb200gsr.so (hand-written sm_100a CUDA) through the C ABI of
include/b200gsr.h; PyTorch provides device memory, the stream and autograd plumbing only.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib
from . import parallel as _parallel


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    score_flag: bool = False


# ------------------------------------------------------------------------------------------
# Workspace: one transient scratch buffer + pair-capacity estimate per (device, stream).
# ------------------------------------------------------------------------------------------
class _Workspace:
    def __init__(self, device: torch.device):
        self.device = device
        self.scratch: Optional[torch.Tensor] = None
        self.capacity = 0                       # pair capacity used for the next call
        self.notify = torch.zeros(4, dtype=torch.int32).pin_memory()   # device-mapped host words
        self.seq = 0
        self.last_pairs = 0

    def ensure_scratch(self, nbytes: int) -> torch.Tensor:
        if self.scratch is None or self.scratch.numel() < nbytes:
            self.scratch = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=self.device)
        return self.scratch


_workspaces: dict = {}
_MIN_CAPACITY = 1 << 20
_POLL_TIMEOUT_S = 60.0


def _workspace(device: torch.device) -> _Workspace:
    stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = _workspaces[key] = _Workspace(device)
    return ws


def set_workspace_capacity(max_pairs: int, device=None) -> None:
    """Optional: pre-size the (tile, Gaussian) pair capacity of the current stream's workspace."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _workspace(dev).capacity = int(max_pairs)


def _round_cap(n: int) -> int:
    g = 1 << 18
    return max(_MIN_CAPACITY, (int(n) + g - 1) // g * g)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:          # the kernels use 16-byte vector loads on rows
        t = t.clone()
    return t


def _make_params(rs: GaussianRasterizationSettings, P: int, M: int, keep: list, dev) -> _lib.Params:
    # the per-view constants live on the Gaussians' device (upstream requires CUDA tensors here too)
    bg = _f32c(rs.bg.detach().to(dev)); vm = _f32c(rs.viewmatrix.detach().to(dev))
    pm = _f32c(rs.projmatrix.detach().to(dev)); cp = _f32c(rs.campos.detach().to(dev))
    keep.extend([bg, vm, pm, cp])
    return _lib.Params(P, M, int(rs.sh_degree), int(rs.image_height), int(rs.image_width),
                       float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier),
                       int(bool(rs.prefiltered)), int(bool(rs.score_flag)),
                       bg.data_ptr(), vm.data_ptr(), pm.data_ptr(), cp.data_ptr())


class _State:
    """Everything backward needs that is not a tensor input."""
    __slots__ = ("params_keep", "P", "M", "capacity", "saved", "rs")


def _forward_impl(rs, means3D, shs, colors, opac, scales, rots, cov3d):
    lib = _lib.load()
    dev = means3D.device
    if dev.type != "cuda":
        raise RuntimeError("diff_gaussian_rasterization (b200gsr): inputs must be CUDA tensors; "
                           "there is no CPU fallback")
    P = int(means3D.shape[0])
    M = int(shs.shape[1]) if shs is not None else 0
    H, W = int(rs.image_height), int(rs.image_width)
    keep: list = []
    prm = _make_params(rs, P, M, keep, dev)
    color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
    depth_alpha = torch.empty(2, H, W, dtype=torch.float32, device=dev)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    score = torch.zeros(P, dtype=torch.float32, device=dev) if rs.score_flag else None
    ws = _workspace(dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    cap = ws.capacity if ws.capacity > 0 else _round_cap(6 * P)
    if torch.cuda.current_device() != dev.index:
        torch.cuda.set_device(dev)        # the library launches on the CURRENT device (callers use cuda:0)
    while True:
        sl = _lib.scratch_layout(P, H, W, cap)
        vl = _lib.saved_layout(P, H, W, cap)
        scratch = ws.ensure_scratch(sl.total)
        saved = torch.empty(vl.total, dtype=torch.uint8, device=dev)
        ws.seq = (ws.seq + 1) & 0x7FFFFFFF or 1
        rc = lib.b200gsr_forward(C.byref(prm), _ptr(means3D), _ptr(shs), _ptr(colors), _ptr(opac),
                                 _ptr(scales), _ptr(rots), _ptr(cov3d), _ptr(color), _ptr(depth_alpha),
                                 _ptr(radii), _ptr(score), _ptr(scratch), scratch.numel(), _ptr(saved),
                                 saved.numel(), cap, C.c_void_p(ws.notify.data_ptr()), ws.seq, stream)
        if rc:
            msg = _lib.last_error()
            if rc == -1:
                raise Exception(msg)
            raise RuntimeError(f"b200gsr_forward failed ({rc}): {msg}")
        # Wait only for the tile scan (project + scan kernels); sort/composite keep running.
        t0 = time.perf_counter()
        n = ws.notify
        while int(n[0]) != ws.seq:
            if time.perf_counter() - t0 > _POLL_TIMEOUT_S:
                torch.cuda.synchronize(dev)
                if int(n[0]) == ws.seq:
                    break
                raise RuntimeError("b200gsr_forward: device never reported the pair count")
        pairs = int(n[1]) & 0xFFFFFFFF
        ws.last_pairs = pairs
        if pairs <= cap:
            # high-water mark with 25% head-room: the capacity only costs 8 B per pair in `saved`, and
            # views of one training step differ a lot in pair count (random cameras), so never shrink
            ws.capacity = max(ws.capacity, _round_cap(int(pairs * 1.25)), _MIN_CAPACITY)
            break
        cap = ws.capacity = _round_cap(int(pairs * 1.5))   # overflow: re-issue with enough room
        if score is not None:
            score.zero_()
    st = _State()
    st.params_keep = keep; st.P = P; st.M = M; st.capacity = cap; st.saved = saved; st.rs = rs
    return color, radii, depth_alpha, score, st


def grad_sections(P: int, n_col: int, has_sr: bool):
    """Offsets (in floats) of the parameter-gradient sections inside the flat buffer that backward
    fills and (under view sharding) all-reduces: means3D[P,3], opac[P,1], col[P,n_col], then
    scales[P,3]+rots[P,4] or cov[P,6].  Every section starts on a 256-byte boundary because the
    kernels use 16-byte vector stores.  Returns (offsets dict, total floats)."""
    widths = [("means3D", 3), ("opac", 1), ("col", n_col)] + ([("scales", 3), ("rots", 4)] if has_sr else [("cov", 6)])
    offs, o = {}, 0
    for name, wdt in widths:
        offs[name] = o
        o += (P * wdt + 63) // 64 * 64
    return offs, o


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        means3D = _f32c(means3D); sh = _f32c(sh); colors_precomp = _f32c(colors_precomp)
        opacities = _f32c(opacities); scales = _f32c(scales); rotations = _f32c(rotations)
        cov3Ds_precomp = _f32c(cov3Ds_precomp)
        color, radii, depth_alpha, score, st = _forward_impl(
            raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp)
        ctx.st = st
        ctx.has = (sh is not None, colors_precomp is not None, scales is not None, cov3Ds_precomp is not None)
        tensors = [means3D, opacities, radii, depth_alpha]
        for t in (sh, colors_precomp, scales, rotations, cov3Ds_precomp):
            tensors.append(t if t is not None else torch.empty(0, device=means3D.device))
        ctx.save_for_backward(*tensors)
        ctx.mark_non_differentiable(radii)
        if raster_settings.score_flag:
            ctx.mark_non_differentiable(score)
            return score, color, radii, depth_alpha
        return color, radii, depth_alpha

    @staticmethod
    def backward(ctx, *grads):
        st = ctx.st
        rs = st.rs
        if rs.score_flag:
            _, g_color, _, g_da = grads
        else:
            g_color, _, g_da = grads
        means3D, opacities, radii, depth_alpha, sh, colors, scales, rots, cov3d = ctx.saved_tensors
        has_sh, has_col, has_sr, has_cov = ctx.has
        sh = sh if has_sh else None; colors = colors if has_col else None
        scales = scales if has_sr else None; rots = rots if has_sr else None
        cov3d = cov3d if has_cov else None
        dev = means3D.device
        P, M = st.P, st.M
        H, W = int(rs.image_height), int(rs.image_width)
        g_color = torch.zeros(3, H, W, device=dev) if g_color is None else _f32c(g_color)
        g_da = torch.zeros(2, H, W, device=dev) if g_da is None else _f32c(g_da)

        # one flat buffer for every parameter gradient: a single NCCL all-reduce when views are
        # sharded across ranks (dreamscene_b200.parallel), and one allocation otherwise
        n_col = 3 * M if has_sh else 3
        offs, o = grad_sections(P, n_col, has_sr)
        flat = torch.empty(max(o, 1), dtype=torch.float32, device=dev)
        sec = lambda name, wdt: flat[offs[name]:offs[name] + P * wdt]
        d_means3D = sec("means3D", 3).view(P, 3)
        d_opac = sec("opac", 1).view(P, 1)
        d_colsh = sec("col", n_col)
        d_sh = d_colsh.view(P, M, 3) if has_sh else None
        d_colors = d_colsh.view(P, 3) if has_col else None
        if has_sr:
            d_scales = sec("scales", 3).view(P, 3)
            d_rots = sec("rots", 4).view(P, 4)
            d_cov = None
        else:
            d_scales = d_rots = None
            d_cov = sec("cov", 6).view(P, 6)
        d_means2D = torch.empty(P, 3, dtype=torch.float32, device=dev)

        if P > 0:
            lib = _lib.load()
            if torch.cuda.current_device() != dev.index:
                torch.cuda.set_device(dev)
            keep: list = []
            prm = _make_params(rs, P, M, keep, dev)
            ws = _workspace(dev)
            sl = _lib.scratch_layout(P, H, W, st.capacity)
            scratch = ws.ensure_scratch(sl.total)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = lib.b200gsr_backward(
                C.byref(prm), _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacities), _ptr(scales),
                _ptr(rots), _ptr(cov3d), _ptr(radii), _ptr(depth_alpha), _ptr(g_color), _ptr(g_da),
                _ptr(st.saved), st.saved.numel(), _ptr(scratch), scratch.numel(), st.capacity,
                _ptr(d_means3D), _ptr(d_means2D), _ptr(d_sh), _ptr(d_colors), _ptr(d_opac),
                _ptr(d_scales), _ptr(d_rots), _ptr(d_cov), stream)
            if rc:
                raise RuntimeError(f"b200gsr_backward failed ({rc}): {_lib.last_error()}")
            _parallel.maybe_all_reduce(flat)
        return (d_means3D, d_means2D, d_sh, d_colors, d_opac, d_scales, d_rots, d_cov, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """Frustum (near-plane) visibility mask; unused by DreamScene, kept for API parity."""
        rs = self.raster_settings
        with torch.no_grad():
            pos = _f32c(positions)
            vis = torch.empty(pos.shape[0], dtype=torch.uint8, device=pos.device)
            vm, pm = _f32c(rs.viewmatrix), _f32c(rs.projmatrix)
            rc = _lib.load().b200gsr_mark_visible(
                int(pos.shape[0]), _ptr(pos), _ptr(vm), _ptr(pm), _ptr(vis),
                C.c_void_p(torch.cuda.current_stream(pos.device).cuda_stream))
            if rc:
                raise RuntimeError(f"b200gsr_mark_visible failed ({rc}): {_lib.last_error()}")
        return vis.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, self.raster_settings)
