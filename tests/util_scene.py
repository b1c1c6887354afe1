"""Shared helpers for the parity tests: scene construction, oracle runs, decoding the CUDA
library's saved buffers."""

import numpy as np
import torch

from oracle import splat_ref as O
from harness import cameras, synthetic


def make_inputs(P, H, W, seed=0, sh_max=3, sh_degree=None, radius=0.5, phi=0.0, fovx=0.55,
                opacity="sigmoid_normal", cam_radius=3.5, exact_knn=None, scale_mul=1.0):
    sc = synthetic.ball_scene(P, radius=radius, sh_degree_max=sh_max, seed=seed, opacity=opacity,
                              exact_knn=True if exact_knn is None else exact_knn)
    sc["scales"] = sc["scales"] * scale_mul
    cam = cameras.orbit_camera(radius=cam_radius, phi_deg=phi, fovx=fovx, height=H, width=W)
    deg = sh_max if sh_degree is None else sh_degree
    return sc, cam, deg


def oracle_settings(cam, deg, bg=(1.0, 1.0, 1.0), score=False, scale_modifier=1.0):
    return O.Settings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                      torch.tensor(bg, dtype=torch.float32), scale_modifier, cam.world_view_transform,
                      cam.full_proj_transform, deg, cam.camera_center, False, score)


def cuda_settings(cam, deg, bg=(1.0, 1.0, 1.0), score=False, scale_modifier=1.0, device="cuda"):
    from dreamscene_b200 import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx,
        tanfovy=cam.tanfovy, bg=torch.tensor(bg, dtype=torch.float32, device=device),
        scale_modifier=scale_modifier, viewmatrix=cam.world_view_transform.to(device),
        projmatrix=cam.full_proj_transform.to(device), sh_degree=deg,
        campos=cam.camera_center.to(device), prefiltered=False, score_flag=score)


def decode_saved(saved: torch.Tensor, P, H, W, cap):
    """-> dict(num_pairs, tile_start[int64], idx[int64 per pair], depth_bits[uint32 per pair],
    n_contrib[H,W])."""
    from dreamscene_b200 import _lib
    vl = _lib.saved_layout(P, H, W, cap)
    raw = saved.cpu().numpy()
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    header = raw[vl.header:vl.header + 32].view(np.uint32)
    D = int(header[0])
    ts = raw[vl.tile_start:vl.tile_start + 4 * (ntiles + 1)].view(np.uint32).astype(np.int64)
    wo = raw[vl.work_order:vl.work_order + 4 * ntiles].view(np.uint32).astype(np.int64)
    nc = raw[vl.n_contrib:vl.n_contrib + 4 * H * W].view(np.uint32).reshape(H, W)
    n = min(D, cap)
    keys = raw[vl.keys:vl.keys + 8 * n].view(np.uint64)
    geom = raw[vl.geom:vl.geom + 48 * P].view(np.uint32).reshape(P, 12)
    fill = raw[vl.header + 4 * 32:vl.header + 4 * 64].view(np.uint32).astype(np.int64)     # backward work lists
    items = raw[vl.bwd_items:vl.bwd_items + 4 * 32 * ntiles * 8].view(np.uint32).reshape(32, ntiles * 8) \
        if vl.bwd_items < vl.total and len(raw) >= vl.total else None
    return dict(num_pairs=D, header=header.copy(), tile_start=ts, work_order=wo, n_contrib=nc,
                bwd_fill=fill, bwd_items=items,
                idx=(keys & np.uint64(0xFFFFFFFF)).astype(np.int64),
                depth_bits=(keys >> np.uint64(32)).astype(np.uint32), geom_f32=geom.view(np.float32))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
