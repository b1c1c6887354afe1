"""View sharding, factored SH exchange (dreamscene_b200.parallel.exchange_factored): one GPU plays every rank.

dL/dsh of a view = basis(view direction) x dL/d(clamped colour).  The in-backward reduction therefore sends [P, 3]
colour gradients + the camera centre per rank (one all-gather) and rebuilds the summed [P, M, 3] rows with
b200gsr_sh_grad_expand.  Here the per-rank payloads are captured from real backwards of different views, stacked as the
all-gather would deliver them, expanded, and compared with the sum of the dense per-view SH gradients."""
import ctypes as C

import pytest
import torch

from tests import util_scene as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sh_max,deg", [(3, 3), (3, 1), (1, 1), (3, 0)])
def test_factored_payloads_rebuild_the_sum_of_the_dense_sh_gradients(sh_max, deg, monkeypatch):
    from dreamscene_b200 import GaussianRasterizer, _lib, parallel
    H = W = 96
    P, B = 5000, 3
    sc, _, _ = U.make_inputs(P, H, W, seed=11, sh_max=sh_max)
    M = sc["shs"].shape[1]
    cams = [U.cameras.orbit_camera(phi_deg=115.0 * k, theta_deg=55.0 + 9 * k, height=H, width=W) for k in range(B)]
    S = [U.cuda_settings(c, deg) for c in cams]
    names = ("means3D", "opacities", "shs", "scales", "rotations")
    g = torch.Generator(device="cuda").manual_seed(2)
    gcs = [torch.randn(3, H, W, device="cuda", generator=g) / (H * W) for _ in range(B)]
    gds = [torch.randn(2, H, W, device="cuda", generator=g) / (H * W) for _ in range(B)]

    def run(view, t):
        out = GaussianRasterizer(S[view])(means3D=t["means3D"], means2D=torch.zeros(P, 3, device="cuda", requires_grad=True),
                                          opacities=t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
        torch.autograd.backward([out[0], out[2]], [gcs[view], gds[view]])

    # dense reference: per-view gradients, summed
    dense = {k: 0 for k in names}
    for v in range(B):
        t = {k: sc[k].cuda().requires_grad_(True) for k in names}
        run(v, t)
        for k in names:
            dense[k] = dense[k] + t[k].grad
    # "sharded": every view's backward takes the factored path; the exchange is captured instead of sent
    payloads, flats = [], []

    def fake_exchange(flat, dcol, P_, M_, deg_, means3D):
        payloads.append(dcol.clone()); flats.append(flat.clone())
        return torch.zeros(P_, M_, 3, device=dcol.device)

    monkeypatch.setattr(parallel, "reduction_active", lambda: True)
    monkeypatch.setattr(parallel, "factored_sh_exchange", lambda: True)
    monkeypatch.setattr(parallel, "exchange_factored", fake_exchange)
    other = {k: 0 for k in names if k != "shs"}
    for v in range(B):
        t = {k: sc[k].cuda().requires_grad_(True) for k in names}
        run(v, t)
        for k in other:
            other[k] = other[k] + t[k].grad
    monkeypatch.undo()
    assert len(payloads) == B and payloads[0].numel() == parallel.factored_stride(P)
    for v in range(B):      # the payload carries the view's camera centre behind the [P, 3] colour gradients
        assert torch.equal(payloads[v][3 * P:3 * P + 3].cpu(), cams[v].camera_center.float().reshape(3))
    # the other parameter gradients do not depend on how the SH gradient leaves the kernel (two runs of the
    # backward differ only by the order of its fp32 atomics)
    for k in other:
        assert U.rel_err(other[k], dense[k]) < 1e-5, k
    gathered = torch.stack(payloads).contiguous()
    d_sh = torch.full((P, M, 3), float("nan"), device="cuda")
    rc = _lib.load().b200gsr_sh_grad_expand(P, M, deg, B, C.c_void_p(t["means3D"].data_ptr()), C.c_void_p(gathered.data_ptr()),
                                            gathered.shape[1], C.c_void_p(d_sh.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    assert torch.isfinite(d_sh).all()
    assert float(d_sh[:, (deg + 1) ** 2:].abs().max()) == 0.0 if (deg + 1) ** 2 < M else True
    # two backward runs differ by the order of their fp32 atomics (~1e-6 relative): the bounds leave room for that
    scale = float(dense["shs"].abs().max())
    assert scale > 0 and float((d_sh - dense["shs"]).abs().max()) <= 2e-5 * scale
    assert U.rel_err(d_sh, dense["shs"]) < 1e-5


def test_expand_rejects_bad_arguments():
    from dreamscene_b200 import _lib
    lib = _lib.load()
    z = torch.zeros(64, device="cuda")
    p = C.c_void_p(z.data_ptr())
    assert lib.b200gsr_sh_grad_expand(4, 16, 3, 0, p, p, 64, p, None) != 0          # no views
    assert lib.b200gsr_sh_grad_expand(4, 16, 3, 65, p, p, 64, p, None) != 0         # too many views
    assert lib.b200gsr_sh_grad_expand(4, 4, 3, 1, p, p, 64, p, None) != 0           # M smaller than the degree needs
    assert lib.b200gsr_sh_grad_expand(4, 16, 3, 1, p, p, 12, p, None) != 0          # stride without room for the camera
    assert lib.b200gsr_sh_grad_expand(0, 16, 3, 1, None, None, 64, None, None) == 0   # empty scene
