"""rasterize_views (one binning/sort/composite pass for B views) against the per-view loop:
identical images and radii per view, bit-identical sorted lists, gradients equal to the sum over views."""
import numpy as np
import pytest
import torch

from tests import util_scene as U

pytestmark = pytest.mark.gpu


def _views(B, H, W, sh_degrees, bgs):
    cams = [U.cameras.orbit_camera(phi_deg=37.0 * k, theta_deg=60.0 + 7 * k, height=H, width=W) for k in range(B)]
    return cams, [U.cuda_settings(c, d, bg=b) for c, d, b in zip(cams, sh_degrees, bgs)]


@pytest.mark.parametrize("H,W", [(128, 128), (90, 150)])
def test_views_equal_the_per_view_loop_forward_and_backward(H, W):
    from dreamscene_b200 import GaussianRasterizer
    from dreamscene_b200.multiview import rasterize_views
    B, P = 3, 6000
    sc, _, _ = U.make_inputs(P, H, W, seed=4)
    cams, S = _views(B, H, W, sh_degrees=[3, 0, 2], bgs=[(1, 1, 1), (0, 0, 0), (0.2, 0.5, 0.9)])
    names = ("means3D", "opacities", "shs", "scales", "rotations")
    g = torch.Generator(device="cuda").manual_seed(3)
    gcs = [torch.randn(3, H, W, device="cuda", generator=g) / (H * W) for _ in range(B)]
    gds = [torch.randn(2, H, W, device="cuda", generator=g) / (H * W) for _ in range(B)]
    # per-view loop (the reference's call pattern)
    t1 = {k: sc[k].cuda().requires_grad_(True) for k in names}
    m1 = [torch.zeros(P, 3, device="cuda", requires_grad=True) for _ in range(B)]
    ref = [GaussianRasterizer(S[v])(means3D=t1["means3D"], means2D=m1[v], opacities=t1["opacities"], shs=t1["shs"],
                                    scales=t1["scales"], rotations=t1["rotations"]) for v in range(B)]
    torch.autograd.backward([r[0] for r in ref] + [r[2] for r in ref], gcs + gds)
    # fused
    t2 = {k: sc[k].cuda().requires_grad_(True) for k in names}
    m2 = [torch.zeros(P, 3, device="cuda", requires_grad=True) for _ in range(B)]
    got = rasterize_views(S, t2["means3D"], t2["opacities"], shs=t2["shs"], scales=t2["scales"], rotations=t2["rotations"],
                          means2D=m2)
    torch.autograd.backward([r[0] for r in got] + [r[2] for r in got], gcs + gds)
    for v in range(B):
        assert got[v][0].shape == (3, H, W) and got[v][2].shape == (2, H, W)
        assert torch.equal(got[v][1], ref[v][1])                                 # radii
        assert torch.equal(got[v][0], ref[v][0]) and torch.equal(got[v][2], ref[v][2])   # same lists, same arithmetic
        assert U.rel_err(m2[v].grad, m1[v].grad) < 1e-5
    for k in names:
        assert U.rel_err(t2[k].grad, t1[k].grad) < 1e-5, k                       # accumulated over views in the kernel


def test_views_with_per_view_parameters_and_score():
    """scene_render augments shs / scales separately per view: per-view lists of those tensors, shared rest."""
    from dreamscene_b200 import GaussianRasterizer
    from dreamscene_b200.multiview import rasterize_views
    B, P, H, W = 2, 3000, 96, 96
    sc, _, _ = U.make_inputs(P, H, W, seed=6, sh_max=1)
    cams = [U.cameras.orbit_camera(phi_deg=50.0 * k, height=H, width=W) for k in range(B)]
    S = [U.cuda_settings(c, 1, score=True) for c in cams]
    base = {k: v.cuda() for k, v in sc.items()}
    shs_v = [(base["shs"] * (1 + 0.1 * k)).requires_grad_(True) for k in range(B)]
    sca_v = [(base["scales"] * (1 + 0.05 * k)).requires_grad_(True) for k in range(B)]
    shared = {k: base[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "rotations")}
    got = rasterize_views(S, shared["means3D"], shared["opacities"], shs=shs_v, scales=sca_v, rotations=shared["rotations"])
    (sum(r[1].sum() for r in got) + sum(r[3].sum() for r in got)).backward()
    gs = {k: v.grad.clone() for k, v in shared.items()}
    gl = [t.grad.clone() for t in shs_v + sca_v]
    for t in list(shared.values()) + shs_v + sca_v:
        t.grad = None
    ref = [GaussianRasterizer(S[v])(means3D=shared["means3D"], means2D=torch.zeros(P, 3, device="cuda"),
                                    opacities=shared["opacities"], shs=shs_v[v], scales=sca_v[v],
                                    rotations=shared["rotations"]) for v in range(B)]
    (sum(r[1].sum() for r in ref) + sum(r[3].sum() for r in ref)).backward()
    for v in range(B):
        assert U.rel_err(got[v][0], ref[v][0]) < 1e-5                              # important_score per view
        assert torch.equal(got[v][1], ref[v][1]) and torch.equal(got[v][2], ref[v][2])
    for k in shared:
        assert U.rel_err(gs[k], shared[k].grad) < 1e-5, k
    for a, t in zip(gl, shs_v + sca_v):
        assert U.rel_err(a, t.grad) < 1e-5
