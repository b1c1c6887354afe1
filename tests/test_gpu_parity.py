"""GPU parity tests: the CUDA path (through the public Python surface -> C ABI) against the oracle.

Tolerances (BASELINE.json north_star): forward colour/depth/alpha 1e-4 abs fp32, tile/sort
indices bit-exact, backward 1e-3 rel.  Because alpha>=1/255 and T<1e-4 are discontinuous tests,
two fp32 implementations whose exp() differ by 1 ulp may disagree on a vanishing fraction of
(pixel, Gaussian) pairs; those pixels are bounded separately (<= 1e-4 of all pixels, each within
one blend weight 1/255)."""
import numpy as np
import pytest
import torch

from oracle import splat_ref as O
from tests import util_scene as U

pytestmark = pytest.mark.gpu

FWD_ATOL = 1e-4
OUTLIER_FRAC = 1e-4
OUTLIER_MAX = 6e-3


def run_cuda(sc, cam, deg, score=False, grads=None, bg=(1.0, 1.0, 1.0), use="sh+sr", scale_modifier=1.0):
    from dreamscene_b200 import GaussianRasterizer
    from dreamscene_b200 import rasterizer as R
    dev = "cuda"
    S = U.cuda_settings(cam, deg, bg, score, scale_modifier)
    t = {k: v.detach().clone().to(dev).requires_grad_(grads is not None) for k, v in sc.items()}
    m2d = torch.zeros(t["means3D"].shape[0], 3, device=dev, requires_grad=grads is not None)
    kw = dict(means3D=t["means3D"], means2D=m2d, opacities=t["opacities"])
    if "colors" in use:
        kw["colors_precomp"] = t["colors_precomp"]
    else:
        kw["shs"] = t["shs"]
    if "cov" in use:
        kw["cov3D_precomp"] = t["cov3D_precomp"]
    else:
        kw["scales"] = t["scales"]; kw["rotations"] = t["rotations"]
    out = GaussianRasterizer(S)(**kw)
    res = dict(zip(("score", "color", "radii", "depth_alpha") if score else ("color", "radii", "depth_alpha"), out))
    if grads is not None:
        gc, gd = grads
        loss = (res["color"] * gc.to(dev)).sum() + (res["depth_alpha"] * gd.to(dev)).sum()
        loss.backward()
        res["grads"] = {k: v.grad.detach().cpu() for k, v in t.items() if v.grad is not None}
        res["grads"]["means2D"] = m2d.grad.detach().cpu()
    torch.cuda.synchronize()
    return res


def run_oracle(sc, cam, deg, score=False, grads=None, bg=(1.0, 1.0, 1.0), use="sh+sr", dtype=torch.float32,
               decisions=None, scale_modifier=1.0, record=False):
    """record=True: an fp32 run also records its per-(entry, pixel) blend decisions; they travel in
    r["decisions"] and an fp64 run given those decisions replays them (oracle.splat_ref.composite)."""
    S = U.oracle_settings(cam, deg, bg, score, scale_modifier)
    t = {k: v.detach().clone().requires_grad_(grads is not None) for k, v in sc.items()}
    m2d = torch.zeros(t["means3D"].shape[0], 3, requires_grad=grads is not None)
    kw = dict(means3D=t["means3D"], opacities=t["opacities"], means2D=m2d)
    if "colors" in use:
        kw["colors_precomp"] = t["colors_precomp"]
    else:
        kw["shs"] = t["shs"]
    if "cov" in use:
        kw["cov3D_precomp"] = t["cov3D_precomp"]
    else:
        kw["scales"] = t["scales"]; kw["rotations"] = t["rotations"]
    r = O.rasterize(S, dtype=dtype, decisions=decisions, record_blend={} if record else None, **kw)
    if grads is not None:
        gc, gd = grads
        loss = (r["color"] * gc.to(dtype)).sum() + (r["depth_alpha"] * gd.to(dtype)).sum()
        loss.backward()
        r["grads"] = {k: v.grad.detach() for k, v in t.items() if v.grad is not None}
        r["grads"]["means2D"] = m2d.grad.detach()
    return r


def check_images(cu, ref):
    for name in ("color", "depth_alpha"):
        a, b = cu[name].detach().cpu(), ref[name].detach().float()
        diff = (a - b).abs()
        bad = diff > FWD_ATOL
        frac = bad.float().mean().item()
        assert frac <= OUTLIER_FRAC, f"{name}: {frac:.2e} of values differ by more than {FWD_ATOL}"
        assert diff.max().item() <= OUTLIER_MAX * max(1.0, b.abs().max().item()), \
            f"{name}: max abs diff {diff.max().item()}"


def check_lists(sc, cam, ref):
    """bit-exact: radii, per-tile ranges, sorted Gaussian ids, depth bits."""
    from dreamscene_b200 import rasterizer as R
    dev = torch.device("cuda", torch.cuda.current_device())
    S = U.cuda_settings(cam, 3 if sc["shs"].shape[1] == 16 else 1)
    t = {k: v.to(dev) for k, v in sc.items()}
    with torch.no_grad():
        color, radii, da, _, st = R._forward_impl(S, t["means3D"], t["shs"], None, t["opacities"],
                                                  t["scales"], t["rotations"], None)
    torch.cuda.synchronize()
    P = sc["means3D"].shape[0]
    dec = U.decode_saved(st.saved, P, cam.image_height, cam.image_width, st.capacity)
    assert dec["num_pairs"] == len(ref["point_list"])
    np.testing.assert_array_equal(radii.cpu().numpy(), ref["radii"].numpy())
    np.testing.assert_array_equal(dec["tile_start"][:-1], ref["ranges"][:, 0])
    np.testing.assert_array_equal(dec["tile_start"][1:], ref["ranges"][:, 1])
    np.testing.assert_array_equal(dec["idx"], ref["point_list"])
    ref_bits = (ref["keys"] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    np.testing.assert_array_equal(dec["depth_bits"], ref_bits)
    return dec


def test_cfg1_forward_and_indices_10k_256():
    sc, cam, deg = U.make_inputs(10000, 256, 256)
    ref = run_oracle(sc, cam, deg, score=True)
    cu = run_cuda(sc, cam, deg, score=True)
    np.testing.assert_array_equal(cu["radii"].cpu().numpy(), ref["radii"].numpy())
    check_images(cu, ref)
    dec = check_lists(sc, cam, ref)
    nc_bad = (dec["n_contrib"] != ref["n_contrib"].numpy()).mean()
    assert nc_bad <= 1e-3, f"n_contrib mismatch fraction {nc_bad}"
    assert U.rel_err(cu["score"], ref["score"]) < 1e-4


def test_backward_work_lists_cover_every_blended_block_once():
    """The forward files every 8x4-pixel block that blended an entry under the size class of its consumed list
    length (common.cuh::gsr_bwd_class); the backward pops the classes longest first.  Checked against the
    n_contrib plane of the same forward: exact multiset of (class, tile*8 + block)."""
    from dreamscene_b200 import rasterizer as R
    sc, cam, deg = U.make_inputs(20000, 200, 136, seed=5)          # ragged image: partial tiles at both edges
    H, W = cam.image_height, cam.image_width
    dev = torch.device("cuda", torch.cuda.current_device())
    S = U.cuda_settings(cam, deg)
    t = {k: v.to(dev) for k, v in sc.items()}
    color, radii, da, _, st = R._forward_impl(S, t["means3D"], t["shs"], None, t["opacities"], t["scales"],
                                              t["rotations"], None, with_backward=True)
    torch.cuda.synchronize()
    dec = U.decode_saved(st.saved, sc["means3D"].shape[0], H, W, st.capacity)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    nc = np.zeros((gy * 16, gx * 16), np.int64)
    nc[:H, :W] = dec["n_contrib"]
    # block index inside the tile: (blk & 1) -> x half (8 px), (blk >> 1) -> y quarter (4 px)
    nb = nc.reshape(gy, 4, 4, gx, 2, 8).max(axis=(2, 5))           # [ty, yq, tx, xh]
    expect = set()
    for ty in range(gy):
        for yq in range(4):
            for tx in range(gx):
                for xh in range(2):
                    n = int(nb[ty, yq, tx, xh])
                    if n > 0:
                        e = n.bit_length() - 1
                        k = min(31, 2 * e + (((n >> (e - 1)) & 1) if e > 0 else 0))
                        expect.add((k, (ty * gx + tx) * 8 + yq * 2 + xh))
    got = set()
    for k in range(32):
        for it in dec["bwd_items"][k, :dec["bwd_fill"][k]]:
            assert (k, int(it)) not in got
            got.add((k, int(it)))
    assert got == expect and len(expect) > 50
    assert len({k for k, _ in expect}) >= 4                        # several size classes in use


def test_backward_matches_fp64_oracle_on_fp32_lists():
    sc, cam, deg = U.make_inputs(3000, 128, 128, seed=3)
    H = W = 128
    g = torch.Generator().manual_seed(1)
    grads = (torch.randn(3, H, W, generator=g) / (H * W), torch.randn(2, H, W, generator=g) / (H * W))
    ref32 = run_oracle(sc, cam, deg, record=True)
    ref = run_oracle(sc, cam, deg, grads=grads, dtype=torch.float64, decisions=ref32["decisions"])
    cu = run_cuda(sc, cam, deg, grads=grads)
    check_images(cu, ref32)
    for k in ("means3D", "scales", "rotations", "opacities", "shs", "means2D"):
        e = U.rel_err(cu["grads"][k], ref["grads"][k])
        assert e < 1e-3, f"grad {k}: rel err {e}"
    assert float(cu["grads"]["means2D"][:, 2].abs().max()) == 0.0


@pytest.mark.parametrize("deg,sh_max", [(0, 3), (1, 3), (2, 3), (1, 1), (0, 0), (2, 2)])
def test_sh_degrees_and_strides(deg, sh_max):
    sc, cam, _ = U.make_inputs(1500, 96, 96, seed=5, sh_max=sh_max)
    H = W = 96
    g = torch.Generator().manual_seed(2)
    grads = (torch.randn(3, H, W, generator=g), torch.randn(2, H, W, generator=g))
    ref32 = run_oracle(sc, cam, deg, record=True)
    ref = run_oracle(sc, cam, deg, grads=grads, dtype=torch.float64, decisions=ref32["decisions"])
    cu = run_cuda(sc, cam, deg, grads=grads)
    check_images(cu, ref32)
    for k in ("means3D", "shs", "opacities"):
        assert U.rel_err(cu["grads"][k], ref["grads"][k]) < 1e-3, k
    ncoef = (deg + 1) ** 2
    assert float(cu["grads"]["shs"][:, ncoef:].abs().max() if ncoef < sc["shs"].shape[1] else 0.0) == 0.0


def test_precomputed_colour_and_covariance_inputs():
    sc, cam, deg = U.make_inputs(2000, 112, 80, seed=7)
    H, W = 112, 80
    sc2 = dict(sc)
    sc2["colors_precomp"] = torch.rand(2000, 3, generator=torch.Generator().manual_seed(4))
    sc2["cov3D_precomp"] = O.cov3d_from_scale_rot(sc["scales"], sc["rotations"], 1.0, torch.float32)
    for k in ("shs", "scales", "rotations"):
        sc2.pop(k)
    g = torch.Generator().manual_seed(9)
    grads = (torch.randn(3, H, W, generator=g), torch.randn(2, H, W, generator=g))
    ref32 = run_oracle(sc2, cam, deg, use="colors+cov")
    ref = run_oracle(sc2, cam, deg, grads=grads, use="colors+cov", dtype=torch.float64,
                     decisions=ref32["decisions"])
    cu = run_cuda(sc2, cam, deg, grads=grads, use="colors+cov")
    np.testing.assert_array_equal(cu["radii"].cpu().numpy(), ref32["radii"].numpy())
    check_images(cu, ref32)
    for k in ("means3D", "colors_precomp", "cov3D_precomp", "opacities"):
        assert U.rel_err(cu["grads"][k], ref["grads"][k]) < 1e-3, k


@pytest.mark.parametrize("H,W,fovx", [(150, 200, 0.7), (67, 33, 0.55), (16, 16, 0.3), (270, 480, 0.96)])
def test_non_square_partial_tiles(H, W, fovx):
    sc, cam, deg = U.make_inputs(4000, H, W, seed=11, fovx=fovx)
    assert cam.tanfovx != cam.tanfovy or H == W
    ref = run_oracle(sc, cam, deg)
    cu = run_cuda(sc, cam, deg)
    np.testing.assert_array_equal(cu["radii"].cpu().numpy(), ref["radii"].numpy())
    check_images(cu, ref)
    check_lists(sc, cam, ref)


def test_empty_and_culled_inputs():
    from dreamscene_b200 import GaussianRasterizer
    sc, cam, deg = U.make_inputs(64, 64, 64, seed=1)
    S = U.cuda_settings(cam, deg, bg=(0.2, 0.5, 0.9))
    # P = 0
    e = lambda *s: torch.zeros(*s, device="cuda", requires_grad=True)
    color, radii, da = GaussianRasterizer(S)(means3D=e(0, 3), means2D=e(0, 3), opacities=e(0, 1),
                                             shs=e(0, 16, 3), scales=e(0, 3), rotations=e(0, 4))
    (color.sum() + da.sum()).backward()
    assert radii.numel() == 0
    assert torch.allclose(color[:, 0, 0].cpu(), torch.tensor([0.2, 0.5, 0.9]))
    assert float(da[0].abs().max()) == 0.0 and float(da[1].min()) == 1.0
    # everything behind the camera
    sc["means3D"] = sc["means3D"] + cam.camera_center * 2.0
    cu = run_cuda(sc, cam, deg, grads=(torch.ones(3, 64, 64), torch.ones(2, 64, 64)), bg=(0.2, 0.5, 0.9))
    assert int(cu["radii"].abs().sum()) == 0
    for k, v in cu["grads"].items():
        assert float(v.abs().max()) == 0.0, k


def test_zero_scales_and_no_grad_mode():
    sc, cam, deg = U.make_inputs(2000, 96, 96, seed=13)
    sc["scales"][::3] = 0.0            # scene_gaussian.py:855-857 clamps noised scales at exactly 0
    sc["scales"][1::7, 1] = 0.0
    H = W = 96
    grads = (torch.ones(3, H, W) / 100, torch.ones(2, H, W) / 100)
    ref32 = run_oracle(sc, cam, deg, record=True)
    ref = run_oracle(sc, cam, deg, grads=grads, dtype=torch.float64, decisions=ref32["decisions"])
    cu = run_cuda(sc, cam, deg, grads=grads)
    check_images(cu, ref32)
    for k, v in cu["grads"].items():
        assert torch.isfinite(v).all(), k
        assert U.rel_err(v, ref["grads"][k]) < 2e-3, k
    with torch.no_grad():
        cu2 = run_cuda(sc, cam, deg)
    assert torch.equal(cu2["color"], cu["color"])


def test_long_tile_lists_exercise_both_sort_paths_and_overflow_retry():
    """All Gaussians inside a few tiles: lists > 4096 (big in-smem sort) and > 16384 (out-of-core
    path); also starts from a tiny capacity so the overflow -> retry path runs."""
    from dreamscene_b200 import rasterizer as R
    P, H, W = 40000, 48, 48
    sc, cam, deg = U.make_inputs(P, H, W, seed=17, radius=0.3, cam_radius=6.0, exact_knn=False, scale_mul=0.5)
    ref = run_oracle(sc, cam, deg)
    assert (ref["ranges"][:, 1] - ref["ranges"][:, 0]).max() > 16384
    dstate = R._device_state(torch.device("cuda", torch.cuda.current_device()))
    old = (R._pair_mode, R._MIN_CAPACITY, R._MIN_PAIRS_PER_GAUSSIAN, dstate.capacity, dstate.user_capacity)
    R.flush_checks()
    R.set_pair_count_mode("sync")
    R._MIN_CAPACITY, R._MIN_PAIRS_PER_GAUSSIAN = 1024, 0
    R.set_workspace_capacity(1024)
    try:
        cu = run_cuda(sc, cam, deg)
    finally:
        R._pair_mode, R._MIN_CAPACITY, R._MIN_PAIRS_PER_GAUSSIAN, dstate.capacity, dstate.user_capacity = old
    np.testing.assert_array_equal(cu["radii"].cpu().numpy(), ref["radii"].numpy())
    check_lists(sc, cam, ref)
    check_images(cu, ref)


def test_equal_depth_runs_and_skewed_depths_sort_exactly():
    """100 exact duplicates of each of 80 positions (runs of equal depth bits, ordered by index)
    plus two tight depth clusters (skewed bucket occupancy -> radix fallback)."""
    P0, dup, H, W = 80, 100, 64, 64
    sc, cam, deg = U.make_inputs(P0, H, W, seed=19, radius=0.05, exact_knn=False, scale_mul=40.0)
    sc = {k: v.repeat_interleave(dup, dim=0).contiguous() for k, v in sc.items()}
    # second cluster: same duplicates pushed away from the camera along the view axis
    far = {k: v.clone() for k, v in sc.items()}
    fwd = -cam.camera_center / cam.camera_center.norm()
    far["means3D"] = far["means3D"] + fwd * 1.5
    sc = {k: torch.cat([sc[k], far[k]], dim=0).contiguous() for k in sc}
    ref = run_oracle(sc, cam, deg)
    n = ref["ranges"][:, 1] - ref["ranges"][:, 0]
    assert n.max() > 4096
    cu = run_cuda(sc, cam, deg)
    np.testing.assert_array_equal(cu["radii"].cpu().numpy(), ref["radii"].numpy())
    check_lists(sc, cam, ref)
    check_images(cu, ref)


def test_screen_filling_gaussians_scale_modifier_and_score():
    """A few huge Gaussians (tile rect = whole image, thousands of tiles each) mixed with small ones,
    scale_modifier != 1, score_flag on, 1080p-like aspect with partial tiles."""
    H, W = 135, 240
    sc, cam, deg = U.make_inputs(3000, H, W, seed=29)
    sc["scales"][:5] = 2.0                      # bigger than the scene: cover every tile
    sc["opacities"][:5] = 0.3
    g = torch.Generator().manual_seed(3)
    grads = (torch.randn(3, H, W, generator=g) / (H * W), torch.randn(2, H, W, generator=g) / (H * W))
    ref32 = run_oracle(sc, cam, deg, score=True, scale_modifier=0.7)
    assert int(ref32["pre"]["touched"].max()) == ((H + 15) // 16) * ((W + 15) // 16)
    ref = run_oracle(sc, cam, deg, grads=grads, dtype=torch.float64, decisions=ref32["decisions"], scale_modifier=0.7)
    cu = run_cuda(sc, cam, deg, score=True, grads=grads, scale_modifier=0.7)
    np.testing.assert_array_equal(cu["radii"].cpu().numpy(), ref32["radii"].numpy())
    check_images(cu, ref32)
    assert U.rel_err(cu["score"], ref32["score"]) < 1e-4
    for k in ("means3D", "scales", "rotations", "opacities", "shs", "means2D"):
        assert U.rel_err(cu["grads"][k], ref["grads"][k]) < 1e-3, k


def test_single_gaussian_and_random_background():
    sc, cam, deg = U.make_inputs(1, 64, 64, seed=31)
    sc["means3D"][:] = 0.0
    sc["scales"][:] = torch.tensor([0.08, 0.05, 0.03])
    sc["opacities"][:] = 0.8
    bg = (0.1, 0.7, 0.4)
    ref = run_oracle(sc, cam, deg, bg=bg)
    cu = run_cuda(sc, cam, deg, bg=bg, grads=(torch.ones(3, 64, 64), torch.zeros(2, 64, 64)))
    assert int(cu["radii"][0]) == int(ref["radii"][0]) > 0
    check_images(cu, ref)
    assert torch.isfinite(cu["grads"]["means3D"]).all()


@pytest.mark.parametrize("P", [2, 3, 7, 1001])
def test_odd_point_counts_forward_backward(P):
    """Gradient sections of the flat buffer must stay 16-byte aligned for any P."""
    sc, cam, deg = U.make_inputs(P, 48, 48, seed=37 + P, exact_knn=False, scale_mul=3.0)
    g = torch.Generator().manual_seed(P)
    grads = (torch.randn(3, 48, 48, generator=g), torch.randn(2, 48, 48, generator=g))
    ref32 = run_oracle(sc, cam, deg, record=True)
    ref = run_oracle(sc, cam, deg, grads=grads, dtype=torch.float64, decisions=ref32["decisions"])
    cu = run_cuda(sc, cam, deg, grads=grads)
    check_images(cu, ref32)
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        assert U.rel_err(cu["grads"][k], ref["grads"][k]) < 1e-3, k
    # also through an unaligned view of a larger tensor
    from dreamscene_b200 import GaussianRasterizer
    S = U.cuda_settings(cam, deg)
    big = torch.zeros(P * 3 + 1, device="cuda")
    big[1:] = sc["means3D"].reshape(-1).cuda()
    t = {k: v.cuda() for k, v in sc.items()}
    c2, r2, d2 = GaussianRasterizer(S)(means3D=big[1:].view(P, 3), means2D=torch.zeros(P, 3, device="cuda"),
                                       opacities=t["opacities"], shs=t["shs"], scales=t["scales"],
                                       rotations=t["rotations"])
    assert torch.equal(c2.cpu(), cu["color"].cpu())


def test_large_tile_grid_uses_the_global_atomic_binning_fallback():
    """1792x1792 = 12544 tiles > the smem multisplit limit (12288): exercises the privatised
    global-counter path; lists and images must still match."""
    H = W = 1792
    sc, cam, deg = U.make_inputs(800, H, W, seed=41, exact_knn=False, scale_mul=0.5)
    ref = run_oracle(sc, cam, deg)
    cu = run_cuda(sc, cam, deg)
    np.testing.assert_array_equal(cu["radii"].cpu().numpy(), ref["radii"].numpy())
    check_lists(sc, cam, ref)
    check_images(cu, ref)


def test_tile_grid_at_the_multisplit_limit():
    """2048x1536 = exactly 12288 tiles: the largest grid on the shared-memory binning path."""
    H, W = 1536, 2048
    sc, cam, deg = U.make_inputs(600, H, W, seed=43, exact_knn=False, scale_mul=0.5)
    ref = run_oracle(sc, cam, deg)
    cu = run_cuda(sc, cam, deg)
    np.testing.assert_array_equal(cu["radii"].cpu().numpy(), ref["radii"].numpy())
    check_lists(sc, cam, ref)
    check_images(cu, ref)


def test_api_errors_match_reference_messages():
    from dreamscene_b200 import GaussianRasterizer
    sc, cam, deg = U.make_inputs(16, 32, 32)
    S = U.cuda_settings(cam, deg)
    t = {k: v.cuda() for k, v in sc.items()}
    z = torch.zeros(16, 3, device="cuda")
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        GaussianRasterizer(S)(means3D=t["means3D"], means2D=z, opacities=t["opacities"],
                              scales=t["scales"], rotations=t["rotations"])
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        GaussianRasterizer(S)(means3D=t["means3D"], means2D=z, opacities=t["opacities"], shs=t["shs"])
