"""Full-size checks at BASELINE.json's headline configuration (1M Gaussians, 1024x1024): the
oracle is too slow to render the whole frame, so the CUDA path is checked through size-independent
properties plus an oracle comparison on a sample of tiles."""
import numpy as np
import pytest
import torch

from oracle import splat_ref as O
from tests import util_scene as U

pytestmark = pytest.mark.gpu

P, H, W = 1_000_000, 1024, 1024


@pytest.fixture(scope="module")
def big():
    sc, cam, deg = U.make_inputs(P, H, W, seed=0)     # the bench workload (exact 3-NN scales)
    dev = torch.device("cuda", torch.cuda.current_device())
    t = {k: v.to(dev) for k, v in sc.items()}
    return sc, cam, deg, t, dev


def _forward(t, cam, deg, dev, bg=(1.0, 1.0, 1.0)):
    from dreamscene_b200 import rasterizer as R
    S = U.cuda_settings(cam, deg, bg, device=dev)
    with torch.no_grad():
        color, radii, da, _, st = R._forward_impl(S, t["means3D"], t["shs"], None, t["opacities"], t["scales"],
                                                  t["rotations"], None)
    torch.cuda.synchronize()
    return color, radii, da, st


def test_lists_are_tile_major_depth_sorted_and_complete(big):
    sc, cam, deg, t, dev = big
    color, radii, da, st = _forward(t, cam, deg, dev)
    dec = U.decode_saved(st.saved, P, H, W, st.capacity)
    D = dec["num_pairs"]
    ts = dec["tile_start"]
    assert ts[0] == 0 and ts[-1] == D and np.all(np.diff(ts) >= 0)
    # every Gaussian appears once per touched tile: sum of rect areas == D (checksum of checksums)
    rect_area = np.bincount(dec["idx"], minlength=P)
    assert rect_area.sum() == D and (rect_area[radii.cpu().numpy() == 0] == 0).all()
    assert (rect_area[radii.cpu().numpy() > 0] > 0).all()
    # inside every tile the 64-bit keys (depth bits, idx) are strictly increasing
    keys = (dec["depth_bits"].astype(np.uint64) << np.uint64(32)) | dec["idx"].astype(np.uint64)
    inc = keys[1:] > keys[:-1]
    boundary = np.zeros(D - 1, bool)
    inner = ts[1:-1]
    boundary[inner[(inner > 0) & (inner < D)] - 1] = True
    assert np.all(inc | boundary)
    # the depth bits stored in the key are the Gaussian's view depth (geom record field 7)
    np.testing.assert_array_equal(dec["depth_bits"], dec["geom_f32"].view(np.uint32)[dec["idx"], 7])
    # outputs are sane: transmittance in [0,1], colour finite, n_contrib within the tile's list
    T = da[1].cpu().numpy()
    assert T.min() >= 0.0 and T.max() <= 1.0 and np.isfinite(color.cpu().numpy()).all()
    nc = dec["n_contrib"].astype(np.int64)
    n_tile = np.diff(ts).reshape(H // 16, W // 16)
    assert (nc.reshape(H // 16, 16, W // 16, 16).max(axis=(1, 3)) <= n_tile).all()


def test_forward_is_deterministic_and_background_enters_linearly(big):
    sc, cam, deg, t, dev = big
    c1, r1, da1, _ = _forward(t, cam, deg, dev, bg=(1.0, 1.0, 1.0))
    c2, r2, da2, _ = _forward(t, cam, deg, dev, bg=(1.0, 1.0, 1.0))
    assert torch.equal(c1, c2) and torch.equal(da1, da2) and torch.equal(r1, r2)
    c0, _, da0, _ = _forward(t, cam, deg, dev, bg=(0.0, 0.0, 0.0))
    assert torch.equal(da0, da1)
    # colour(bg) = C + T*bg  =>  colour(1) - colour(0) = T for every channel
    for ch in range(3):
        assert torch.allclose(c1[ch] - c0[ch], da1[1], atol=2e-7)


def test_backward_is_linear_in_the_incoming_gradient(big):
    from dreamscene_b200 import GaussianRasterizer
    sc, cam, deg, t, dev = big
    S = U.cuda_settings(cam, deg, device=dev)
    g = torch.Generator().manual_seed(5)
    gc = (torch.randn(3, H, W, generator=g) / (H * W)).to(dev)
    gd = (torch.randn(2, H, W, generator=g) / (H * W)).to(dev)

    def grads(scale):
        p = {k: v.clone().requires_grad_(True) for k, v in t.items()}
        m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
        color, radii, da = GaussianRasterizer(S)(means3D=p["means3D"], means2D=m2d, opacities=p["opacities"],
                                                 shs=p["shs"], scales=p["scales"], rotations=p["rotations"])
        torch.autograd.backward([color, da], [gc * scale, gd * scale])
        return {k: v.grad for k, v in p.items()}, m2d.grad

    g1, m1 = grads(1.0)
    g2, m2 = grads(2.0)
    for k in g1:
        assert torch.isfinite(g1[k]).all()
        assert U.rel_err(g2[k], 2.0 * g1[k]) < 1e-5, k
    assert U.rel_err(m2, 2.0 * m1) < 1e-5
    assert float(m1[:, 2].abs().max()) == 0.0


def _arbitrate_radii(sc, cam, cuda_radii, oracle_radii):
    """The torch oracle and the CUDA kernel evaluate the same IEEE fp32 operation sequence, so the
    radii must be identical.  If they ever differ (seen once, on one host CPU type, for 1 of 1e6
    Gaussians), dump the inputs for offline analysis and let the numpy-fp32 scalar arbiter decide
    which side deviates: CUDA deviating is a failure; the torch oracle deviating on this host is
    reported as a skip (the checker, not the product, is off)."""
    bad = np.nonzero(cuda_radii != oracle_radii)[0]
    if bad.size == 0:
        return
    import os
    from oracle import fp32_arbiter as A
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez("gpurun_out/radii_mismatch.npz", idx=bad, means3D=sc["means3D"][bad].numpy(),
             scales=sc["scales"][bad].numpy(), rotations=sc["rotations"][bad].numpy(), cuda=cuda_radii[bad],
             oracle=oracle_radii[bad], view=cam.world_view_transform.numpy(), proj=cam.full_proj_transform.numpy(),
             tanfov=np.array([cam.tanfovx, cam.tanfovy]))
    arb = np.array([A.radius_rect(sc["means3D"][i].numpy(), sc["scales"][i].numpy(), sc["rotations"][i].numpy(),
                                  cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(), cam.tanfovx,
                                  cam.tanfovy, H, W)["radius"] for i in bad[:64]])
    assert np.array_equal(arb, cuda_radii[bad[:64]]), \
        f"CUDA radii deviate from the numpy-fp32 arbiter for Gaussians {bad[:8]}: {cuda_radii[bad[:8]]} vs {arb[:8]}"
    pytest.skip(f"torch oracle deviates from both CUDA and the numpy-fp32 arbiter on this host for "
                f"{bad.size} of {P} Gaussians (inputs dumped to gpurun_out/radii_mismatch.npz)")


def test_sampled_tiles_match_the_oracle_at_full_size(big):
    """cfg3: complete sorted list bit-exact; forward on 24 list-length-stratified tiles within the
    measured budgets; COMPLETE parameter gradients vs the fp64 oracle with the incoming gradients
    masked to those tiles (long lists: fp32 atomic accumulation over thousands of contributors,
    rcp.approx drift over hundreds of back-to-front steps, the ring wrap)."""
    from tests import parity_budgets as B
    from tests import parity_tools as PT
    sc, cam, deg, t, dev = big
    color, radii, da, st = _forward(t, cam, deg, dev)
    S, pre, keys, pl, ranges, odec = PT.oracle_lists(sc, cam, deg)
    _arbitrate_radii(sc, cam, radii.cpu().numpy(), pre["radii"].numpy())
    dec = U.decode_saved(st.saved, P, H, W, st.capacity)
    assert dec["num_pairs"] == len(pl)
    np.testing.assert_array_equal(dec["idx"], pl)                      # the full multi-million-entry sorted list
    np.testing.assert_array_equal(dec["tile_start"][:-1], ranges[:, 0])
    tiles = PT.sample_tiles(ranges, 24)
    mask = PT.tile_mask(tiles, H, W)
    with torch.no_grad():
        oc, oda, onc, _ = O.composite(pre, pl, ranges, S, tiles=tiles)
    stats = PT.forward_stats(color, da, oc, oda, cu_nc=dec["n_contrib"], ref_nc=onc.numpy(), mask=mask)
    B.check_forward("cfg3_1M_1024", stats)
    # backward: gradients flow only into the sampled tiles, the parameter gradient is compared in full
    g = torch.Generator().manual_seed(13)
    gc = torch.randn(3, H, W, generator=g) / (H * W) * mask
    gd = torch.randn(2, H, W, generator=g) / (H * W) * mask
    blend = PT.record_blend_decisions(pre, odec, S, tiles)        # fp32 per-pixel decisions, replayed in fp64
    want = PT.oracle_backward_on_tiles(sc, cam, deg, tiles, gc, gd, odec, group=3, blend=blend)
    got = PT.cuda_forward_backward(sc, cam, deg, gc, gd, device=dev)
    errs = PT.grad_errors(got["grads"], want)
    for k in ("means3D", "scales", "rotations", "opacities", "shs", "means2D"):
        assert errs[k]["rel_l2"] < B.BWD_REL, (k, errs[k])
