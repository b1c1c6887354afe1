"""CPU-only property tests of the oracle (hypothesis): invariants any correct restatement of the
rasterizer must satisfy, on small random scenes.  They guard the checker itself."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import splat_ref as O
from tests import util_scene as U


def _render(P, H, W, seed, deg, bg=(1.0, 1.0, 1.0), score=False, **kw):
    sc, cam, _ = U.make_inputs(P, H, W, seed=seed, exact_knn=False, **kw)
    S = U.oracle_settings(cam, deg, bg, score)
    with torch.no_grad():
        r = O.rasterize(S, sc["means3D"], sc["opacities"], shs=sc["shs"], scales=sc["scales"],
                        rotations=sc["rotations"])
    return sc, cam, r


@settings(max_examples=8, deadline=None)
@given(seed=st.integers(0, 10_000), P=st.integers(1, 400), H=st.integers(8, 70), W=st.integers(8, 70),
       deg=st.integers(0, 3))
def test_lists_sorted_ranges_partition_and_outputs_bounded(seed, P, H, W, deg):
    sc, cam, r = _render(P, H, W, seed, deg)
    keys, pl, ranges = r["keys"], r["point_list"], r["ranges"]
    D = len(pl)
    assert D == int(r["pre"]["touched"].sum())
    if D:
        assert np.all(keys[1:] >= keys[:-1])                       # tile-major, depth-sorted
        same = keys[1:] == keys[:-1]
        assert np.all(pl[1:][same] > pl[:-1][same])                # equal keys keep index order (stable)
    assert ranges[0, 0] == 0 and ranges[-1, 1] == D
    assert np.all(ranges[1:, 0] == ranges[:-1, 1])                 # ranges partition the list
    T = r["depth_alpha"][1]
    assert float(T.min()) >= 0.0 and float(T.max()) <= 1.0
    assert float(r["depth_alpha"][0].min()) >= 0.0                  # depths are > 0.2, weights >= 0
    assert torch.isfinite(r["color"]).all()
    vis = r["radii"] > 0
    assert torch.equal(vis, r["pre"]["visible"])
    n = ranges[:, 1] - ranges[:, 0]
    gx = (W + 15) // 16
    nc = r["n_contrib"].numpy()
    for t in np.nonzero(n)[0][:6]:
        ty, tx = divmod(int(t), gx)
        assert nc[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16].max() <= n[t]


@settings(max_examples=5, deadline=None)
@given(seed=st.integers(0, 10_000), P=st.integers(5, 300))
def test_background_enters_linearly_and_score_is_total_blend_weight(seed, P):
    _, _, r1 = _render(P, 40, 48, seed, 2, bg=(1.0, 1.0, 1.0), score=True)
    _, _, r0 = _render(P, 40, 48, seed, 2, bg=(0.0, 0.0, 0.0))
    T = r1["depth_alpha"][1]
    for c in range(3):
        assert torch.allclose(r1["color"][c] - r0["color"][c], T, atol=1e-6)
    # sum of importance scores == sum over pixels of (1 - T) when every weight is counted once
    assert abs(float(r1["score"].sum()) - float((1.0 - T).sum())) < 1e-3 * max(1.0, float((1.0 - T).sum()))


def test_culled_gaussians_get_exactly_zero_gradients():
    sc, cam, deg = U.make_inputs(200, 32, 32, seed=3)
    sc["means3D"][::2] += cam.camera_center * 2.0            # every other Gaussian behind the camera
    S = U.oracle_settings(cam, deg)
    t = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    r = O.rasterize(S, t["means3D"], t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
    (r["color"].sum() + r["depth_alpha"].sum()).backward()
    culled = r["radii"] == 0
    assert culled[::2].all()
    for k, v in t.items():
        assert float(v.grad[culled].abs().max()) == 0.0, k


def test_fp64_run_on_fp32_decisions_agrees_with_fp32_run():
    sc, cam, deg = U.make_inputs(300, 48, 48, seed=9)
    S = U.oracle_settings(cam, deg)
    kw = dict(shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    with torch.no_grad():
        r32 = O.rasterize(S, sc["means3D"], sc["opacities"], **kw)
        r64 = O.rasterize(S, sc["means3D"], sc["opacities"], dtype=torch.float64, decisions=r32["decisions"], **kw)
    assert (r64["color"].float() - r32["color"]).abs().max() < 5e-3
    assert ((r64["color"].float() - r32["color"]).abs() > 1e-4).float().mean() < 1e-3
