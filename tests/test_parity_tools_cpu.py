"""CPU checks of the checker: the tile-sampled, group-wise oracle backward used for the full-size
GPU parity tests must equal the plain whole-frame oracle autograd with the same masked gradients."""
import numpy as np
import torch

from oracle import splat_ref as O
from tests import parity_tools as PT
from tests import util_scene as U


def test_tile_sampled_oracle_backward_equals_whole_frame_autograd():
    H = W = 64
    sc, cam, deg = U.make_inputs(400, H, W, seed=3)
    S, pre, keys, pl, ranges, dec = PT.oracle_lists(sc, cam, deg)
    tiles = PT.sample_tiles(ranges, 5)
    assert 1 <= len(tiles) <= 5 and len(set(tiles)) == len(tiles)
    m = PT.tile_mask(tiles, H, W)
    assert int(m.sum()) == 256 * len(tiles)
    g = torch.Generator().manual_seed(0)
    gc = torch.randn(3, H, W, generator=g) * m
    gd = torch.randn(2, H, W, generator=g) * m
    got = PT.oracle_backward_on_tiles(sc, cam, deg, tiles, gc, gd, dec, dtype=torch.float64, group=2)
    # reference: whole frame, fp64 on the same fp32 decisions
    t = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    m2d = torch.zeros(400, 3, requires_grad=True)
    r = O.rasterize(S, t["means3D"], t["opacities"], shs=t["shs"], scales=t["scales"], rotations=t["rotations"],
                    means2D=m2d, dtype=torch.float64, decisions=dec)
    ((r["color"] * gc.double()).sum() + (r["depth_alpha"] * gd.double()).sum()).backward()
    for k in ("means3D", "opacities", "shs", "scales", "rotations"):
        assert U.rel_err(got[k], t[k].grad) < 1e-12, k
    assert U.rel_err(got["means2D"], m2d.grad) < 1e-12


def test_forward_stats_counts_outliers_inside_the_mask_only():
    H = W = 32
    a = torch.zeros(3, H, W); b = torch.zeros(3, H, W)
    da = torch.zeros(2, H, W); db = torch.zeros(2, H, W)
    a[0, 0, 0] = 5e-4          # inside tile 0
    a[1, 20, 20] = 1.0         # outside the mask
    da[1, 3, 3] = 2e-4
    m = PT.tile_mask([0], H, W)
    st = PT.forward_stats(a, da, b, db, cu_nc=np.zeros((H, W)), ref_nc=np.zeros((H, W)), mask=m)
    assert st["color"]["n"] == 3 * 256 and st["color"]["n_bad"] == 1 and abs(st["color"]["max_abs"] - 5e-4) < 1e-9
    assert st["T"]["n_bad"] == 1 and st["depth"]["n_bad"] == 0 and st["n_contrib"]["n_mismatch"] == 0
    full = PT.forward_stats(a, da, b, db)
    assert full["color"]["n_bad"] == 2 and full["color"]["max_abs"] == 1.0
