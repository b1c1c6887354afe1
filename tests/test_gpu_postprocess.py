"""SURVEY 8(f1, post-processing): fused disparity vs the reference's PyTorch expression
(/root/reference/scene_gaussian.py:871-881), values and gradients, single view and batch."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_disparity(depth_alpha, focal):
    depth, alpha = torch.chunk(depth_alpha, 2)
    disp = focal / (depth + (alpha * 10) + 1e-5)
    try:
        min_d = disp[alpha <= 0.1].min()
    except Exception:
        min_d = disp.min()
    disp = torch.clamp((disp - min_d) / (disp.max() - min_d), 0.0, 1.0)
    return disp, alpha


def _fake_depth_alpha(H, W, seed, opaque=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    T = torch.rand(H, W, device="cuda", generator=g)
    if opaque:
        T = torch.where(torch.rand(H, W, device="cuda", generator=g) < 0.5, T * 0.05, T)
    else:
        T = 0.2 + 0.8 * T                         # no pixel with alpha <= 0.1: the fallback branch
    depth = (1.0 - T) * (2.0 + 3.0 * torch.rand(H, W, device="cuda", generator=g))
    return torch.stack([depth, T])


@pytest.mark.parametrize("opaque", [True, False])
def test_disparity_matches_reference_expression_and_autograd(opaque):
    from dreamscene_b200.postprocess import disparity_from_depth_alpha
    H, W = 96, 80
    focal = 1 / (2 * math.tan(0.96 / 2))
    da = _fake_depth_alpha(H, W, 3, opaque)
    a = da.clone().requires_grad_(True)
    b = da.clone().requires_grad_(True)
    d1, al1 = disparity_from_depth_alpha(a, focal)
    d2, al2 = reference_disparity(b, focal)
    assert d1.shape == d2.shape == (1, H, W) and torch.equal(al1, al2)
    assert torch.allclose(d1, d2, rtol=0, atol=2e-6)
    assert float(d1.min()) == 0.0 and float(d1.max()) == 1.0
    g = torch.Generator(device="cuda").manual_seed(9)
    gd, ga = torch.randn(1, H, W, device="cuda", generator=g), torch.randn(1, H, W, device="cuda", generator=g)
    torch.autograd.backward([d1, al1], [gd, ga])
    torch.autograd.backward([d2, al2], [gd, ga])
    err = float((a.grad - b.grad).abs().max() / b.grad.abs().max())
    assert err < 1e-4, err


def test_disparity_batch_equals_per_view_and_backward_is_repeatable():
    from dreamscene_b200.postprocess import disparity_from_depth_alpha
    H, W, B = 64, 64, 4
    focals = [1 / (2 * math.tan(f / 2)) for f in (0.55, 0.96, 0.7, 0.96)]
    das = torch.stack([_fake_depth_alpha(H, W, 10 + k, k != 2) for k in range(B)]).requires_grad_(True)
    disp, alpha = disparity_from_depth_alpha(das, focals)
    assert disp.shape == (B, 1, H, W)
    for k in range(B):
        d, a = reference_disparity(das[k].detach(), focals[k])
        assert torch.allclose(disp[k], d, rtol=0, atol=2e-6) and torch.equal(alpha[k], a)
    loss = (disp * disp).sum()
    g1 = torch.autograd.grad(loss, das, retain_graph=True)[0]
    g2 = torch.autograd.grad(loss, das)[0]
    assert torch.allclose(g1, g2, rtol=1e-5, atol=1e-8)
