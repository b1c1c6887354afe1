"""the fused scene assembly (activations + cat + augmentation, one kernel each way)
against the reference's PyTorch expressions (harness/scene_ref.py restates scene_gaussian.py:753-857)."""
import pytest
import torch

from harness.scene_ref import reference_assemble

pytestmark = pytest.mark.gpu


def _groups(sizes, M, seed=0, requires_grad=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    out = []
    for n in sizes:
        d = {"_xyz": r(n, 3), "_opacity": r(n, 1) * 2, "_scaling": r(n, 3) * 0.5 - 3.0, "_rotation": r(n, 4),
             "_features_dc": r(n, 1, 3), "_features_rest": r(n, M - 1, 3) * 0.1}
        out.append({k: v.requires_grad_(requires_grad) for k, v in d.items()})
    return out


@pytest.mark.parametrize("M,sizes", [(4, [1000, 1, 513, 4097]), (16, [300, 777]), (9, [129, 64]), (1, [50])])
def test_assemble_matches_reference_expressions_forward_and_backward(M, sizes):
    from dreamscene_b200.scene import assemble_scene
    groups = _groups(sizes, M)
    P = sum(sizes)
    gen = torch.Generator(device="cuda").manual_seed(7)
    drawn = assemble_scene(groups, noise="torch", generator=gen)
    gen = torch.Generator(device="cuda").manual_seed(7)          # the reference's draws: shs first, then scales
    z_shs = torch.randn(P, M, 3, device="cuda", generator=gen)
    z_scales = torch.randn(P, 3, device="cuda", generator=gen)
    same_stream = reference_assemble(groups, z_shs, z_scales)
    assert torch.equal(drawn[4], same_stream[4]) and torch.equal(drawn[2], same_stream[2])   # same RNG stream as the reference
    z_scales[::5] = -10.0                                        # 1 + 0.1118 z < 0: exercises the clamp at exactly 0
    got = assemble_scene(groups, noise="torch", z_shs=z_shs, z_scales=z_scales)
    ref_groups = [{k: v.detach().clone().requires_grad_(True) for k, v in g.items()} for g in groups]
    want = reference_assemble(ref_groups, z_shs, z_scales)
    names = ("means3D", "opacity", "scales", "rotations", "shs")
    for a, b, k in zip(got, want, names):
        assert a.shape == b.shape, k
        if k == "rotations":
            assert torch.allclose(a, b, rtol=0, atol=2e-7), k        # sum-of-squares order: <= 1 ulp
        else:
            assert torch.equal(a, b), (k, float((a - b).abs().max()))
    assert float(got[2].min()) >= 0.0 and bool((got[2] == 0).any())   # the clamp is exercised
    w = [torch.randn_like(t) for t in got]
    torch.autograd.backward(list(got), w)
    torch.autograd.backward(list(want), w)
    for g, r in zip(groups, ref_groups):
        for k in g:
            if g[k].numel() == 0:
                continue
            err = float((g[k].grad - r[k].grad).abs().max() / r[k].grad.abs().max().clamp_min(1e-30))
            assert err < 2e-6, (k, err)


def test_assemble_without_augmentation_and_fused_noise_statistics():
    from dreamscene_b200.scene import assemble_scene
    groups = _groups([20000, 30000], 4, seed=1)
    plain = assemble_scene(groups, shs_aug=False, scale_aug=False)
    want = reference_assemble(groups)
    for a, b, k in zip(plain, want, range(5)):
        assert torch.allclose(a, b, rtol=0, atol=2e-7), k
    # in-kernel Philox noise: same seed -> same output, different seed -> different; the implied
    # standard normals have mean 0 / variance 1 and the backward regenerates exactly the forward's factors
    a1 = assemble_scene(groups, noise="fused", seed=123)
    a2 = assemble_scene(groups, noise="fused", seed=123)
    a3 = assemble_scene(groups, noise="fused", seed=124)
    assert torch.equal(a1[4], a2[4]) and torch.equal(a1[2], a2[2]) and not torch.equal(a1[4], a3[4])
    z = (a1[4] / plain[4] - 1.0) / (0.2 ** 0.5)                   # shs = v (1 + c z)
    z = z[torch.isfinite(z) & (plain[4].abs() > 1e-3)]
    assert abs(float(z.mean())) < 0.01 and abs(float(z.var()) - 1.0) < 0.02
    assert abs(float((z ** 4).mean()) - 3.0) < 0.15              # Gaussian kurtosis
    for g in groups:
        for v in g.values():
            v.grad = None
    a1[4].sum().backward()
    dshs = torch.cat([torch.cat((g["_features_dc"].grad, g["_features_rest"].grad), dim=1) for g in groups])
    assert torch.allclose(dshs, a1[4].detach() / plain[4].detach(), rtol=2e-5, atol=1e-4)   # d(v(1+cz))/dv = 1+cz


def test_assemble_feeds_the_rasterizer_end_to_end():
    from dreamscene_b200 import GaussianRasterizer
    from dreamscene_b200.scene import assemble_scene
    from tests import util_scene as U
    sc, cam, deg = U.make_inputs(3000, 96, 96, seed=5, sh_max=1)
    raw = {"_xyz": sc["means3D"], "_opacity": torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)), "_scaling": sc["scales"].log(),
           "_rotation": sc["rotations"] * 1.7, "_features_dc": sc["shs"][:, :1], "_features_rest": sc["shs"][:, 1:]}
    halves = [{k: v[:1234].cuda().contiguous().requires_grad_(True) for k, v in raw.items()},
              {k: v[1234:].cuda().contiguous().requires_grad_(True) for k, v in raw.items()}]
    m, o, s, r, f = assemble_scene(halves, shs_aug=False, scale_aug=False)
    S = U.cuda_settings(cam, deg)
    color, radii, da = GaussianRasterizer(S)(means3D=m, means2D=torch.zeros_like(m, requires_grad=True), opacities=o,
                                             shs=f, scales=s, rotations=r)
    color.sum().backward()
    t = {k: v.cuda() for k, v in sc.items()}
    c2, _, _ = GaussianRasterizer(S)(means3D=t["means3D"], means2D=torch.zeros_like(m), opacities=t["opacities"], shs=t["shs"],
                                     scales=t["scales"], rotations=t["rotations"])
    assert torch.allclose(color, c2, atol=2e-5)
    assert all(torch.isfinite(v.grad).all() and float(v.grad.abs().max()) > 0 for h in halves for v in h.values())


def test_assemble_multi_view_equals_per_view_calls():
    """views=B: one pass over the raw parameters, B independently augmented copies of shs / scales; values and leaf
    gradients equal B separate single-view calls with the same draws (torch noise) / are reproducible (Philox)."""
    from dreamscene_b200.scene import assemble_scene
    B, M = 3, 4
    groups = _groups([700, 1300, 50], M, seed=9)
    P = 2050
    gen = torch.Generator(device="cuda").manual_seed(11)
    zs = torch.randn(B, P, M, 3, device="cuda", generator=gen)
    zc = torch.randn(B, P, 3, device="cuda", generator=gen)
    m, o, sc, r, sh = assemble_scene(groups, noise="torch", z_shs=zs, z_scales=zc, views=B)
    assert sc.shape == (B, P, 3) and sh.shape == (B, P, M, 3)
    w_sc, w_sh = torch.randn_like(sc), torch.randn_like(sh)
    w_m, w_o, w_r = torch.randn_like(m), torch.randn_like(o), torch.randn_like(r)
    torch.autograd.backward([m, o, sc, r, sh], [w_m, w_o, w_sc, w_r, w_sh])
    got = [{k: v.grad.clone() for k, v in g.items()} for g in groups]
    for g in groups:
        for v in g.values():
            v.grad = None
    for v in range(B):
        m1, o1, sc1, r1, sh1 = assemble_scene(groups, noise="torch", z_shs=zs[v], z_scales=zc[v])
        assert torch.equal(sc1, sc[v]) and torch.equal(sh1, sh[v]) and torch.equal(m1, m) and torch.equal(r1, r)
        torch.autograd.backward([m1, o1, sc1, r1, sh1],
                                [w_m if v == 0 else torch.zeros_like(w_m), w_o if v == 0 else torch.zeros_like(w_o), w_sc[v],
                                 w_r if v == 0 else torch.zeros_like(w_r), w_sh[v]])
    for g, gg in zip(groups, got):
        for k in g:
            err = float((gg[k] - g[k].grad).abs().max() / g[k].grad.abs().max().clamp_min(1e-30))
            assert err < 5e-6, (k, err)
    a = assemble_scene(groups, noise="fused", seed=5, views=B)
    b = assemble_scene(groups, noise="fused", seed=5, views=B)
    assert torch.equal(a[4], b[4]) and not torch.equal(a[4][0], a[4][1])         # reproducible, views differ
    single = assemble_scene(groups, noise="fused", seed=5)
    assert torch.equal(single[4], a[4][0]) and torch.equal(single[2], a[2][0])    # view 0 == the single-view stream
