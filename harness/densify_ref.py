"""The reference's densification logic restated functionally in PyTorch (harness code: the checker for
dreamscene_b200.densify; follows /root/reference/gs_renderer.py:868-1081 statement by statement, with
dicts in place of the optimizer and the torch.normal draws passed in as standard normals z)."""
import torch

NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def build_rotation(r):        # gs_renderer.py:124-147
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r_ * z); R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y); R[:, 2, 1] = 2 * (y * z + r_ * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _cat(p, a, new):          # cat_tensors_to_optimizer
    p = {k: torch.cat((p[k], new[k]), dim=0) for k in NAMES}
    if a is not None:
        a = {k: tuple(torch.cat((m, torch.zeros_like(new[k])), dim=0) for m in a[k]) for k in NAMES}
    return p, a


def _prune(p, a, mask):       # prune_points / _prune_optimizer
    keep = ~mask
    p = {k: v[keep] for k, v in p.items()}
    if a is not None:
        a = {k: tuple(m[keep] for m in v) for k, v in a.items()}
    return p, a


def densify_and_prune(params, adam, accum, denom, max_grad, min_opacity, extent, max_screen_size, percent_dense, N, z):
    p = {k: params[k].clone() for k in NAMES}
    a = None if adam is None else {k: tuple(m.clone() for m in adam[k]) for k in NAMES}
    grads = accum / denom
    grads[grads.isnan()] = 0.0
    # densify_and_clone
    sel = torch.where(torch.norm(grads, dim=-1) >= max_grad, True, False)
    sel = torch.logical_and(sel, torch.max(torch.exp(p["scaling"]), dim=1).values <= percent_dense * extent)
    p, a = _cat(p, a, {k: p[k][sel] for k in NAMES})
    # densify_and_split (grads padded with zeros for the clones)
    n_init = p["xyz"].shape[0]
    padded = torch.zeros(n_init, device=grads.device)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = torch.where(padded >= max_grad, True, False)
    sel = torch.logical_and(sel, torch.max(torch.exp(p["scaling"]), dim=1).values > percent_dense * extent)
    stds = torch.exp(p["scaling"])[sel].repeat(N, 1)
    samples = z[:stds.shape[0]] * stds                                  # torch.normal(mean=0, std=stds)
    rots = build_rotation(p["rotation"][sel]).repeat(N, 1, 1)
    new = {"xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + p["xyz"][sel].repeat(N, 1),
           "scaling": torch.log(torch.exp(p["scaling"])[sel].repeat(N, 1) / (0.8 * N)),
           "rotation": p["rotation"][sel].repeat(N, 1), "f_dc": p["f_dc"][sel].repeat(N, 1, 1),
           "f_rest": p["f_rest"][sel].repeat(N, 1, 1), "opacity": p["opacity"][sel].repeat(N, 1)}
    p, a = _cat(p, a, new)
    p, a = _prune(p, a, torch.cat((sel, torch.zeros(N * int(sel.sum()), device=sel.device, dtype=bool))))
    # final prune; max_radii2D was reset to zeros by densification_postfix
    prune_mask = (torch.sigmoid(p["opacity"]) < min_opacity).squeeze()
    if max_screen_size:
        big_vs = torch.zeros(p["xyz"].shape[0], device=sel.device) > max_screen_size
        big_ws = torch.exp(p["scaling"]).max(dim=1).values > 0.1 * extent
        prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_vs), big_ws)
    return _prune(p, a, prune_mask)
