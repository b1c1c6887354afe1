"""Deterministic synthetic Gaussian scenes for tests and bench.

Recipes restate the reference's initialisers (no reference import):
  * positions: uniform ball, gs_renderer.py:359-367
  * scales: sqrt(mean squared distance to 3 nearest neighbours), gs_renderer.py:590-594
    (exact k-d tree 3-NN by default = the reference recipe; `exact_knn=False` substitutes the
    expected 3-NN distance of a uniform ball, which is what round 1 benchmarked and gives ~30%
    more tile pairs), times exp(N(0,0.3)) anisotropy
  * SH: dc = RGB2SH(U(0,1)) (gs_renderer.py:585, utils/sh_utils.py:122-123), rest N(0,0.05)
"""
from __future__ import annotations

import math

import numpy as np
import torch

SH_C0 = 0.28209479177387814


def ball_scene(P: int, radius: float = 0.5, sh_degree_max: int = 3, seed: int = 0,
               opacity: str = "sigmoid_normal", exact_knn: bool = True):
    """Returns dict of CPU fp32 tensors: means3D[P,3], scales[P,3] (post-exp), rotations[P,4]
    (unit), opacities[P,1] (post-sigmoid), shs[P,M,3]."""
    rng = np.random.RandomState(seed)
    phis = rng.random_sample(P) * 2 * np.pi
    costheta = rng.random_sample(P) * 2 - 1
    thetas = np.arccos(costheta)
    mu = rng.random_sample(P)
    r = radius * np.cbrt(mu)
    xyz = np.stack((r * np.sin(thetas) * np.cos(phis), r * np.sin(thetas) * np.sin(phis),
                    r * np.cos(thetas)), axis=1).astype(np.float32)
    if exact_knn and P >= 4:
        from scipy.spatial import cKDTree
        d, _ = cKDTree(xyz).query(xyz, k=4, workers=-1)
        dist2 = (d[:, 1:] ** 2).mean(1)
    else:
        # expected k-th NN distance in a uniform density n: r_k^3 ~ k / (4/3 pi n)
        n = P / (4.0 / 3.0 * math.pi * radius ** 3)
        rk2 = [(k / (4.0 / 3.0 * math.pi * n)) ** (2.0 / 3.0) for k in (1, 2, 3)]
        dist2 = np.full(P, float(np.mean(rk2)))
    base = np.sqrt(np.maximum(dist2, 1e-7))[:, None].repeat(3, 1)
    scales = (base * np.exp(rng.normal(0, 0.3, (P, 3)))).astype(np.float32)
    q = rng.normal(0, 1, (P, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    if opacity == "sigmoid_normal":
        op = 1.0 / (1.0 + np.exp(-rng.normal(0, 1.5, (P, 1))))
    elif opacity == "init":          # gs_renderer.py:598
        op = np.full((P, 1), 0.1)
    elif opacity == "uniform":
        op = rng.uniform(0.05, 0.95, (P, 1))
    else:
        raise ValueError(opacity)
    M = (sh_degree_max + 1) ** 2
    shs = np.zeros((P, M, 3), np.float32)
    shs[:, 0, :] = (rng.random_sample((P, 3)) - 0.5) / SH_C0
    if M > 1:
        shs[:, 1:, :] = rng.normal(0, 0.05, (P, M - 1, 3))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return dict(means3D=t(xyz), scales=t(scales), rotations=t(q), opacities=t(op), shs=t(shs))
