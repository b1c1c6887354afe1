"""The PyTorch glue of DreamScene's scene_render restated (harness code: what dreamscene_b200.scene
replaces; used by tests and benchmarks/scene_step.py, never by the product)."""
from typing import Sequence

import torch


def _get(group, name):
    return group[name] if isinstance(group, dict) else getattr(group, name)


def reference_assemble(groups: Sequence, z_shs=None, z_scales=None):
    """The reference expressions in plain PyTorch (scene_gaussian.py:753-857; used by tests and the
    cfg5 benchmark as the thing being replaced).  z_* = standard-normal draws or None (no augmentation)."""
    cat = lambda name: torch.cat([_get(g, name) for g in groups])
    means3D = cat("_xyz")
    opacity = torch.sigmoid(cat("_opacity"))
    scales = torch.exp(cat("_scaling"))
    rotations = torch.nn.functional.normalize(cat("_rotation"))
    shs = torch.cat([torch.cat((_get(g, "_features_dc"), _get(g, "_features_rest")), dim=1) for g in groups])
    if z_shs is not None:
        shs = shs + z_shs * ((0.2 ** 0.5) * shs)
    if z_scales is not None:
        scales = torch.clamp(scales + z_scales * ((0.2 ** 0.5) * scales / 4), 0.0)
    return means3D, opacity, scales, rotations, shs
