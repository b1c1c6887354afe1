"""View-sharded data parallelism.

DreamScene renders the C_batch_size views of a step sequentially on one GPU and lets autograd
sum the per-view parameter gradients (/root/reference/training/scene_trainer.py:801-829,881).
Here each rank renders its own views with replicated Gaussian parameters; the rasterizer's
backward all-reduces (SUM) its flat parameter-gradient buffer over NCCL/NVLink before returning,
so every rank ends the step with the same summed gradients as the sequential loop.
Per-view quantities (means2D grad, radii, visibility) are NOT reduced, as in the reference,
which only uses the last view's (training/object_trainer.py:385-390).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

_group = None
_enabled = False


def enable_view_sharding(group: Optional["dist.ProcessGroup"] = None) -> None:
    """After this call every rasterizer backward all-reduces its parameter gradients."""
    global _group, _enabled
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _group, _enabled = group, True


def disable_view_sharding() -> None:
    global _group, _enabled
    _group, _enabled = None, False


def is_enabled() -> bool:
    return _enabled


def maybe_all_reduce(flat: torch.Tensor) -> None:
    if _enabled and dist.get_world_size(_group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_group)


def shard_views(num_views: int, rank: Optional[int] = None, world: Optional[int] = None):
    """Views {rank, rank+world, ...} of a batch (scene_trainer.py:801 loop index i)."""
    if rank is None:
        rank = dist.get_rank(_group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(_group) if dist.is_initialized() else 1
    return list(range(rank, num_views, world))
