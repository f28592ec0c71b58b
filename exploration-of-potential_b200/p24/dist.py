"""Multi-GPU sharding of the loss path: the batch shards by image across the GPUs of one box, one
process per GPU; the only collective is the SUM all-reduce of the 28 loss sums (24 per-ray IoU sums, obj BCE, cls BCE,
num_fg, num_gt), enqueued on the compute stream between the sums kernel and the finalize kernel.  Assignments are
strictly per image, so no other data crosses GPUs; the postprocess needs no collective at all.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``total`` images for ``rank`` (earlier ranks take the remainder)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(outputs: torch.Tensor, labels: torch.Tensor, rank: int, world: int):
    lo, hi = shard_range(outputs.shape[0], rank, world)
    return outputs[lo:hi], labels[lo:hi]


def allreduce_sums(sums28: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the 28-float vector (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
    if sums28.numel() != 28:
        raise ValueError("expected the 28 loss sums")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums28, op=dist.ReduceOp.SUM, group=group)
    return sums28


def attach(loss_function, group=None):
    """Make ``Loss_Function.forward`` shard-aware: sums are all-reduced over ``group`` (default WORLD) before the
    normalisation, so every rank computes the same loss / weights / state."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised")
    loss_function.process_group = group if group is not None else dist.group.WORLD
    return loss_function
