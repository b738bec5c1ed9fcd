"""Batch sharding of restoration inference across the GPUs of one box (SURVEY.md §8 e).

Images are independent (per-sample styles, no batch statistics in the hot-path layers), so the job is
split into contiguous slices, one process per GPU, weights replicated, and NO data-path collective:
the only communication is a barrier and a MAX-reduction of the device-side elapsed time for reporting.
"""
from __future__ import annotations

from typing import Iterator, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``total`` items for ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def micro_batches(lo: int, hi: int, micro: int) -> Iterator[Tuple[int, int]]:
    """[lo, hi) in steps of ``micro`` (last one ragged)."""
    if micro <= 0:
        raise ValueError("micro batch must be positive")
    for s in range(lo, hi, micro):
        yield s, min(s + micro, hi)


def max_over_ranks(seconds: float, device=None) -> float:
    """MAX-reduce a per-rank elapsed time (works with nccl on GPU and gloo on CPU)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device or "cpu")
    if t.device.type == "cuda":
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def restore_sharded(net, decoder, low_imgs, codes, noise_z, rank: int, world: int, micro: int = 8, out=None):
    """Run this rank's slice of a job through the fused hot path; returns (lo, hi, restored slice)."""
    import torch

    from . import fastpath

    lo, hi = shard_range(low_imgs.shape[0], rank, world)
    if out is None:
        out = torch.empty((hi - lo,) + tuple(low_imgs.shape[1:]), dtype=torch.float32, device=low_imgs.device)
    for s, e in micro_batches(lo, hi, micro):
        restored, _ = fastpath.restore_faces(net, decoder, low_imgs[s:e], codes[s:e], [noise_z[s:e]])
        out[s - lo:e - lo] = restored
    return lo, hi, out
