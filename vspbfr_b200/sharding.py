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


def restore_from_host(net, decoder, low_h, codes_h, noise_z_h, out_h, micro: int = 32, device=None):
    """Restore a job whose inputs and outputs live in (pinned) HOST memory: micro-batches are copied in on one
    stream, processed on the current stream and copied out on a third, so the PCIe transfers of batch m+1 / m-1
    overlap the kernels of batch m (restoration_test.py:125-157 moves every batch synchronously).

    low_h [N,3,S,S], codes_h [N,18,512], noise_z_h [N,512] -> out_h [N,3,S,S] (filled asynchronously; the call
    returns after the last device->host copy has been enqueued AND completed)."""
    import torch

    from . import fastpath

    device = torch.device(device if device is not None else torch.cuda.current_device())
    n = low_h.shape[0]
    compute = torch.cuda.current_stream(device)
    h2d, d2h = torch.cuda.Stream(device), torch.cuda.Stream(device)
    spans = list(micro_batches(0, n, micro))

    def stage_in(span):
        s, e = span
        with torch.cuda.stream(h2d):
            t = tuple(x[s:e].to(device, non_blocking=True) for x in (low_h, codes_h, noise_z_h))
            ev = torch.cuda.Event()
            ev.record(h2d)
        return t, ev

    nxt = stage_in(spans[0]) if spans else None
    for i, (s, e) in enumerate(spans):
        (lo, co, zz), ready = nxt
        nxt = stage_in(spans[i + 1]) if i + 1 < len(spans) else None      # prefetch while this batch computes
        compute.wait_event(ready)
        for t in (lo, co, zz):
            t.record_stream(compute)
        restored, _ = fastpath.restore_faces(net, decoder, lo, co, [zz])
        done = torch.cuda.Event()
        done.record(compute)
        d2h.wait_event(done)
        restored.record_stream(d2h)
        with torch.cuda.stream(d2h):
            out_h[s:e].copy_(restored, non_blocking=True)
    compute.wait_stream(d2h)
    return out_h
