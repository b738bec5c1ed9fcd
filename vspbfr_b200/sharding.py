"""Batch sharding of restoration inference across the GPUs of one box (SURVEY.md §8 e).

Images are independent (per-sample styles, no batch statistics in the hot-path layers), so the job is
split into contiguous slices, one process per GPU, weights replicated, and NO data-path collective:
the only communication is a barrier and a MAX-reduction of the device-side elapsed time for reporting.
"""
from __future__ import annotations

import os
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


def restore_from_host(net, decoder, low_h, codes_h, noise_z_h, out_h, micro: int = 32, device=None, restorer=None):
    """Restore a job whose inputs and outputs live in (pinned) HOST memory: micro-batches are copied in on one
    stream, processed on the current stream and copied out on a third, so the PCIe transfers of batch m+1 / m-1
    overlap the kernels of batch m (restoration_test.py:125-157 moves every batch synchronously).  Device staging
    buffers are two persistent sets guarded by events — nothing is allocated or freed across streams inside the loop.

    low_h [N,3,S,S], codes_h [N,18,512], noise_z_h [N,512] -> out_h [N,3,S,S] fp32, or — when ``out_h`` is a uint8
    [N,S,S,3] buffer — the images quantised on the device exactly as ``save_image(normalize=True, range=(-1, 1))`` would
    (imageio.quantize_u8: a quarter of the device->host bytes); returns when every device->host copy has completed.  ``restorer`` (e.g. a ``fastpath.GraphedRestorer`` captured for ``micro``) replaces the eager
    ``fastpath.restore_faces`` call for full micro-batches: ``restorer(low, codes, z) -> (restored, image)`` must return
    tensors it does not overwrite later."""
    import torch

    from . import fastpath

    device = torch.device(device if device is not None else torch.cuda.current_device())
    as_u8 = out_h.dtype == torch.uint8
    if as_u8:
        from .imageio import quantize_u8

    def result_rows(t):
        return quantize_u8(t) if as_u8 else t

    n = low_h.shape[0]
    compute = torch.cuda.current_stream(device)
    h2d, d2h = torch.cuda.Stream(device), torch.cuda.Stream(device)
    spans = list(micro_batches(0, n, micro))
    if not spans:
        return out_h
    stage = [tuple(torch.empty((micro,) + tuple(x.shape[1:]), dtype=x.dtype, device=device) for x in (low_h, codes_h, noise_z_h))
             for _ in range(2)]
    loaded = [torch.cuda.Event() for _ in range(2)]       # staging set i holds its batch
    loaded_small = [torch.cuda.Event() for _ in range(2)] # ... its codes and noise vectors (copied first: the decoder half needs only them)
    split = restorer is not None and hasattr(restorer, "graph_restore") and os.environ.get("VSP_NO_SPLIT_H2D") is None
    consumed = [torch.cuda.Event() for _ in range(2)]     # the kernels reading staging set i have been enqueued and finished
    copied = [None, None]                                  # (event, tensor) of the device->host copy two batches back
    h2d.wait_stream(compute)

    def stage_in(i):
        s, e = spans[i]
        k = i & 1
        with torch.cuda.stream(h2d):
            if i >= 2:
                h2d.wait_event(consumed[k])
            for dst, src in zip(stage[k][1:], (codes_h, noise_z_h)):
                dst[:e - s].copy_(src[s:e], non_blocking=True)
            loaded_small[k].record(h2d)
            stage[k][0][:e - s].copy_(low_h[s:e], non_blocking=True)
            loaded[k].record(h2d)

    stage_in(0)
    for i, (s, e) in enumerate(spans):
        k = i & 1
        if i + 1 < len(spans):
            stage_in(i + 1)                                # prefetch while this batch computes
        lo, co, zz = (t[:e - s] for t in stage[k])
        if split and e - s == micro and getattr(restorer, "tail_groups", 1) > 1:
            # grouped tail: the restorer's last level runs in sample groups; each group's rows go to the host while the
            # following groups still compute (the only exposed copy is the last group's)
            compute.wait_event(loaded_small[k])
            for c in copied:
                if c is not None:
                    compute.wait_event(c[0])               # the static output buffer is free again
            evs, keep = [], []

            def on_group(g, glo, ghi, restored, s=s, evs=evs, keep=keep):
                rows = result_rows(restored[glo:ghi])
                keep.append(rows)                          # alive until the caller's final synchronisation
                ready = torch.cuda.Event()
                ready.record(compute)
                d2h.wait_event(ready)
                with torch.cuda.stream(d2h):
                    out_h[s + glo:s + ghi].copy_(rows, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(d2h)
                evs.append(ev)

            restored, _ = restorer(lo, co, zz, clone=False, before_low=lambda k=k: compute.wait_event(loaded[k]),
                                   on_group=on_group)
            consumed[k].record(compute)
            copied[k] = (evs[-1], (restored, keep))        # keep the group tensors alive until their copies are done
            continue
        if split and e - s == micro:
            # the style-decoder half starts as soon as the codes are on the device; the image copy overlaps it
            compute.wait_event(loaded_small[k])
            restored, _ = restorer(lo, co, zz, before_low=lambda k=k: compute.wait_event(loaded[k]))
        else:
            compute.wait_event(loaded[k])
            if restorer is not None and e - s == micro:
                restored, _ = restorer(lo, co, zz)
            else:
                restored, _ = fastpath.restore_faces(net, decoder, lo, co, [zz])
        restored = result_rows(restored)
        consumed[k].record(compute)
        if copied[k] is not None:
            copied[k][0].synchronize()                     # long finished; lets the tensor two batches back be freed safely
        d2h.wait_event(consumed[k])
        with torch.cuda.stream(d2h):
            out_h[s:e].copy_(restored, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(d2h)
        copied[k] = (ev, restored)                         # keep `restored` alive until its copy has completed
    for c in copied:
        if c is not None:
            c[0].synchronize()
    return out_h
