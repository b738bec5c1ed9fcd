#!/usr/bin/env python
"""Profiling companion of ``bench.py --workload train`` (BASELINE.json configs[4]): runs vspbfr_b200.train_step.TrainStep
eagerly and prints the kernels with the most device time of one iteration (torch profiler).

    python tools/train_step_bench.py [--size 512] [--batch 4] [--top 40]
The numbers for the record come from ``python bench.py --workload train`` (1 GPU: whole-iteration CUDA graph) and
``python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench.py --workload train --gpus N`` (DDP)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200 import _lib  # noqa: E402
from vspbfr_b200.train_step import TrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    _lib.load()
    ts = TrainStep(args.size, args.batch, "cuda")
    for _ in range(2):
        ts.step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        ts.step()
        torch.cuda.synchronize()
    rows = [(ev.key, ev.self_device_time_total, ev.count) for ev in prof.key_averages() if ev.self_device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    total = sum(r[1] for r in rows)
    print(f"device time of one eager iteration: {total / 1e3:.2f} ms over {sum(r[2] for r in rows)} kernels")
    for k, t, n in rows[:args.top]:
        print(f"{t / 1e3:8.3f} ms {100 * t / total:5.1f}% x{n:<4d} {k[:120]}")


if __name__ == "__main__":
    main()
