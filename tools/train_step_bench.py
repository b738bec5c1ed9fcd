#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[4]): one iteration of restoration_train.py:153-256 on the rebuilt
layers — D logistic step, R1 regulariser (double backward through the Discriminator, restoration_train.py:200-216),
generator non-saturating step, EMA accumulate — 512x512, batch 4 per GPU, random-init, synthetic data.
LPIPS / ArcFace losses need pretrained nets (no network): percept/id weights are 0, as SURVEY.md §8(d) prescribes.

    python tools/train_step_bench.py [--steps 3] [--batch 4] [--size 512]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py   (DDP, NCCL)

Prints one JSON line on rank 0: images/s over all ranks, ms per step (max over ranks, CUDA events), and the share of
the step spent in gradient all-reduce as seen by a second run with the all-reduce disabled (`no_sync`)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import autograd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200 import _lib  # noqa: E402
from vspbfr_b200.op import conv2d_gradfix  # noqa: E402
from vspbfr_b200.restorenet import Discriminator, Restoration_net, mixing_noise  # noqa: E402
from vspbfr_b200.stylegan2 import Generator  # noqa: E402


def d_logistic_loss(real_pred, fake_pred):          # restoration_train.py:56-60
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def d_r1_loss(real_pred, real_img):                  # restoration_train.py:63-73
    with conv2d_gradfix.no_weight_gradients():
        grad_real, = autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_nonsaturating_loss(fake_pred):                 # restoration_train.py:76-79
    return F.softplus(-fake_pred).mean()


def requires_grad(model, flag=True):
    for p in model.parameters():
        p.requires_grad = flag


def accumulate(model1, model2, decay=0.999):         # restoration_train.py:33-38
    par1, par2 = dict(model1.named_parameters()), dict(model2.named_parameters())
    for k in par1.keys():
        par1[k].data.mul_(decay).add_(par2[k].data, alpha=1 - decay)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--r1", type=float, default=10.0)
    ap.add_argument("--d_reg_every", type=int, default=16)
    ap.add_argument("--profile", type=int, default=0, help="N > 0: print the N kernels with the most device time of one eager iteration")
    ap.add_argument("--graph", type=int, default=0,
                    help="1 (single GPU only): capture the whole iteration (D, R1, G, both Adam steps, EMA) as ONE CUDA graph")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    torch.manual_seed(0)
    dec_size = max(args.size * 2, 16)
    generator = Restoration_net(args.size, 512, 8, channel_multiplier=2).to(dev)
    g_ema = Restoration_net(args.size, 512, 8, channel_multiplier=2).to(dev).eval()
    g_ema.load_state_dict(generator.state_dict())
    discriminator = Discriminator(args.size, channel_multiplier=2).to(dev)
    decoder = Generator(dec_size, 512, 8, channel_multiplier=2).to(dev).eval()
    g_module, d_module = generator, discriminator
    if world > 1:
        generator = torch.nn.parallel.DistributedDataParallel(generator, device_ids=[local], broadcast_buffers=False)
        discriminator = torch.nn.parallel.DistributedDataParallel(discriminator, device_ids=[local], broadcast_buffers=False)
    g_reg_ratio, d_reg_ratio = 4 / 5, args.d_reg_every / (args.d_reg_every + 1)
    use_graph = bool(args.graph) and world == 1
    g_optim = torch.optim.Adam(generator.parameters(), lr=0.002 * g_reg_ratio, betas=(0.0, 0.99 ** g_reg_ratio),
                               capturable=use_graph)
    d_optim = torch.optim.Adam(discriminator.parameters(), lr=0.002 * d_reg_ratio, betas=(0.0, 0.99 ** d_reg_ratio),
                               capturable=use_graph)

    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    real_img = (torch.rand(args.batch, 3, args.size, args.size, generator=g) * 2 - 1).to(dev)
    low_img = (torch.rand(args.batch, 3, args.size, args.size, generator=g) * 2 - 1).to(dev)
    codes = torch.randn(args.batch, 18, 512, generator=g).to(dev)          # stands in for e4e + code diffuser

    def step(sync=True):
        import contextlib
        def nosync(m):      # a fresh context per use (generator-based context managers are single-shot)
            return m.no_sync() if (world > 1 and not sync) else contextlib.nullcontext()
        with torch.no_grad():
            _, de_feats = decoder([codes], input_is_latent=True, return_features=True)
        # ---- D
        requires_grad(generator, False)
        requires_grad(discriminator, True)
        noise = mixing_noise(args.batch, 512, 0.9, dev)
        with torch.no_grad():
            restored = generator(low_img, de_feats, codes, noise)
        with nosync(discriminator):
            fake_pred = discriminator(restored.detach())
            real_pred = discriminator(real_img)
            d_loss = d_logistic_loss(real_pred, fake_pred)
            discriminator.zero_grad()
            d_loss.backward()
        d_optim.step()
        # ---- R1 (forced every step here: the double-backward path is what this benchmark is about)
        tmp = real_img.detach().clone().requires_grad_(True)
        with nosync(discriminator):
            real_pred = discriminator(tmp)
            r1 = d_r1_loss(real_pred, tmp)
            discriminator.zero_grad()
            (args.r1 / 2 * r1 * args.d_reg_every + 0 * real_pred[0]).backward()
        d_optim.step()
        # ---- G
        requires_grad(generator, True)
        requires_grad(discriminator, False)
        noise = mixing_noise(args.batch, 512, 0.9, dev)
        with nosync(generator):
            restored = generator(low_img, de_feats, codes, noise)
            g_loss = g_nonsaturating_loss(discriminator(restored))
            generator.zero_grad()
            g_loss.backward()
        g_optim.step()
        accumulate(g_ema, g_module, 0.5 ** (32 / (10 * 1000)))
        return d_loss.detach(), r1.detach(), g_loss.detach()

    def timed(n, sync=True):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            out = step(sync)
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    eager_step = step
    if use_graph:
        # whole-iteration capture: eager warm-up on a side stream (optimizer state, cuDNN plans), then ONE graph whose replay
        # is the complete iteration — the eager loop is bound by the host (same step time at batch 4 and 8)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(3, args.warmup)):
                eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        generator.zero_grad(set_to_none=True)
        discriminator.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            static_out = eager_step()
        per_graph = _lib.launch_count() - n0

        def step(sync=True):
            graph.replay()
            return static_out
    for _ in range(args.warmup):
        step()
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
            eager_step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=args.profile, max_name_column_width=90),
              file=sys.stderr)
        print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=25,
                                                                 max_name_column_width=40, max_shapes_column_width=90),
              file=sys.stderr)
    l0 = _lib.launch_count()
    t, (d_l, r1_l, g_l) = timed(args.steps)
    launches = _lib.launch_count() - l0
    if use_graph:
        launches = per_graph * args.steps
    t_nosync = timed(args.steps, sync=False)[0] if world > 1 else t
    if rank == 0:
        print(json.dumps({
            "metric": "train_images_per_sec_512", "value": world * args.batch * args.steps / t, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "ms_per_step": 1e3 * t / args.steps,
            "ms_per_step_no_allreduce": 1e3 * t_nosync / args.steps,
            "allreduce_exposed_frac": max(0.0, 1 - t_nosync / t),
            "config": {"workload": "restoration_train.py step (D + R1 double backward + G + EMA), percept/id weights 0",
                       "size": args.size, "batch_per_gpu": args.batch, "parallelism": f"DDP x{world} (NCCL)",
                       "launch": "one CUDA graph replay per iteration" if use_graph else "eager"},
            "losses": {"d": float(d_l), "r1": float(r1_l), "g": float(g_l)},
            "gpu_launches": int(launches), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
