"""GPU tests of the training step (BASELINE configs[4]): the iteration runs, every parameter is updated, and DDP's
all-reduced gradients equal the average of the same shards' gradients computed in one process ("DDP grads ==
single process").  The two ranks of the DDP test share cuda:0 (the test box has one GPU) and talk over gloo; NCCL carries the
same all-reduce in ``bench.py --workload train --gpus N``."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vspbfr_b200.restorenet import Discriminator, Restoration_net
from vspbfr_b200.stylegan2 import Generator
from vspbfr_b200.train_step import TrainStep, d_logistic_loss, g_nonsaturating_loss

pytestmark = pytest.mark.gpu
DEV = "cuda"
SIZE = 32


def test_train_step_updates_every_parameter_and_replays_as_a_graph():
    ts = TrainStep(SIZE, 2, DEV, capturable=True)
    before = {k: v.detach().clone() for k, v in list(ts.g_module.named_parameters()) + list(ts.d_module.named_parameters())}
    d, r1, g = ts.step()
    assert all(torch.isfinite(v) for v in (d, r1, g))
    after = dict(list(ts.g_module.named_parameters()) + list(ts.d_module.named_parameters()))
    same = [k for k, v in before.items() if torch.equal(v, after[k])]
    assert not same, f"parameters not updated: {same[:5]}"
    replay = ts.capture()
    assert ts.graph_launches > 100
    snap = ts.g_module.conv1.fusion[0].weight.detach().clone()
    d2, r12, g2 = replay()
    torch.cuda.synchronize()
    assert all(torch.isfinite(v) for v in (d2, r12, g2))
    assert not torch.equal(snap, ts.g_module.conv1.fusion[0].weight)          # the replay trains


def test_flat_gradient_buffers_stay_attached_and_train():
    """reduce="flat": every .grad is (and stays, after backward passes and optimizer steps) a view into the network's one
    flat buffer — the thing the single NCCL all-reduce per backward operates on — and the graph replay trains."""
    ts = TrainStep(SIZE, 2, DEV, capturable=True, reduce="flat")

    def attached(module, flat):
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        off = 0
        for p in module.parameters():
            assert p.grad is not None and p.grad.data_ptr() == lo + off * 4 and p.grad.data_ptr() < hi
            off += p.numel()
        assert off == flat.numel()

    attached(ts.g_module, ts.flat_g)
    attached(ts.d_module, ts.flat_d)
    ts.step()
    attached(ts.g_module, ts.flat_g)
    attached(ts.d_module, ts.flat_d)
    assert float(ts.flat_g.abs().sum()) > 0 and float(ts.flat_d.abs().sum()) > 0
    replay = ts.capture()
    snap = ts.d_module.final_linear[1].weight.detach().clone()
    d, r1, g = replay()
    torch.cuda.synchronize()
    assert all(torch.isfinite(v) for v in (d, r1, g))
    assert not torch.equal(snap, ts.d_module.final_linear[1].weight)
    attached(ts.g_module, ts.flat_g)


def _nets(seed=7):
    torch.manual_seed(seed)
    # eval(): the only train/eval difference in these networks is Restoration_net's Dropout2d(0.5) on the global code
    # (models/RestoreNet.py:909), whose random mask would differ between the processes being compared
    g = Restoration_net(SIZE, 512, 2, channel_multiplier=2).to(DEV).eval()
    d = Discriminator(SIZE, channel_multiplier=2).to(DEV)
    dec = Generator(SIZE * 2, 512, 2, channel_multiplier=2).to(DEV).eval()
    return g, d, dec


def _batch(n):
    gen = torch.Generator().manual_seed(11)
    return (torch.rand(n, 3, SIZE, SIZE, generator=gen) * 2 - 1, torch.rand(n, 3, SIZE, SIZE, generator=gen) * 2 - 1,
            torch.randn(n, 18, 512, generator=gen), torch.randn(n, 512, generator=gen))


def _losses_backward(g, d, dec, real, low, codes, z):
    """G loss through D, and D's logistic loss: the two gradient sets DDP all-reduces."""
    real, low, codes, z = (t.to(DEV) for t in (real, low, codes, z))
    with torch.no_grad():
        _, feats = dec([codes], input_is_latent=True, return_features=True)
    restored = g(low, feats, codes, [z])
    g_loss = g_nonsaturating_loss(d(restored))
    d_loss = d_logistic_loss(d(real), d(restored.detach()))
    g.zero_grad()
    d.zero_grad()
    # separate backward passes so each network's gradient comes from its own loss only
    gg = torch.autograd.grad(g_loss, [p for p in g.parameters()], retain_graph=True)
    d_loss.backward()
    return gg


def _ddp_worker(rank, world, port, per_rank, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, d, dec = _nets()
        gd = torch.nn.parallel.DistributedDataParallel(g, broadcast_buffers=False)
        dd = torch.nn.parallel.DistributedDataParallel(d, broadcast_buffers=False)
        real, low, codes, z = _batch(world * per_rank)
        # rank r owns rows r, r + world, ...: the single-process reference's minibatch-stddev groups (strided by batch/group,
        # models/RestoreNet.py:1250-1258) then coincide with the per-rank batches
        rows = slice(rank, None, world)
        real, low, codes, z = (t[rows].to(DEV) for t in (real, low, codes, z))
        with torch.no_grad():
            _, feats = dec([codes], input_is_latent=True, return_features=True)
        restored = gd(low, feats, codes, [z])
        for p in d.parameters():
            p.requires_grad_(False)
        g_loss = g_nonsaturating_loss(d(restored))
        gd.zero_grad()
        g_loss.backward()                                    # DDP all-reduces (averages) G's gradients here
        for p in d.parameters():
            p.requires_grad_(True)
        d_loss = d_logistic_loss(dd(real), dd(restored.detach()))
        dd.zero_grad()
        d_loss.backward()
        torch.cuda.synchronize()
        if rank == 0:
            # numpy, not tensors: a tensor in a Queue travels as a shared-memory handle that dies with this process
            q.put(({k: p.grad.cpu().numpy() for k, p in g.named_parameters()},
                   {k: p.grad.cpu().numpy() for k, p in d.named_parameters()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_ddp_gradients_equal_single_process_gradients():
    world, per_rank = 2, 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    g_ddp, d_ddp = q.get(timeout=900)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single process: the same two shards one after the other, gradients averaged (what DDP's all-reduce computes).  The
    # shards are NOT fused into one batch of 8: the style MLP / modulation linears are library GEMMs whose fp32 summation
    # order depends on the batch size (measured 1e-6), and this random-init network amplifies such differences through
    # bf16 rounding flips to ~10 % of a gradient — a property of the network, not of the all-reduce under test.
    # What remains between the two sides is the order of fp32 atomics (split-K weight gradients, bias reductions) followed by bf16 rounding of the next gradient: measured 2e-3, bound 1e-2.
    g, d, dec = _nets()
    batch = _batch(world * per_rank)
    g_sum, d_sum = None, None
    for r in range(world):
        shard = [t[r::world] for t in batch]
        gg = _losses_backward(g, d, dec, *shard)
        dgr = [p.grad.detach().clone() for p in d.parameters()]
        g_sum = list(gg) if g_sum is None else [a + b for a, b in zip(g_sum, gg)]
        d_sum = dgr if d_sum is None else [a + b for a, b in zip(d_sum, dgr)]
    for (k, _), want in zip(g.named_parameters(), g_sum):
        if k.endswith("noise.weight"):
            continue        # its gradient is sum(dy * noise) with noise drawn per call from each process's own generator
        got, want = torch.from_numpy(g_ddp[k]).to(DEV), want / world
        scale = float(want.abs().max()) + 1e-12
        assert float((got - want).abs().max()) <= 1e-2 * scale, (k, float((got - want).abs().max()), scale)
    for (k, _), want in zip(d.named_parameters(), d_sum):
        got, want = torch.from_numpy(d_ddp[k]).to(DEV), want / world
        scale = float(want.abs().max()) + 1e-12
        assert float((got - want).abs().max()) <= 1e-2 * scale, (k, float((got - want).abs().max()), scale)


def test_generator_step_includes_the_reference_loss_networks():
    """restoration_train.py:236-245: with the default weights (0.5 / 0.1) the generator loss carries the LPIPS-VGG and the
    ArcFace-ResNet identity terms (lossnets.py, frozen, eval mode); with both weights 0 the step is the GAN terms alone."""
    full = TrainStep(SIZE, 2, DEV, seed=3)
    bare = TrainStep(SIZE, 2, DEV, seed=3, percept_loss_weight=0.0, id_loss_weight=0.0)
    assert full.percept_loss is not None and full.id_loss is not None and bare.percept_loss is None and bare.id_loss is None
    assert not full.percept_loss.model.training and not full.id_loss.Z.training
    assert not any(p.requires_grad for p in list(full.percept_loss.parameters()) + list(full.id_loss.parameters()))
    torch.manual_seed(0)
    _, _, g_full = full.step()
    torch.manual_seed(0)
    _, _, g_bare = bare.step()
    assert torch.isfinite(g_full) and torch.isfinite(g_bare) and float((g_full - g_bare).abs()) > 0
