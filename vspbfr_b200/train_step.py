"""One training iteration of the restoration GAN on the rebuilt layers (BASELINE.json configs[4]).

Mirror of the loop body of /root/reference/restoration_train.py:159-256 — discriminator logistic step (:181-193), R1
regulariser with its double backward through the Discriminator (:63-73, :200-216), generator non-saturating step
(:220-250) and the EMA update (:46-51, :255) — with the loss helpers under the reference's own names
(``d_logistic_loss`` :56-60, ``d_r1_loss`` :63-73, ``g_nonsaturating_loss`` :76-79, ``requires_grad`` :41-43,
``accumulate`` :46-51).  The w+ codes stand in for the frozen e4e encoder + code diffuser (:166-167, outside the hot
path); the LPIPS-VGG and ArcFace terms of the generator step (:236-245, default weights 0.5 / 0.1) run on ``lossnets.py`` —
the reference's network structures, randomly initialised because the pretrained checkpoints cannot be downloaded here.  Data parallelism is the reference's: one process per GPU,
``DistributedDataParallel`` around generator and discriminator (:431-445), gradients all-reduced by NCCL during backward.

Every convolution of the step — forward, input gradient, weight gradient and the second-order terms of R1 — runs on the
tcgen05 kernels through ``op.conv2d_gradfix`` / ``op.modconv``; there is no library convolution on this path.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F
from torch import autograd

from .op import conv2d_gradfix
from .restorenet import Discriminator, Restoration_net, mixing_noise
from .stylegan2 import Generator


def requires_grad(model, flag=True):
    for p in model.parameters():
        p.requires_grad = flag


def accumulate(model1, model2, decay=0.999):
    """EMA of ``model2``'s parameters into ``model1`` (one fused multi-tensor update instead of 2 launches per tensor)."""
    p1, p2 = dict(model1.named_parameters()), dict(model2.named_parameters())
    dst = [p1[k].data for k in p1]
    src = [p2[k].data for k in p1]
    torch._foreach_mul_(dst, decay)
    torch._foreach_add_(dst, src, alpha=1 - decay)


def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    with conv2d_gradfix.no_weight_gradients():
        grad_real, = autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_nonsaturating_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


class TrainStep:
    """Networks, optimizers and synthetic batch of one rank; ``step()`` = one full iteration.

    Gradient exchange across ``world`` ranks (the default process group must already be initialised):
      * ``reduce="ddp"``  — the reference's way: generator and discriminator wrapped in DistributedDataParallel
        (restoration_train.py:431-445), bucketed all-reduce overlapped with backward by the reducer's hooks; eager only.
      * ``reduce="flat"`` — every parameter's ``.grad`` is a view into one flat fp32 buffer per network, averaged with ONE
        NCCL all-reduce per backward (D logistic, D R1, G: 116 + 116 + 450 MB over NVLink) between the phases of the
        iteration.  The phases hold no collective, so each replays as a CUDA graph (``capture()``): the eager iteration is
        bound by the host (~170 ms of Python / autograd for ~107 ms of GPU work), and a few ms of un-overlapped all-reduce
        cost far less than that.
    ``capturable`` keeps the Adam step counters on the device (needed for capture)."""

    def __init__(self, size=512, batch=4, device="cuda", world=1, local_rank=0, rank=0, r1=10.0, d_reg_every=16,
                 style_dim=512, n_mlp=8, mixing=0.9, capturable=False, seed=0, reduce="ddp", percept_loss_weight=0.5,
                 id_loss_weight=0.1, lpips_weights=None, arcface_path=None):
        if reduce not in ("ddp", "flat"):
            raise ValueError(f"reduce must be 'ddp' or 'flat', got {reduce!r}")
        self.size, self.batch, self.world, self.device = size, batch, world, torch.device(device)
        self.r1, self.d_reg_every, self.style_dim, self.mixing = r1, d_reg_every, style_dim, mixing
        self.reduce = reduce
        dev = self.device
        torch.manual_seed(seed)
        self.g_module = Restoration_net(size, style_dim, n_mlp, channel_multiplier=2).to(dev)
        self.g_ema = Restoration_net(size, style_dim, n_mlp, channel_multiplier=2).to(dev).eval()
        self.g_ema.load_state_dict(self.g_module.state_dict())
        self.d_module = Discriminator(size, channel_multiplier=2).to(dev)
        self.decoder = Generator(max(size * 2, 16), style_dim, n_mlp, channel_multiplier=2).to(dev).eval()
        self.generator, self.discriminator = self.g_module, self.d_module
        self.flat_g = self.flat_d = None
        if world > 1 and reduce == "ddp":
            ddp = torch.nn.parallel.DistributedDataParallel
            ids = [local_rank] if dev.type == "cuda" else None
            self.generator = ddp(self.g_module, device_ids=ids, broadcast_buffers=False)
            self.discriminator = ddp(self.d_module, device_ids=ids, broadcast_buffers=False)
        if reduce == "flat":
            self.flat_g, self.flat_d = _flatten_grads(self.g_module), _flatten_grads(self.d_module)
        g_ratio, d_ratio = 4 / 5, d_reg_every / (d_reg_every + 1)          # restoration_train.py:410-422
        self.g_optim = torch.optim.Adam(self.generator.parameters(), lr=0.002 * g_ratio, betas=(0.0, 0.99 ** g_ratio),
                                        capturable=capturable)
        self.d_optim = torch.optim.Adam(self.discriminator.parameters(), lr=0.002 * d_ratio, betas=(0.0, 0.99 ** d_ratio),
                                        capturable=capturable)
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        self.real_img = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).to(dev)
        self.low_img = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).to(dev)
        self.codes = torch.randn(batch, 18, style_dim, generator=g).to(dev)
        self._de_feats = None
        self._losses = [None, None, None]
        # loss networks of the generator step (restoration_train.py:116,143,236-245; the reference's default weights 0.5 / 0.1):
        # frozen LPIPS-VGG and ArcFace ResNet-101, randomly initialised unless checkpoints are given (lossnets.py)
        self.percept_loss_weight, self.id_loss_weight = float(percept_loss_weight), float(id_loss_weight)
        self.percept_loss = self.id_loss = None
        if self.percept_loss_weight > 0 or self.id_loss_weight > 0:
            from . import lossnets
            torch.manual_seed(seed + 1)
            cl = torch.channels_last if dev.type == "cuda" else torch.contiguous_format     # NHWC: the library's tensor-core path
            if self.percept_loss_weight > 0:
                self.percept_loss = lossnets.PerceptualLoss(lin_weights_path=lpips_weights).to(dev).to(memory_format=cl)
            if self.id_loss_weight > 0:
                self.id_loss = lossnets.IDLoss(arcface_path).to(dev).to(memory_format=cl)
            self._loss_format = cl

    def grad_bytes(self):
        """fp32 gradient bytes all-reduced per iteration: D twice (logistic + R1 steps), G once."""
        n = lambda m: sum(p.numel() for p in m.parameters()) * 4
        return {"generator": n(self.g_module), "discriminator": n(self.d_module), "per_step": n(self.g_module) + 2 * n(self.d_module)}

    # ---- gradient bookkeeping -------------------------------------------------------------------------------------
    def _zero(self, module, flat):
        if flat is not None:
            flat.zero_()                               # .grad tensors are views of `flat`: one memset
        else:
            module.zero_grad()

    def _exchange(self, flat, sync):
        """Average the gradients of the backward that just ran (flat mode; DDP's reducer has already done it)."""
        if flat is not None and self.world > 1 and sync:
            import torch.distributed as dist
            if dist.get_backend() == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            else:                                       # gloo (CPU-side tests) has no AVG
                dist.all_reduce(flat)
                flat.div_(self.world)

    def _nosync(self, module, sync):
        ddp = self.world > 1 and self.reduce == "ddp"
        return module.no_sync() if (ddp and not sync) else contextlib.nullcontext()

    # ---- the four phases of restoration_train.py:159-256; a gradient exchange sits between consecutive phases ---------
    def _phase_d(self, sync=True):
        """Style-decoder features, generator forward (no grad), discriminator logistic loss and its backward."""
        dev = self.device
        with torch.no_grad():
            _, self._de_feats = self.decoder([self.codes], input_is_latent=True, return_features=True)
        requires_grad(self.generator, False)
        requires_grad(self.discriminator, True)
        noise = mixing_noise(self.batch, self.style_dim, self.mixing, dev)
        with torch.no_grad():
            restored = self.generator(self.low_img, self._de_feats, self.codes, noise)
        with self._nosync(self.discriminator, sync):
            d_loss = d_logistic_loss(self.discriminator(self.real_img), self.discriminator(restored.detach()))
            self._zero(self.discriminator, self.flat_d)
            d_loss.backward()
        self._losses[0] = d_loss.detach()

    def _phase_r1(self, sync=True):
        """Discriminator update, then the R1 penalty and its (double) backward — every iteration here: that path is what
        configs[4] is about; the reference runs it every d_reg_every-th iteration with the same weight."""
        self.d_optim.step()
        tmp = self.real_img.detach().clone().requires_grad_(True)
        with self._nosync(self.discriminator, sync):
            real_pred = self.discriminator(tmp)
            r1_loss = d_r1_loss(real_pred, tmp)
            self._zero(self.discriminator, self.flat_d)
            (self.r1 / 2 * r1_loss * self.d_reg_every + 0 * real_pred[0]).backward()
        self._losses[1] = r1_loss.detach()

    def _phase_g(self, sync=True):
        """Discriminator update, then the generator's non-saturating loss and its backward."""
        self.d_optim.step()
        requires_grad(self.generator, True)
        requires_grad(self.discriminator, False)
        noise = mixing_noise(self.batch, self.style_dim, self.mixing, self.device)
        with self._nosync(self.generator, sync):
            restored = self.generator(self.low_img, self._de_feats, self.codes, noise)
            g_loss = g_nonsaturating_loss(self.discriminator(restored))
            if self.percept_loss is not None or self.id_loss is not None:
                pred = restored.contiguous(memory_format=self._loss_format)
                real = self.real_img.detach().contiguous(memory_format=self._loss_format)
            if self.percept_loss is not None:          # restoration_train.py:236-239
                g_loss = g_loss + self.percept_loss(pred, real).sum() * self.percept_loss_weight
            if self.id_loss is not None:               # :242-245
                g_loss = g_loss + self.id_loss(pred, real) * self.id_loss_weight
            self._zero(self.generator, self.flat_g)
            g_loss.backward()
        self._losses[2] = g_loss.detach()

    def _phase_update(self):
        """Generator update and EMA."""
        self.g_optim.step()
        accumulate(self.g_ema, self.g_module, 0.5 ** (32 / (10 * 1000)))

    def step(self, sync=True):
        """One eager iteration; ``sync=False`` skips the gradient exchange (to measure its exposed cost)."""
        self._phase_d(sync)
        self._exchange(self.flat_d, sync)
        self._phase_r1(sync)
        self._exchange(self.flat_d, sync)
        self._phase_g(sync)
        self._exchange(self.flat_g, sync)
        self._phase_update()
        return tuple(self._losses)

    def capture(self, warmup=3):
        """The iteration as CUDA graphs: ONE graph on a single GPU; with ``reduce="flat"`` on several GPUs one graph per
        phase (shared memory pool) with the NCCL all-reduces issued between the replays.  Returns
        ``replay(sync=True) -> (d, r1, g)`` losses (static tensors)."""
        from . import _lib
        if self.world > 1 and self.reduce != "flat":
            raise RuntimeError("graph capture on several GPUs needs reduce='flat' (DDP's reducer is not captured)")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(3, warmup)):
                self.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.flat_g is None:
            self.generator.zero_grad(set_to_none=True)
            self.discriminator.zero_grad(set_to_none=True)
        n0 = _lib.launch_count()
        if self.world == 1:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.step()
            graphs = [graph]
        else:
            graphs, pool = [], None
            for phase in (self._phase_d, self._phase_r1, self._phase_g, self._phase_update):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    phase()
                pool = pool or g.pool()
                graphs.append(g)
        self._graphs = graphs
        self.graph_launches = _lib.launch_count() - n0          # this library's kernels inside one replayed iteration
        out = tuple(self._losses)

        def replay(sync=True):
            if len(graphs) == 1:
                graphs[0].replay()
                return out
            graphs[0].replay()
            self._exchange(self.flat_d, sync)
            graphs[1].replay()
            self._exchange(self.flat_d, sync)
            graphs[2].replay()
            self._exchange(self.flat_g, sync)
            graphs[3].replay()
            return out
        return replay


def _flatten_grads(module):
    """Give every parameter of ``module`` a ``.grad`` that is a view into one flat fp32 buffer (returned): autograd
    accumulates into existing gradients in place, so one all-reduce / one memset covers the whole network."""
    params = [p for p in module.parameters()]
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    return flat
