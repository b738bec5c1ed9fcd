"""CPU oracle for the Discriminator and the R1 penalty (test infrastructure).

Functional torch-CPU restatement, driven by a plain ``state_dict``, of
/root/reference/models/RestoreNet.py:1137-1265 (``ConvLayer`` / ``ResBlock`` / ``Discriminator.forward`` with its
minibatch-stddev block :1250-1258) and of the R1 step of /root/reference/restoration_train.py:66-73, :200-216.  Built
from ``upfirdn2d_native_port``, the leaky-ReLU bias-act and plain ``F.conv2d`` (what ``conv2d_gradfix.conv2d`` reduces to
on the CPU, op/conv2d_gradfix.py:34-42); differentiable to any order.  Pinned against tests/golden/gradfix.npz (outputs
of the real reference) by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .upfirdn2d_ref import upfirdn2d_native_port

SQRT2 = math.sqrt(2.0)


class _RoundBf16(torch.autograd.Function):
    """Round to bf16 in the forward AND round the cotangent in the backward — what a bf16-operand GEMM does to both passes."""

    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().to(g.dtype)


_emulate_bf16 = False


def _conv2d(x, w, bias=None, **kw):
    if _emulate_bf16:
        x, w = _RoundBf16.apply(x), _RoundBf16.apply(w)
    return F.conv2d(x, w, bias, **kw)


class bf16_operands:
    """``with bf16_operands():`` the oracle's convolutions take bf16-rounded operands (fp32 accumulation) — the precision
    floor of ANY bf16 tensor-core implementation of the same network, used by the GPU tests to size their tolerances."""

    def __enter__(self):
        global _emulate_bf16
        self._saved, _emulate_bf16 = _emulate_bf16, True

    def __exit__(self, *exc):
        global _emulate_bf16
        _emulate_bf16 = self._saved


def _blur(x, pad):
    k = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=x.dtype)
    k = torch.outer(k, k)
    return upfirdn2d_native_port(x, k / k.sum(), 1, 1, pad)


def _lrelu(x, bias=None):
    if bias is not None:
        x = x + bias.reshape(1, -1, *([1] * (x.ndim - 2)))
    return F.leaky_relu(x, 0.2) * SQRT2


def _conv_layer(sd, p, x, k, downsample=False, activate=True):
    """ConvLayer (models/RestoreNet.py:1137-1179): [Blur] + EqualConv2d + [FusedLeakyReLU]; module indices shift by one when
    the blur is present."""
    i = 0
    if downsample:
        q = 2 + (k - 1)                                         # (len(blur_kernel) - factor) + (k - 1)
        x = _blur(x, ((q + 1) // 2, q // 2))
        i = 1
    w = sd[f"{p}{i}.weight"]
    scale = 1.0 / math.sqrt(w.shape[1] * k * k)
    bias = sd.get(f"{p}{i}.bias")
    x = _conv2d(x, w * scale, bias, stride=2 if downsample else 1, padding=0 if downsample else k // 2)
    if activate:
        x = _lrelu(x, sd.get(f"{p}{i + 1}.bias"))
    return x


def _equal_linear(sd, p, x, act=False):
    w = sd[p + "weight"]
    scale = 1.0 / math.sqrt(w.shape[1])
    if act:
        return _lrelu(F.linear(x, w * scale), sd[p + "bias"])
    return F.linear(x, w * scale, sd[p + "bias"])


def discriminator_ref(sd, img, stddev_group=4):
    """Discriminator.forward (models/RestoreNet.py:1244-1265)."""
    n_blocks = len({k.split(".")[1] for k in sd if k.startswith("encoder_convs.")})
    out = _conv_layer(sd, "encoder_input_convs.", img, 1)
    for b in range(n_blocks):
        p = f"encoder_convs.{b}."
        y = _conv_layer(sd, p + "conv1.", out, 3)
        y = _conv_layer(sd, p + "conv2.", y, 3, downsample=True)
        skip = _conv_layer(sd, p + "skip.", out, 1, downsample=True, activate=False)
        out = (y + skip) / SQRT2
    batch, channel, height, width = out.shape
    group = min(batch, stddev_group)
    sdv = out.view(group, -1, 1, channel, height, width)
    sdv = torch.sqrt(sdv.var(0, unbiased=False) + 1e-8)
    sdv = sdv.mean([2, 3, 4], keepdim=True).squeeze(2).repeat(group, 1, height, width)
    out = _conv_layer(sd, "final_conv.", torch.cat([out, sdv], 1), 3)
    out = _equal_linear(sd, "final_linear.0.", out.reshape(batch, -1), act=True)
    return _equal_linear(sd, "final_linear.1.", out)


def r1_step_ref(sd, real, r1_weight=10.0, d_reg_every=16):
    """One R1 regularisation step (restoration_train.py:66-73, :200-216): returns (pred, grad_real, r1, {param: grad})."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "kernel" not in k}
    full = dict(sd)
    full.update(params)
    real = real.detach().clone().requires_grad_(True)
    pred = discriminator_ref(full, real)
    (grad_real,) = torch.autograd.grad(pred.sum(), real, create_graph=True)
    r1 = grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()
    keys = list(params)
    grads = torch.autograd.grad(r1_weight / 2 * r1 * d_reg_every + 0 * pred[0].sum(), [params[k] for k in keys], allow_unused=True)
    return pred.detach(), grad_real.detach(), r1.detach(), {k: g for k, g in zip(keys, grads)}
