"""fused_leaky_relu / FusedLeakyReLU — host side of the sm_100a bias + activation op.

Mirrors /root/reference/op/fused_act.py:126-233: same names and signatures, the forward
saves the OUTPUT (not the input) and the backward gates on its sign; the second-order
backward is the same kernel applied to (gradgrad_input + gradgrad_bias).  Differences by
design: the bias-gradient reduction is fused into the backward kernel (the reference runs
a separate ``sum``, :139-145), and there is no CPU branch (so the reference CPU branch's
quirk of ignoring ``negative_slope``, :222, does not exist here).
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function

from .. import _lib


def _geometry(x):
    """(n, step_b, size_b): bias broadcasts on dim 1 (op/fused_bias_act_kernel.cu:84-88)."""
    step_b = 1
    for d in x.shape[2:]:
        step_b *= d
    return x.numel(), step_b, (x.shape[1] if x.ndim > 1 else 1)


def bias_act_raw(x, bias, ref, act, grad, alpha, scale):
    """``fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)`` of the reference."""
    if x.device.type != "cuda":
        raise RuntimeError("input must be a CUDA tensor (vspbfr_b200 has no CPU path)")
    if x.dtype != torch.float32:
        raise RuntimeError("fused_bias_act: only float32 is supported")
    x = x.contiguous()
    if bias is not None and bias.numel() == 0:
        bias = None
    if ref is not None and ref.numel() == 0:
        ref = None
    n, step_b, size_b = _geometry(x)
    if bias is not None:
        if bias.device != x.device or bias.dtype != torch.float32:
            raise RuntimeError("bias must be a float32 CUDA tensor on the input's device")
        if bias.numel() != size_b:
            raise RuntimeError(f"bias has {bias.numel()} elements, expected {size_b} (dim 1 of the input)")
        bias = bias.contiguous()
    if ref is not None:
        if ref.shape != x.shape:
            raise RuntimeError("refer must have the input's shape")
        ref = ref.contiguous()
    y = torch.empty_like(x)
    if n == 0:
        return y
    with _lib.device_guard(x.device):
        rc = _lib.load().vsp_bias_act_f32(_lib.ptr(x), _lib.ptr(bias), _lib.ptr(ref), _lib.ptr(y), n, step_b, size_b,
                                          act, grad, alpha, scale, _lib.stream_ptr())
    _lib.check(rc, "bias_act")
    return y


def bias_act_bwd_raw(grad_out, out, want_bias, alpha, scale):
    grad_out = grad_out.contiguous()
    out = out.contiguous()
    n, step_b, size_b = _geometry(grad_out)
    dx = torch.empty_like(grad_out)
    dbias = torch.empty(size_b, dtype=torch.float32, device=grad_out.device) if want_bias else None
    if n == 0:
        if dbias is not None:
            dbias.zero_()
        return dx, dbias
    with _lib.device_guard(grad_out.device):
        rc = _lib.load().vsp_bias_act_bwd_f32(_lib.ptr(grad_out), _lib.ptr(out), _lib.ptr(dx), _lib.ptr(dbias),
                                              n, step_b, size_b, alpha, scale, _lib.stream_ptr())
    _lib.check(rc, "bias_act_bwd")
    return dx, dbias


class FusedLeakyReLUFunctionBackward(Function):
    """op/fused_act.py:126-165."""

    @staticmethod
    def forward(ctx, grad_output, out, bias, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input, grad_bias = bias_act_bwd_raw(grad_output, out, bias, negative_slope, scale)
        if grad_bias is None:
            grad_bias = grad_output.new_empty(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        (out,) = ctx.saved_tensors
        gg_bias = gradgrad_bias if (gradgrad_bias is not None and gradgrad_bias.numel()) else None
        gradgrad_out = _SecondOrder.apply(gradgrad_input, gg_bias, out, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None, None


class _SecondOrder(Function):
    """gradgrad_out = scale*lrelu'(out)*(gg_in + gg_b): linear in (gg_in, gg_b), so it is its
    own closed set (its backward is the first-order backward again)."""

    @staticmethod
    def forward(ctx, gg_in, gg_bias, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.has_bias = gg_bias is not None
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return bias_act_raw(gg_in, gg_bias, out, 3, 1, negative_slope, scale)

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        gi, gb = FusedLeakyReLUFunctionBackward.apply(g, out, ctx.has_bias, ctx.negative_slope, ctx.scale)
        return gi, (gb if ctx.has_bias else None), None, None, None


class FusedLeakyReLUFunction(Function):
    """op/fused_act.py:168-196."""

    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        ctx.bias = bias is not None
        out = bias_act_raw(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.bias,
                                                                     ctx.negative_slope, ctx.scale)
        if not ctx.bias:
            grad_bias = None
        return grad_input, grad_bias, None, None


class FusedUpFirDnLeakyReLU(Function):
    """``fused_leaky_relu(upfirdn2d(x, k, up, down, pad), bias)`` in one kernel (forward); the
    backward chains the two reference backward formulas."""

    @staticmethod
    def forward(ctx, input, kernel, bias, up, down, pad, negative_slope, scale):
        from .upfirdn2d import upfirdn2d_raw

        out = upfirdn2d_raw(input, kernel, up, down, pad, bias=bias, act=3, alpha=negative_slope, scale=scale)
        ctx.save_for_backward(kernel, out)
        ctx.has_bias = bias is not None
        ctx.cfg = (up, down, pad, negative_slope, scale, tuple(input.shape))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        from .upfirdn2d import UpFirDn2dBackward

        kernel, out = ctx.saved_tensors
        up, down, pad, slope, scale, in_size = ctx.cfg
        g_act, g_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.has_bias, slope, scale)
        grad_input = None
        if ctx.needs_input_grad[0]:
            kh, kw = kernel.shape
            out_h, out_w = out.shape[2], out.shape[3]
            g_pad = (kw - pad[0] - 1, in_size[3] * up[0] - out_w * down[0] + pad[0] - up[0] + 1,
                     kh - pad[2] - 1, in_size[2] * up[1] - out_h * down[1] + pad[2] - up[1] + 1)
            grad_input = UpFirDn2dBackward.apply(g_act, kernel, torch.flip(kernel, [0, 1]), up, down, pad, g_pad,
                                                 in_size, (out_h, out_w))
        return grad_input, None, (g_bias if ctx.has_bias else None), None, None, None, None, None


class FusedLeakyReLU(nn.Module):
    """op/fused_act.py:199-213 — parameter name ``bias`` kept for state_dict compatibility."""

    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        if bias:
            self.bias = nn.Parameter(torch.zeros(channel))
        else:
            self.bias = None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:216-233 (CUDA branch)."""
    if not (torch.is_grad_enabled() and (input.requires_grad or (bias is not None and bias.requires_grad))):
        return bias_act_raw(input, bias, None, 3, 0, negative_slope, scale)     # nothing to record
    return FusedLeakyReLUFunction.apply(input.contiguous(), bias, negative_slope, scale)
