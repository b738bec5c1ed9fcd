"""upfirdn2d — host side of the sm_100a up-FIR-down resampler.

Mirrors /root/reference/op/upfirdn2d.py:217-362: same public function
``upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))`` (scalar or ``(x, y)`` factors, 2- or
4-pad incl. negative pads), same autograd structure — the op is linear, so its gradient is
the same op with swapped factors, flipped filter and ``g_pad`` (:309-312), and the gradient
of that is the original op again (closed under differentiation, any order).

Unlike the reference there is no CPU branch: CPU tensors raise.  The CUDA work happens in
``vsp_upfirdn2d_f32`` (include/vsp_b200.h), called through ctypes.
"""
from __future__ import annotations

from collections import abc

import torch
from torch.autograd import Function

from .. import _lib


def _check_cuda(t, name):
    if t.device.type != "cuda":
        raise RuntimeError(f"{name} must be a CUDA tensor (vspbfr_b200 has no CPU path)")


def upfirdn2d_raw(input, kernel, up, down, pad, bias=None, act=0, alpha=0.2, scale=1.0):
    """One kernel launch: [N,C,H,W] fp32 -> [N,C,H',W'] fp32 (optionally + bias + lrelu*scale)."""
    _check_cuda(input, "input")
    _check_cuda(kernel, "kernel")
    if input.dtype != torch.float32 or kernel.dtype != torch.float32:
        raise RuntimeError("upfirdn2d: only float32 input/kernel are supported")
    if input.ndim != 4 or kernel.ndim != 2:
        raise RuntimeError("upfirdn2d: expected input [N,C,H,W] and kernel [kh,kw]")
    up_x, up_y = up
    down_x, down_y = down
    pad_x0, pad_x1, pad_y0, pad_y1 = pad
    n, c, in_h, in_w = input.shape
    kh, kw = kernel.shape
    lib = _lib.load()
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) // down_y          # op/upfirdn2d.py:301-302 (== vsp_upfirdn2d_out_size)
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) // down_x
    x = input.contiguous()
    k = kernel.contiguous()
    out = torch.empty((n, c, max(out_h, 0), max(out_w, 0)), dtype=torch.float32, device=input.device)
    if out.numel() == 0:
        return out
    if bias is not None:
        bias = bias.contiguous()
    with _lib.device_guard(input.device):
        rc = lib.vsp_upfirdn2d_f32(_lib.ptr(x), _lib.ptr(k), _lib.ptr(out), n * c, in_h, in_w, kh, kw,
                                   up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1,
                                   _lib.ptr(bias), c, act, alpha, scale, _lib.stream_ptr())
    _lib.check(rc, "upfirdn2d")
    return out


class UpFirDn2dBackward(Function):
    """grad_input = upfirdn2d(grad_out, flip(k), up=down, down=up, pad=g_pad); op/upfirdn2d.py:217-283."""

    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        grad_input = upfirdn2d_raw(grad_output.reshape(in_size[0], in_size[1], out_size[0], out_size[1]),
                                   grad_kernel, down, up, g_pad)
        # the transposed op may produce a larger extent than the input when `down` does not
        # divide evenly; the reference's g_pad formula makes them equal (op/upfirdn2d.py:309-312)
        grad_input = grad_input.view(in_size[0], in_size[1], in_size[2], in_size[3])
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.in_size, ctx.out_size = in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        (kernel,) = ctx.saved_tensors
        gradgrad_out = UpFirDn2d.apply(gradgrad_input.reshape(ctx.in_size), kernel, ctx.up, ctx.down, ctx.pad)
        return gradgrad_out, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    """op/upfirdn2d.py:286-343."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        _, _, in_h, in_w = input.shape
        ctx.in_size = tuple(input.shape)
        out = upfirdn2d_raw(input, kernel, up, down, pad)
        out_h, out_w = out.shape[2], out.shape[3]
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        ctx.g_pad = (kernel_w - pad_x0 - 1,
                     in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                     kernel_h - pad_y0 - 1,
                     in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)
        # the flipped filter is only needed by the backward (the reference flips on every call, :299)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]) if ctx.needs_input_grad[0] else kernel)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = None
        if ctx.needs_input_grad[0]:
            grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad,
                                                 ctx.g_pad, ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None  # no kernel gradient, as the reference (:343)


def _normalize(up, down, pad):
    if not isinstance(up, abc.Iterable):
        up = (up, up)
    if not isinstance(down, abc.Iterable):
        down = (down, down)
    pad = tuple(int(p) for p in pad)
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    return (int(up[0]), int(up[1])), (int(down[0]), int(down[1])), pad


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """Same signature and semantics as op/upfirdn2d.py:346-362 (CUDA only)."""
    up, down, pad = _normalize(up, down, pad)
    if not (torch.is_grad_enabled() and input.requires_grad):
        return upfirdn2d_raw(input, kernel, up, down, pad)          # nothing to record: skip the autograd.Function round trip
    return UpFirDn2d.apply(input, kernel, up, down, pad)


def upfirdn2d_bias_act(input, kernel, bias=None, up=1, down=1, pad=(0, 0), negative_slope=0.2, scale=2 ** 0.5):
    """Fused ``fused_leaky_relu(upfirdn2d(x), bias)`` in ONE pass (north_star's optional fused
    epilogue).  Differentiable through the composition of the two autograd Functions'
    backward formulas."""
    from .fused_act import FusedUpFirDnLeakyReLU

    up, down, pad = _normalize(up, down, pad)
    return FusedUpFirDnLeakyReLU.apply(input, kernel, bias, up, down, pad, negative_slope, scale)
