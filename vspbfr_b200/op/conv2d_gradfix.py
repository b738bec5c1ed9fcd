"""conv2d_gradfix — convolution with arbitrarily-high-order gradients through a CLOSED set of ops.

Mirrors /root/reference/op/conv2d_gradfix.py: same module-level API (``conv2d``,
``conv_transpose2d``, ``no_weight_gradients()``, ``enabled``, ``weight_gradients_disabled``) and
the same autograd structure (:134-223): forward conv, input-gradient = the transposed op with
``calc_output_padding`` (:122-132), weight-gradient = a dedicated op whose own backward closes
the set.  The reference's custom path is dead on torch >= 1.9 (``could_use_op`` :78-92 falls back
to plain ``F.conv2d``, and ``no_weight_gradients`` becomes a no-op); here it is live again.

Backends for the three primitive ops (fprop / dgrad / wgrad):
  * ``"tcgen05"`` — the bf16 implicit-GEMM kernels of this repo (csrc/conv_sm100.cu,
    csrc/wgrad_sm100.cu) for the shapes they cover (see ``_tc_supported``);
  * ``"aten"``    — ATen/cuDNN fp32 for everything else (tiny channel counts such as Cin=3,
    exotic strides).  This is the library baseline, not a CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import contextlib

import torch
from torch import autograd
from torch.nn import functional as F

enabled = True
weight_gradients_disabled = False
backend = "tcgen05"  # or "aten"


@contextlib.contextmanager
def no_weight_gradients():
    """op/conv2d_gradfix.py:12-19."""
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    try:
        yield
    finally:
        weight_gradients_disabled = old


def _tuple(xs, ndim=2):
    return tuple(xs) if isinstance(xs, (tuple, list)) else (xs,) * ndim


def _check(input):
    if input.device.type != "cuda":
        raise RuntimeError("conv2d_gradfix: input must be a CUDA tensor (vspbfr_b200 has no CPU path)")


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """op/conv2d_gradfix.py:22-42."""
    _check(input)
    if not enabled:
        return F.conv2d(input, weight, bias, stride, padding, dilation, groups)
    return conv2d_gradfix(False, weight.shape, stride, padding, 0, dilation, groups).apply(input, weight, bias)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """op/conv2d_gradfix.py:45-75."""
    _check(input)
    if not enabled:
        return F.conv_transpose2d(input, weight, bias, stride, padding, output_padding, groups, dilation)
    return conv2d_gradfix(True, weight.shape, stride, padding, output_padding, dilation, groups).apply(
        input, weight, bias)


conv2d_gradfix_cache = dict()


def conv2d_gradfix(transpose, weight_shape, stride, padding, output_padding, dilation, groups):
    """Factory of the closed autograd set, cached per configuration (op/conv2d_gradfix.py:104-227)."""
    ndim = 2
    weight_shape = tuple(weight_shape)
    stride = _tuple(stride)
    padding = _tuple(padding)
    output_padding = _tuple(output_padding)
    dilation = _tuple(dilation)
    key = (transpose, weight_shape, stride, padding, output_padding, dilation, groups)
    if key in conv2d_gradfix_cache:
        return conv2d_gradfix_cache[key]

    common = dict(stride=stride, padding=padding, dilation=dilation, groups=groups)

    def calc_output_padding(input_shape, output_shape):
        if transpose:
            return [0, 0]
        return [input_shape[i + 2] - (output_shape[i + 2] - 1) * stride[i] - (1 - 2 * padding[i])
                - dilation[i] * (weight_shape[i + 2] - 1) for i in range(ndim)]

    class Conv2d(autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias):
            from . import _conv_backend as cb

            if not transpose:
                out = cb.fprop(input, weight, bias, **common)
            else:
                out = cb.fprop_transposed(input, weight, bias, output_padding=output_padding, **common)
            ctx.save_for_backward(input, weight)
            return out

        @staticmethod
        def backward(ctx, grad_output):
            input, weight = ctx.saved_tensors
            grad_input = grad_weight = grad_bias = None
            if ctx.needs_input_grad[0]:
                p = calc_output_padding(input.shape, grad_output.shape)
                grad_input = conv2d_gradfix(not transpose, weight_shape, output_padding=p, **common).apply(
                    grad_output, weight, None)
            if ctx.needs_input_grad[1] and not weight_gradients_disabled:
                grad_weight = Conv2dGradWeight.apply(grad_output, input)
            if ctx.needs_input_grad[2]:
                grad_bias = grad_output.sum((0, 2, 3))
            return grad_input, grad_weight, grad_bias

    class Conv2dGradWeight(autograd.Function):
        @staticmethod
        def forward(ctx, grad_output, input):
            from . import _conv_backend as cb

            grad_weight = cb.wgrad(grad_output, input, weight_shape, transpose=transpose,
                                   output_padding=output_padding, **common)
            ctx.save_for_backward(grad_output, input)
            return grad_weight

        @staticmethod
        def backward(ctx, grad_grad_weight):
            grad_output, input = ctx.saved_tensors
            grad_grad_output = grad_grad_input = None
            if ctx.needs_input_grad[0]:
                grad_grad_output = Conv2d.apply(input, grad_grad_weight, None)
            if ctx.needs_input_grad[1]:
                p = calc_output_padding(input.shape, grad_output.shape)
                grad_grad_input = conv2d_gradfix(not transpose, weight_shape, output_padding=p, **common).apply(
                    grad_output, grad_grad_weight, None)
            return grad_grad_output, grad_grad_input

    conv2d_gradfix_cache[key] = Conv2d
    return Conv2d
