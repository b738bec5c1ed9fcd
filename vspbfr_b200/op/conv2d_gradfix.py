"""conv2d_gradfix — convolutions whose gradients of any order stay inside a closed set of tcgen05 ops.

Public surface of /root/reference/op/conv2d_gradfix.py: ``conv2d`` (:22-42), ``conv_transpose2d`` (:45-75),
``no_weight_gradients()`` (:12-19), the flags ``enabled`` (:8) and ``weight_gradients_disabled`` (:9).  The reference's
custom path is dead on torch >= 1.9 (``could_use_op`` :78-92 falls back to ``F.conv2d``, so ``no_weight_gradients`` does
nothing there); here it is live, and every member of the set runs on this repo's kernels (``_plainconv``):

    ConvOp(spec)          y  = conv(x, w)          forward or transposed, described by a ``ConvSpec``
      d/dx                dx = ConvOp(spec^T)(dy, w)           the transposed op, output padding chosen to restore x's extent
      d/dw                dw = WeightGradOp(spec)(dy, x)       skipped inside ``no_weight_gradients()``
    WeightGradOp(spec)    dw = wgrad(dy, x)
      d/d(dy)             ConvOp(spec)(x, ddw)                 linear in w: the same conv with the incoming cotangent as weight
      d/dx                ConvOp(spec^T)(dy, ddw)

so R1 / path-length double backward (restoration_train.py:66-73, :200-216) never leaves the set.  Two autograd Functions
take the spec as an argument; there is no per-configuration class factory or cache.
"""
from __future__ import annotations

import contextlib

import torch
from torch.nn import functional as F

from ._plainconv import ConvSpec, conv_forward, conv_weight_grad

enabled = True                      # False: plain torch.nn.functional ops (what the reference does on torch >= 1.9)
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    """Skip weight gradients inside the block (R1: only d(score)/d(image) is wanted, restoration_train.py:66-73)."""
    global weight_gradients_disabled
    saved, weight_gradients_disabled = weight_gradients_disabled, True
    try:
        yield
    finally:
        weight_gradients_disabled = saved


def _pair(v):
    if isinstance(v, (tuple, list)):
        if len(v) != 2:
            raise ValueError(f"conv2d_gradfix: expected an int or a pair, got {v!r}")
        return int(v[0]), int(v[1])
    return int(v), int(v)


def _require_cuda(input):
    if input.device.type != "cuda":
        raise RuntimeError("conv2d_gradfix: input must be a CUDA tensor (vspbfr_b200 has no CPU path)")


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    _require_cuda(input)
    if not enabled:
        return F.conv2d(input, weight, bias, stride, padding, dilation, groups)
    spec = ConvSpec(False, _pair(stride), _pair(padding), (0, 0), _pair(dilation), int(groups))
    return ConvOp.apply(input, weight, bias, spec)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    _require_cuda(input)
    if not enabled:
        return F.conv_transpose2d(input, weight, bias, stride, padding, output_padding, groups, dilation)
    spec = ConvSpec(True, _pair(stride), _pair(padding), _pair(output_padding), _pair(dilation), int(groups))
    return ConvOp.apply(input, weight, bias, spec)


def _adjoint(spec: ConvSpec, x_shape, y_shape, k_shape) -> ConvSpec:
    """Spec of the op mapping a cotangent of y back to x's extent.  For a forward conv that is the transposed op whose
    ``output_padding`` makes up for the rows the strided forward never reached (op/conv2d_gradfix.py:122-132); for a
    transposed conv it is the plain forward conv."""
    if spec.transpose:
        return spec._replace(transpose=False, output_padding=(0, 0))
    extra = tuple(x_shape[2 + a] - ((y_shape[2 + a] - 1) * spec.stride[a] - 2 * spec.padding[a]
                                    + spec.dilation[a] * (k_shape[2 + a] - 1) + 1) for a in range(2))
    return spec._replace(transpose=True, output_padding=extra)


class ConvOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weight, bias, spec):
        ctx.spec = spec
        ctx.save_for_backward(input, weight)
        return conv_forward(input, weight, bias, spec)

    @staticmethod
    def backward(ctx, grad_output):
        input, weight = ctx.saved_tensors
        spec = ctx.spec
        want_x, want_w, want_b = ctx.needs_input_grad[:3]
        dx = dw = db = None
        if want_x:
            dx = ConvOp.apply(grad_output, weight, None, _adjoint(spec, input.shape, grad_output.shape, weight.shape))
        if want_w and not weight_gradients_disabled:
            dw = WeightGradOp.apply(grad_output, input, spec, tuple(weight.shape))
        if want_b:
            db = grad_output.sum((0, 2, 3))
        return dx, dw, db, None


class WeightGradOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_output, input, spec, weight_shape):
        ctx.spec, ctx.weight_shape = spec, weight_shape
        ctx.save_for_backward(grad_output, input)
        return conv_weight_grad(grad_output, input, weight_shape, spec)

    @staticmethod
    def backward(ctx, grad_grad_weight):
        grad_output, input = ctx.saved_tensors
        spec = ctx.spec
        gg_out = gg_in = None
        if ctx.needs_input_grad[0]:
            gg_out = ConvOp.apply(input, grad_grad_weight, None, spec)
        if ctx.needs_input_grad[1]:
            gg_in = ConvOp.apply(grad_output, grad_grad_weight, None,
                                 _adjoint(spec, input.shape, grad_output.shape, ctx.weight_shape))
        return gg_out, gg_in, None, None
