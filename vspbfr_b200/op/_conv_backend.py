"""Primitive convolution ops behind conv2d_gradfix: fprop, transposed fprop (dgrad) and wgrad.

NCHW fp32 at the boundary (that is what the reference's layers pass,
models/RestoreNet.py:125-131, :547-553).  Shapes the tcgen05 kernels cover run there (bf16
operands, fp32 accumulate); everything else goes to ATen/cuDNN in fp32 — the same library call
the reference makes (op/conv2d_gradfix.py:34-42) — never to a CPU path.
"""
from __future__ import annotations

import torch
from torch.nn import functional as F

from . import conv2d_gradfix as _cfg


def _tc_supported(input, weight_shape, stride, padding, dilation, groups, transpose):
    """Shapes routed to the tcgen05 implicit GEMM (csrc/conv_sm100.cu)."""
    if _cfg.backend != "tcgen05":
        return False
    try:
        from . import modconv
    except Exception:
        return False
    return modconv.plain_conv_supported(input, weight_shape, stride, padding, dilation, groups, transpose)


def fprop(input, weight, bias, stride, padding, dilation, groups):
    if _tc_supported(input, weight.shape, stride, padding, dilation, groups, False):
        from . import modconv

        return modconv.plain_conv_fprop(input, weight, bias, stride, padding, dilation, groups)
    return F.conv2d(input, weight, bias, stride, padding, dilation, groups)


def fprop_transposed(input, weight, bias, stride, padding, dilation, groups, output_padding):
    if _tc_supported(input, weight.shape, stride, padding, dilation, groups, True):
        from . import modconv

        return modconv.plain_conv_dgrad(input, weight, bias, stride, padding, dilation, groups, output_padding)
    return F.conv_transpose2d(input, weight, bias, stride, padding, output_padding, groups, dilation)


def wgrad(grad_output, input, weight_shape, stride, padding, dilation, groups, transpose, output_padding):
    """Replaces aten::cudnn_convolution(_transpose)_backward_weight (op/conv2d_gradfix.py:180-199),
    an operator that no longer exists in torch 2.x."""
    if not transpose and _tc_supported(input, weight_shape, stride, padding, dilation, groups, False):
        from . import modconv

        gw = modconv.plain_conv_wgrad(grad_output, input, weight_shape, stride, padding, dilation, groups)
        if gw is not None:
            return gw
    w_stub = torch.empty(weight_shape, dtype=input.dtype, device=input.device)
    go, x = grad_output.contiguous(), input.contiguous()
    grads = torch.ops.aten.convolution_backward(
        go, x, w_stub, None, list(stride), list(padding), list(dilation), transpose, list(output_padding), groups,
        [False, True, False])
    return grads[1]
