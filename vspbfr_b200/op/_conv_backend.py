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
    go, x = grad_output.contiguous(), input.contiguous()
    cout, cin, kh, kw = weight_shape
    hw = x.shape[2] * x.shape[3]
    if (not transpose and groups == 1 and kh == 1 and kw == 1 and cin <= 8 and tuple(stride) == (1, 1)
            and tuple(padding) == (0, 0) and x.dtype == torch.float32 and go.dtype == torch.float32 and x.is_cuda
            and hw % 4 == 0 and x.shape[0] <= 65535 and _cfg.backend == "tcgen05"):
        # RGB-side 1x1 layers: streaming reduction instead of the library's tall-skinny GEMM (3.5 ms -> tens of us at 512^2)
        from .. import _lib

        gw = torch.empty(weight_shape, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().vsp_conv1x1_wgrad_small_f32(_lib.ptr(go), _lib.ptr(x), _lib.ptr(gw), x.shape[0], cout, cin, hw,
                                                         _lib.stream_ptr())
        _lib.check(rc, "conv1x1_wgrad_small_f32")
        return gw
    w_stub = torch.empty(weight_shape, dtype=input.dtype, device=input.device)
    grads = torch.ops.aten.convolution_backward(
        go, x, w_stub, None, list(stride), list(padding), list(dilation), transpose, list(output_padding), groups,
        [False, True, False])
    return grads[1]
