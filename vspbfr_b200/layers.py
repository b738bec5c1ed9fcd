"""Synthesis layers of the hot path, rebuilt on the sm_100a ops.

Host-side mirror of the reference's layer classes: same class names, constructor signatures,
parameter/buffer names and shapes (so reference checkpoints load with ``strict=True``), same
forward semantics — but every forward goes through ``vspbfr_b200.op`` (hand-written CUDA behind the
C ABI).  Citations are to /root/reference/models/RestoreNet.py; the e4e twins
(e4e/models/stylegan2/model.py:23-364) have the same structure and are served by the same classes.

These forwards are the differentiable NCHW-fp32 path (training, R1 double backward).  Inference
uses the fused channels-last pipeline in ``vspbfr_b200.fastpath`` that walks the same modules.
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn import functional as F

from .op import FusedLeakyReLU, conv2d_gradfix, fused_leaky_relu, upfirdn2d
from .op.modconv import modulate_input, modulated_conv2d


def make_kernel(k):
    """Normalised 2-D FIR from 1-D taps (models/RestoreNet.py:32-40)."""
    k = torch.as_tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def _updown_pads(blur_kernel, kernel_size, dilation=1):
    """Blur pads around the stride-2 (transposed) convs (models/RestoreNet.py:443-457, :299-313)."""
    p_up = (len(blur_kernel) - 2) - (kernel_size - 1) * dilation
    p_dn = (len(blur_kernel) - 2) + (kernel_size - 1)
    return ((p_up + 1) // 2 + 1, p_up // 2 + 1), ((p_dn + 1) // 2, p_dn // 2)


class PixelNorm(nn.Module):
    def forward(self, input):
        return input * torch.rsqrt(input.square().mean(dim=1, keepdim=True) + 1e-8)


class Upsample(nn.Module):
    """2x FIR upsampling of the RGB skip (models/RestoreNet.py:43-61)."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel) * (factor ** 2))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    """models/RestoreNet.py:64-82 (never instantiated by the reference networks)."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    """models/RestoreNet.py:85-101."""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualConv2d(nn.Module):
    """Equalised-lr convolution (models/RestoreNet.py:104-139; DilatedEqualConv2d :683-722)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True, dilation=1):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        return conv2d_gradfix.conv2d(input, self.weight * self.scale, bias=self.bias, stride=self.stride,
                                     padding=self.padding, dilation=self.dilation)

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]}, {self.weight.shape[2]},"
                f" stride={self.stride}, padding={self.padding})")


class DilatedEqualConv2d(EqualConv2d):
    def __init__(self, in_channel, out_channel, kernel_size, padding=0, stride=1, dilation=1, bias=True):
        super().__init__(in_channel, out_channel, kernel_size, stride=stride, padding=padding, bias=bias,
                         dilation=dilation)


class EqualLinear(nn.Module):
    """models/RestoreNet.py:142-176."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        b = self.bias * self.lr_mul if self.bias is not None else None
        if self.activation:
            return fused_leaky_relu(F.linear(input, self.weight * self.scale), b)
        return F.linear(input, self.weight * self.scale, bias=b)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class _ModulatedBase(nn.Module):
    """Shared body of ModulatedConv2d / Dilated_ModulatedConv2d: weight [1,Cout,Cin,k,k], optional
    blur around the stride-2 forms, forward through the fused tcgen05 Function."""

    def _setup(self, in_channel, out_channel, kernel_size, demodulate, upsample, downsample, blur_kernel, dilation):
        self.eps = 1e-8
        self.kernel_size, self.in_channel, self.out_channel = kernel_size, in_channel, out_channel
        self.upsample, self.downsample = upsample, downsample
        self.blur_kernel = blur_kernel
        self.dilation = dilation
        up_pad, down_pad = _updown_pads(blur_kernel, kernel_size, dilation)
        if upsample:
            self.blur = Blur(blur_kernel, pad=up_pad, upsample_factor=2)
        if downsample:
            self.blur = Blur(blur_kernel, pad=down_pad)
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = ((kernel_size - 1) * dilation) // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.demodulate = demodulate

    def _conv(self, input, s, xs=None):
        """``xs``: ``modulate_input(input, s)`` when the caller shares it between branches (stride-1 form only)."""
        cache = self.__dict__.setdefault("_derived", {})      # sum_t W^2 / packed weights, revalidated per call (not in state_dict)
        if self.upsample:
            return self.blur(modulated_conv2d(input, self.weight, s, self.demodulate, "up", self.dilation, eps=self.eps,
                                              cache=cache))
        if self.downsample:
            return modulated_conv2d(self.blur(input), self.weight, s, self.demodulate, "down", self.dilation, eps=self.eps,
                                    cache=cache)
        return modulated_conv2d(input, self.weight, s, self.demodulate, "same", self.dilation, xs=xs, eps=self.eps,
                                cache=cache)

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample}, downsample={self.downsample})")


class ModulatedConv2d(_ModulatedBase):
    """models/RestoreNet.py:421-555 (both its fused and non-fused branches compute this)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1], fused=True):
        super().__init__()
        self._setup(in_channel, out_channel, kernel_size, demodulate, upsample, downsample, blur_kernel, 1)
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.fused = fused

    def forward(self, input, style):
        return self._conv(input, self.modulation(style))


class Dilated_ModulatedConv2d(_ModulatedBase):
    """models/RestoreNet.py:270-418: takes the ALREADY modulated style (no own ``modulation``)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1], fused=True, dilation=1):
        super().__init__()
        self._setup(in_channel, out_channel, kernel_size, demodulate, upsample, downsample, blur_kernel, dilation)
        self.fused = fused

    def forward(self, input, style):
        return self._conv(input, style)


class NoiseInjection(nn.Module):
    """models/RestoreNet.py:558-569."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            b, _, h, w = image.shape
            noise = image.new_empty(b, 1, h, w).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    """e4e/models/stylegan2/model.py:295-305."""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    """conv -> noise -> bias + leaky ReLU (models/RestoreNet.py:571-605)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True, downsample=False):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    downsample=downsample, blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None):
        return self.activate(self.noise(self.conv(input, style), noise=noise))


class StyledConv_down(StyledConv):
    """models/RestoreNet.py:608-643: ``ModulatedConv2d(downsample=True)``; the ``upsample`` argument
    is accepted and ignored, as in the reference."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__(in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=blur_kernel,
                         demodulate=demodulate, downsample=True)


class ToRGB(nn.Module):
    """1x1 modulated conv (no demod) + bias + upsampled skip (models/RestoreNet.py:647-666)."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None):
        out = self.conv(input, style) + self.bias
        if skip is not None:
            out = out + self.upsample(skip)
        return out


class ConvLayer(nn.Sequential):
    """[Blur] + EqualConv2d + [FusedLeakyReLU] (models/RestoreNet.py:1137-1179)."""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=[1, 3, 3, 1], bias=True,
                 activate=True):
        layers = []
        if downsample:
            _, down_pad = _updown_pads(blur_kernel, kernel_size)
            layers.append(Blur(blur_kernel, pad=down_pad))
            stride, self.padding = 2, 0
        else:
            stride, self.padding = 1, kernel_size // 2
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding, stride=stride,
                                  bias=bias and not activate))
        if activate:
            layers.append(FusedLeakyReLU(out_channel, bias=bias))
        super().__init__(*layers)


class SMART_layer(nn.Module):
    """Four dilated modulated branches sharing one modulation, concatenated, fused by a 3x3 ConvLayer,
    then noise + bias/leaky-ReLU (models/RestoreNet.py:179-268)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True, rates=[1, 2, 4, 8], activate=True):
        super().__init__()
        self.rates = rates
        self.ModulatedConv2ds = nn.ModuleList(
            Dilated_ModulatedConv2d(in_channel, out_channel // len(rates), kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate, dilation=rate)
            for rate in rates)
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.fusion = ConvLayer(out_channel, out_channel, 3)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel) if activate else None

    def _branches(self, input, style):
        s = self.modulation(style)
        if any(b.upsample or b.downsample for b in self.ModulatedConv2ds):
            return [branch(input, s) for branch in self.ModulatedConv2ds]
        xs = modulate_input(input, s)           # the four branches read the same modulated activation: convert it once
        return [branch._conv(input, s, xs=xs) for branch in self.ModulatedConv2ds]

    def forward(self, input, style, noise=None):
        out = self.noise(self.fusion(torch.cat(self._branches(input, style), dim=1)), noise=noise)
        return self.activate(out) if self.activate is not None else out

    def forward_vis(self, input, style, noise=None):
        """As ``forward`` but also returns the branch outputs + result (models/RestoreNet.py:246-268)."""
        outs = self._branches(input, style)
        out = self.noise(self.fusion(torch.cat(outs, dim=1)), noise=noise)
        if self.activate is not None:
            out = self.activate(out)
        outs.append(out)
        return out, outs


class LargeConvLayer(nn.Sequential):
    """Un-modulated counterpart of SMART_layer (models/RestoreNet.py:725-787)."""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=[1, 3, 3, 1], bias=True,
                 activate=True, rates=[1, 2, 4, 8]):
        super().__init__()
        self.downsample = downsample
        if downsample:
            _, down_pad = _updown_pads(blur_kernel, kernel_size)
            self.blur = Blur(blur_kernel, pad=down_pad)
        self.dilated_convs = nn.ModuleList()
        for rate in rates:
            stride = 2 if downsample else 1
            self.padding = ((kernel_size - 1) * rate - (stride if downsample else 0)) // 2
            self.dilated_convs.append(DilatedEqualConv2d(in_channel, out_channel // len(rates), kernel_size,
                                                         padding=self.padding, stride=stride, dilation=rate,
                                                         bias=bias and not activate))
        self.fusion = ConvLayer(out_channel, out_channel, 1)
        self.activate = FusedLeakyReLU(out_channel, bias=bias) if activate else None

    def forward(self, input):
        if self.downsample:
            input = self.blur(input)
        out = self.fusion(torch.cat([conv(input) for conv in self.dilated_convs], dim=1))
        return self.activate(out) if self.activate is not None else out


class ResBlock(nn.Module):
    """Discriminator block (models/RestoreNet.py:1182-1200)."""

    def __init__(self, in_channel, out_channel, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=True)
        self.skip = ConvLayer(in_channel, out_channel, 1, downsample=True, activate=False, bias=False)

    def forward(self, input):
        return (self.conv2(self.conv1(input)) + self.skip(input)) / math.sqrt(2)
