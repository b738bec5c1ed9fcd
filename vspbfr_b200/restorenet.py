"""Restoration_net and Discriminator on the sm_100a ops.

Mirror of /root/reference/models/RestoreNet.py:791-1046 (``Restoration_net``) and :1205-1265
(``Discriminator``): same constructor/forward signatures, same ``state_dict`` keys and shapes
(checked against tests/golden/state_dict_manifest.npz), same construction order (so a given seed
yields the same random initialisation).  ``forward`` here is the differentiable NCHW-fp32 path;
``vspbfr_b200.fastpath.restoration_forward`` runs the same modules as a fused channels-last
bf16 pipeline for inference.
"""
from __future__ import annotations

import math
import random

import torch
from torch import nn

from .layers import (ConvLayer, EqualLinear, LargeConvLayer, PixelNorm, ResBlock, SMART_layer, StyledConv,
                     StyledConv_down, ToRGB)


def channel_table(channel_multiplier=2):
    """Channels per resolution (models/RestoreNet.py:810-820)."""
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier, 128: 128 * channel_multiplier,
            256: 64 * channel_multiplier, 512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}


def make_noise(batch, latent_dim, n_noise, device):
    if n_noise == 1:
        return torch.randn(batch, latent_dim, device=device)
    return torch.randn(n_noise, batch, latent_dim, device=device).unbind(0)


def mixing_noise(batch, latent_dim, prob, device):
    """models/RestoreNet.py:17-22."""
    if prob > 0 and random.random() < prob:
        return make_noise(batch, latent_dim, 2, device)
    return [make_noise(batch, latent_dim, 1, device)]


def assemble_latent(styles, n_latent, inject_index=None):
    """Broadcast / mix the mapped noise styles to [B, n_latent, D] (models/RestoreNet.py:995-1011)."""
    if len(styles) < 2:
        if styles[0].ndim < 3:
            return styles[0].unsqueeze(1).repeat(1, n_latent, 1)
        return styles[0]
    if inject_index is None:
        inject_index = random.randint(1, n_latent - 1)
    first = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
    second = styles[1].unsqueeze(1).repeat(1, n_latent - inject_index, 1)
    return torch.cat([first, second], 1)


class Restoration_net(nn.Module):
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        self.channels = channel_table(channel_multiplier)
        self.blur_kernel = blur_kernel
        wide = 4 * style_dim  # decoder styles: [w+ | mapped noise | x_global(2*style_dim)]
        self.conv1 = SMART_layer(self.channels[4], self.channels[4], 3, wide, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], wide, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        mapping = [PixelNorm()]
        mapping += [EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu") for _ in range(n_mlp)]
        self.style = nn.Sequential(*mapping)
        for layer_idx in range(self.num_layers):
            res = 2 ** ((layer_idx + 5) // 2)
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, res, res))
        in_channel = self.channels[4]
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, wide, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(SMART_layer(out_channel, out_channel, 3, wide, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, wide))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2
        self.encoder_res = [2 ** i for i in range(int(math.log2(size)), 1, -1)]
        self.encoder(im_size=size, channels=self.channels, nc=4, num_styles=self.n_latent, style_channels=style_dim)

    def encoder(self, im_size, channels, nc, num_styles, style_channels, ndf=32):
        """Builds the style-modulated encoder (models/RestoreNet.py:884-912)."""
        self.down_from_big = LargeConvLayer(3, channels[im_size], kernel_size=1)
        self.log_size = int(math.log(im_size, 2))
        in_channel = channels[im_size]
        self.encoder_convs = nn.ModuleList()
        for i in range(self.log_size, 2, -1):
            mid, out_channel = channels[2 ** i], channels[2 ** (i - 1)]
            self.encoder_convs.append(SMART_layer(in_channel, mid, 3, 2 * self.style_dim, blur_kernel=self.blur_kernel))
            self.encoder_convs.append(StyledConv_down(mid, out_channel, 3, 2 * self.style_dim,
                                                      blur_kernel=self.blur_kernel))
            in_channel = out_channel
        self.final_layer = LargeConvLayer(in_channel, channels[4], kernel_size=3)
        self.final_linear = nn.Sequential(
            EqualLinear(channels[4] * 4 * 4, channels[4] * 2, activation="fused_lrelu"), nn.Dropout2d(0.5))
        self.final_transfer = EqualLinear(channels[4] * 2, channels[4] * 4 * 4, activation="fused_lrelu")

    def encoder_forward(self, imgs, latent, noise):
        """models/RestoreNet.py:915-942 — note SMART and its down-conv share one latent index."""
        batch = imgs.shape[0]
        out = self.down_from_big(imgs)
        features = []
        for ii in range(0, len(self.encoder_convs), 2):
            out = self.encoder_convs[ii](out, latent[:, ii], noise[ii])
            features.append(out)
            out = self.encoder_convs[ii + 1](out, latent[:, ii], noise[ii + 1])
        out = self.final_layer(out)
        x_global = self.final_linear(out.view(batch, -1))
        features.append(out + self.final_transfer(x_global).view(batch, -1, 4, 4))
        return x_global, features[::-1]

    def mean_latent(self, n_latent, device):
        return self.style(torch.randn(n_latent, self.style_dim, device=device)).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def prepare_latent(self, pre_styles, noise_styles, inject_index=None, truncation=1, truncation_latent=None,
                       input_is_latent=False):
        """[B, n_latent, 2*style_dim] = cat(w+ from the e4e encoder, mapped noise) (:982-1014)."""
        if not input_is_latent:
            noise_styles = [self.style(s) for s in noise_styles]
        if truncation < 1:
            noise_styles = [truncation_latent + truncation * (s - truncation_latent) for s in noise_styles]
        noise_latent = assemble_latent(noise_styles, self.n_latent, inject_index)
        return torch.cat([pre_styles[:, :noise_latent.shape[1], :], noise_latent], dim=-1)

    def forward(self, images, de_feats, pre_styles, noise_styles, return_latents=False, inject_index=None,
                truncation=1, truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True):
        latent = self.prepare_latent(pre_styles, noise_styles, inject_index, truncation, truncation_latent,
                                     input_is_latent)
        if noise is None:
            noise = ([None] * self.num_layers if randomize_noise
                     else [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)])
        x_global, features = self.encoder_forward(images, torch.flip(latent, dims=[1]).clone(), noise[::-1])

        def sty(i):
            return torch.cat([latent[:, i], x_global], dim=1)

        out = self.conv1(features[0], sty(0), noise=noise[0])
        skip = self.to_rgb1(out, sty(1))
        i = 1
        for up, smart, n_up, n_smart, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2], noise[2::2],
                                                     self.to_rgbs):
            out = up(out, sty(i), noise=n_up)
            level = (i + 1) // 2
            out = out + features[level] + de_feats[level]
            out = smart(out, sty(i + 1), noise=n_smart)
            skip = to_rgb(out, sty(i + 2), skip)
            i += 2
        if return_latents:
            return skip, latent
        return skip


class Discriminator(nn.Module):
    """models/RestoreNet.py:1205-1265."""

    def __init__(self, size, input_channel=3, channel_multiplier=2, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        channels = channel_table(channel_multiplier)
        self.encoder_input_convs = ConvLayer(input_channel, channels[size], 1)
        self.log_size = int(math.log(size, 2))
        in_channel = channels[size]
        self.encoder_convs = nn.ModuleList()
        for i in range(self.log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            self.encoder_convs.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.stddev_group = 4
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(EqualLinear(channels[4] * 4 * 4, channels[4], activation="fused_lrelu"),
                                          EqualLinear(channels[4], 1))

    def forward(self, input):
        out = self.encoder_input_convs(input)
        for block in self.encoder_convs:
            out = block(out)
        batch, channel, height, width = out.shape
        group = min(batch, self.stddev_group)
        stddev = out.view(group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        stddev = torch.sqrt(stddev.var(0, unbiased=False) + 1e-8)
        stddev = stddev.mean([2, 3, 4], keepdims=True).squeeze(2).repeat(group, 1, height, width)
        out = self.final_conv(torch.cat([out, stddev], 1))
        return self.final_linear(out.view(batch, -1))
