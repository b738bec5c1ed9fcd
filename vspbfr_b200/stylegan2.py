"""StyleGAN2 generator used as the frozen *style decoder* (returns per-resolution features).

Mirror of /root/reference/e4e/models/stylegan2/model.py:367-552 (``Generator``) with the same
``state_dict`` layout; built from the shared layer classes of ``vspbfr_b200.layers`` (the e4e layer
twins are structurally identical to the RestoreNet ones).  ``vspbfr_b200.fastpath.generator_forward``
is the fused channels-last inference path over the same modules.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .layers import ConstantInput, EqualLinear, PixelNorm, StyledConv, ToRGB
from .restorenet import assemble_latent, channel_table


class Generator(nn.Module):
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        mapping = [PixelNorm()]
        mapping += [EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu") for _ in range(n_mlp)]
        self.style = nn.Sequential(*mapping)
        self.channels = channel_table(channel_multiplier)
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(self.num_layers):
            res = 2 ** ((layer_idx + 5) // 2)
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, res, res))
        in_channel = self.channels[4]
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2

    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 4, 4, device=device)]
        for i in range(3, self.log_size + 1):
            noises += [torch.randn(1, 1, 2 ** i, 2 ** i, device=device) for _ in range(2)]
        return noises

    def mean_latent(self, n_latent):
        return self.style(torch.randn(n_latent, self.style_dim, device=self.input.input.device)).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def prepare_latent(self, styles, inject_index=None, truncation=1, truncation_latent=None, input_is_latent=False):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        return assemble_latent(styles, self.n_latent, inject_index)

    def forward(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                input_is_latent=False, noise=None, randomize_noise=True, return_features=False):
        latent = self.prepare_latent(styles, inject_index, truncation, truncation_latent, input_is_latent)
        if noise is None:
            noise = ([None] * self.num_layers if randomize_noise
                     else [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)])
        out = self.conv1(self.input(latent), latent[:, 0], noise=noise[0])
        skip = self.to_rgb1(out, latent[:, 1])
        features = [out] if return_features else []
        i = 1
        for up, conv, n_up, n_conv, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2], noise[2::2],
                                                   self.to_rgbs):
            out = up(out, latent[:, i], noise=n_up)
            if return_features:
                features.append(out)
            out = conv(out, latent[:, i + 1], noise=n_conv)
            skip = to_rgb(out, latent[:, i + 2], skip)
            i += 2
        if return_latents:
            return skip, latent
        if return_features:
            return skip, features
        return skip, None
