"""Loss networks of the training step: LPIPS-VGG perceptual loss and the ArcFace-ResNet identity loss of
restoration_train.py:116,143,236-245.  Plain PyTorch (frozen feature extractors evaluated by the library's convolutions —
they are outside the synthesis hot path), with the reference's module structure so that its checkpoints load unchanged:

* ``PerceptualLoss(net="vgg")``  — my_lpips/__init__.py:19-50 -> dist_model.DistModel -> networks_basic.PNetLin
  (:27-92: ScalingLayer, the five VGG-16 feature slices of pretrained_networks.py:97-135, channel-normalised squared
  differences, one 1x1 ``NetLinLayer`` per slice, spatial mean, sum over slices); ``forward(pred, target)`` -> [N,1,1,1].
* ``IDLoss``  — Loss/id_loss.py:7-47: torchvision ``resnet101(num_classes=256)`` on the bilinearly 112x112-resized images,
  L2-normalised embeddings, loss = L1(1, <z_source, z_target>).

The pretrained weights (torchvision VGG-16, my_lpips/weights/v0.1/vgg.pth, the ArcFace ResNet-101) cannot be downloaded in
this environment: both classes take an optional checkpoint path and are otherwise randomly initialised — the training-step
benchmark measures their compute, parity is pinned on identically seeded reference classes (tests/test_lossnets_cpu.py).
"""
from __future__ import annotations

from collections import namedtuple

import torch
import torch.nn.functional as F
from torch import nn


def normalize_tensor(in_feat, eps=1e-10):
    """my_lpips/__init__.py:52-54."""
    norm_factor = torch.sqrt(torch.sum(in_feat ** 2, dim=1, keepdim=True))
    return in_feat / (norm_factor + eps)


def spatial_average(in_tens, keepdim=True):
    return in_tens.mean([2, 3], keepdim=keepdim)


class ScalingLayer(nn.Module):
    """networks_basic.py:94-101."""

    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.Tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, inp):
        return (inp - self.shift) / self.scale


class NetLinLayer(nn.Module):
    """networks_basic.py:102-110: (dropout +) 1x1 convolution without bias."""

    def __init__(self, chn_in, chn_out=1, use_dropout=False):
        super().__init__()
        layers = [nn.Dropout()] if use_dropout else []
        layers += [nn.Conv2d(chn_in, chn_out, 1, stride=1, padding=0, bias=False)]
        self.model = nn.Sequential(*layers)


class vgg16(nn.Module):
    """pretrained_networks.py:97-135: torchvision VGG-16 features cut after relu1_2 / 2_2 / 3_3 / 4_3 / 5_3."""

    def __init__(self, requires_grad=False, pretrained=False):
        super().__init__()
        from torchvision import models as tv

        feats = tv.vgg16(weights="IMAGENET1K_V1" if pretrained else None).features
        self.slice1, self.slice2, self.slice3 = nn.Sequential(), nn.Sequential(), nn.Sequential()
        self.slice4, self.slice5 = nn.Sequential(), nn.Sequential()
        self.N_slices = 5
        for lo, hi, sl in ((0, 4, self.slice1), (4, 9, self.slice2), (9, 16, self.slice3), (16, 23, self.slice4),
                           (23, 30, self.slice5)):
            for x in range(lo, hi):
                sl.add_module(str(x), feats[x])
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, x):
        outs = []
        for sl in (self.slice1, self.slice2, self.slice3, self.slice4, self.slice5):
            x = sl(x)
            outs.append(x)
        return namedtuple("VggOutputs", ["relu1_2", "relu2_2", "relu3_3", "relu4_3", "relu5_3"])(*outs)


class PNetLin(nn.Module):
    """networks_basic.py:27-92 for ``pnet_type='vgg'``, ``lpips=True``, ``spatial=False``, version 0.1."""

    def __init__(self, pnet_rand=True, pnet_tune=False, use_dropout=True):
        super().__init__()
        self.scaling_layer = ScalingLayer()
        self.chns = [64, 128, 256, 512, 512]
        self.L = len(self.chns)
        self.net = vgg16(pretrained=not pnet_rand, requires_grad=pnet_tune)
        self.lin0 = NetLinLayer(self.chns[0], use_dropout=use_dropout)
        self.lin1 = NetLinLayer(self.chns[1], use_dropout=use_dropout)
        self.lin2 = NetLinLayer(self.chns[2], use_dropout=use_dropout)
        self.lin3 = NetLinLayer(self.chns[3], use_dropout=use_dropout)
        self.lin4 = NetLinLayer(self.chns[4], use_dropout=use_dropout)
        self.lins = [self.lin0, self.lin1, self.lin2, self.lin3, self.lin4]

    def forward(self, in0, in1):
        outs0, outs1 = self.net(self.scaling_layer(in0)), self.net(self.scaling_layer(in1))
        val = None
        for kk in range(self.L):
            diff = (normalize_tensor(outs0[kk]) - normalize_tensor(outs1[kk])) ** 2
            res = spatial_average(self.lins[kk].model(diff), keepdim=True)
            val = res if val is None else val + res
        return val


class PerceptualLoss(nn.Module):
    """my_lpips.PerceptualLoss(model="net-lin", net="vgg") as restoration_train.py:143 builds it: ``forward(pred, target)``
    returns the LPIPS distance per image, [N,1,1,1] (the caller sums and weights it, :237)."""

    def __init__(self, lin_weights_path=None, pnet_rand=True):
        super().__init__()
        self.model = PNetLin(pnet_rand=pnet_rand, pnet_tune=False, use_dropout=True)
        if lin_weights_path is not None:          # my_lpips/weights/v0.1/vgg.pth (dist_model.py: load_state_dict(..., strict=False))
            self.model.load_state_dict(torch.load(lin_weights_path, map_location="cpu"), strict=False)
        self.model.eval()                         # dist_model.py: self.net.eval()
        for p in self.model.parameters():
            p.requires_grad = False

    def train(self, mode=True):                   # the metric stays in eval mode (no dropout) whatever the owner does
        return super().train(False)

    def forward(self, pred, target, normalize=False):
        if normalize:
            target, pred = 2 * target - 1, 2 * pred - 1
        return self.model(target, pred)


class IDLoss(nn.Module):
    """Loss/id_loss.py:7-47 (without the optional pixel re-weighting map)."""

    def __init__(self, model_path=None):
        super().__init__()
        from torchvision.models import resnet101

        self.Z = resnet101(num_classes=256).eval()
        self.Z.requires_grad_(False)
        if model_path is not None:
            self.Z.load_state_dict(torch.load(model_path, map_location="cpu"))
        self.l1 = nn.L1Loss()

    def train(self, mode=True):
        return super().train(False)

    def id_loss(self, z_id_x, z_id_y):
        inner_product = torch.bmm(z_id_x.unsqueeze(1), z_id_y.unsqueeze(2)).squeeze()
        return self.l1(torch.ones_like(inner_product), inner_product)

    def get_id(self, target_img):
        return F.normalize(self.Z(F.interpolate(target_img, size=112, mode="bilinear")))

    def forward(self, target_img, source_img):
        z_id = self.get_id(source_img).detach()
        return self.id_loss(z_id, self.get_id(target_img))
