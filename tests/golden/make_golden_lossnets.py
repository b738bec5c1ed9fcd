#!/usr/bin/env python
"""Golden outputs of the reference's loss networks (run in the build container, needs /root/reference):
my_lpips.networks_basic.PNetLin('vgg', pnet_rand=True) and Loss.id_loss.IDLoss on seeded random initialisations and seeded
inputs -> tests/golden/lossnets.npz (state_dict manifests, per-tensor checksums, outputs; inputs are regenerated from the
seeds).  Optional third-party imports of the reference that this image lacks (skimage, IPython, matplotlib) are stubbed —
none of them is on the computed path."""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
for name in ("skimage", "skimage.metrics", "skimage.color", "skimage.transform", "IPython", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["skimage"].__path__ = []
sys.modules["skimage.metrics"].structural_similarity = None
sys.modules["skimage"].color = sys.modules["skimage.color"]
sys.modules["skimage"].transform = sys.modules["skimage.transform"]
sys.modules["IPython"].embed = None
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.path.insert(0, REF)
# Loss/id_loss.py and my_lpips/__init__.py import SVGL (an optional pixel re-weighting used only with a weight map) from
# Loss.e4e_embedding, which drags in the whole e4e encoder and the reference's JIT-compiled op package: stub that module
_stub = types.ModuleType("Loss.e4e_embedding")
_stub.SVGL = None
sys.modules["Loss.e4e_embedding"] = _stub
# the reference asks torchvision for pretrained=False through the pre-0.13 keyword
import torchvision.models as tv  # noqa: E402

_vgg16 = tv.vgg16
tv.vgg16 = lambda pretrained=False, **kw: _vgg16(weights=None)
from my_lpips import networks_basic as nb  # noqa: E402
from Loss import id_loss as ref_id  # noqa: E402


def manifest(sd):
    return np.array([f"{k}|{'x'.join(str(int(s)) for s in v.shape)}" for k, v in sd.items()])


def main():
    out = {}
    torch.manual_seed(31)
    net = nb.PNetLin(pnet_type="vgg", pnet_rand=True, pnet_tune=False, use_dropout=True, spatial=False, version="0.1", lpips=True).eval()
    out["lpips_manifest"] = manifest(net.state_dict())
    out["lpips_sums"] = np.array([float(v.double().sum()) for v in net.state_dict().values()])
    g = torch.Generator().manual_seed(32)
    a = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    b = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    with torch.no_grad():
        out["lpips_out"] = net.forward(b, a).numpy()          # PerceptualLoss.forward(pred=a, target=b) -> model.forward(target, pred)
    torch.manual_seed(33)
    idl = ref_id.IDLoss.__new__(ref_id.IDLoss)                  # the constructor loads a checkpoint and moves to CUDA
    torch.nn.Module.__init__(idl)
    from torchvision.models import resnet101
    idl.Z = resnet101(num_classes=256).eval()
    idl.Z.requires_grad_(False)
    idl.l1 = torch.nn.L1Loss()
    out["id_manifest"] = manifest(idl.Z.state_dict())
    out["id_sums"] = np.array([float(v.double().sum()) for v in idl.Z.state_dict().values()])
    g = torch.Generator().manual_seed(34)
    x = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
    y = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
    with torch.no_grad():
        out["id_loss"] = np.array(float(idl.forward(x, y)))
        out["id_embed"] = idl.get_id(x).numpy()
    np.savez_compressed(os.path.join(OUT, "lossnets.npz"), **out)
    print("lossnets: lpips", out["lpips_out"].ravel(), "id", out["id_loss"])


if __name__ == "__main__":
    main()
