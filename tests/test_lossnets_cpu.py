"""Loss networks of the training step (vspbfr_b200/lossnets.py) against the unmodified reference's classes
(tests/golden/lossnets.npz from tests/golden/make_golden_lossnets.py): same seeded construction -> identical ``state_dict``
layout and parameters -> outputs equal to float rounding, on the CPU in fp32."""
import numpy as np
import torch

from conftest import load_golden
from vspbfr_b200 import lossnets

G = load_golden("lossnets")


def _manifest(sd):
    return [f"{k}|{'x'.join(str(int(s)) for s in v.shape)}" for k, v in sd.items()]


def test_lpips_vgg_matches_reference():
    torch.manual_seed(31)
    net = lossnets.PNetLin(pnet_rand=True, pnet_tune=False, use_dropout=True).eval()
    assert _manifest(net.state_dict()) == [str(s) for s in G["lpips_manifest"]]
    np.testing.assert_allclose([float(v.double().sum()) for v in net.state_dict().values()], G["lpips_sums"], rtol=1e-9, atol=1e-9)
    g = torch.Generator().manual_seed(32)
    a = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    b = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    with torch.no_grad():
        got = net(b, a)
    np.testing.assert_allclose(got.numpy(), G["lpips_out"], rtol=1e-4, atol=1e-7)
    # the wrapper restoration_train.py:143 builds: pred first, eval mode whatever the owner's mode, frozen
    torch.manual_seed(31)
    loss = lossnets.PerceptualLoss()
    loss.train()
    assert not loss.model.training and not any(p.requires_grad for p in loss.parameters())
    a.requires_grad_(True)
    val = loss(a, b)
    np.testing.assert_allclose(val.detach().numpy(), G["lpips_out"], rtol=1e-4, atol=1e-7)
    val.sum().backward()
    assert a.grad is not None and torch.isfinite(a.grad).all() and float(a.grad.abs().max()) > 0


def test_id_loss_matches_reference():
    torch.manual_seed(33)
    idl = lossnets.IDLoss()
    assert _manifest(idl.Z.state_dict()) == [str(s) for s in G["id_manifest"]]
    np.testing.assert_allclose([float(v.double().sum()) for v in idl.Z.state_dict().values()], G["id_sums"], rtol=1e-9, atol=1e-9)
    g = torch.Generator().manual_seed(34)
    x = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
    y = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
    with torch.no_grad():
        np.testing.assert_allclose(idl.get_id(x).numpy(), G["id_embed"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(float(idl(x, y)), float(G["id_loss"]), rtol=1e-3, atol=1e-7)
    x.requires_grad_(True)
    idl(x, y).backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
