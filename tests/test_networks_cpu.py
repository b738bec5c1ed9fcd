"""CPU-side drop-in contract: state_dict keys/shapes of the rebuilt networks equal the reference's
(tests/golden/state_dict_manifest.npz, generated from the reference), and seeded construction
reproduces the reference's random initialisation (per-tensor checksums in networks.npz)."""
import numpy as np
import torch

from conftest import load_golden
from vspbfr_b200.restorenet import Discriminator, Restoration_net
from vspbfr_b200.stylegan2 import Generator

MAN = load_golden("state_dict_manifest")
NET = load_golden("networks")


def _check(mod, prefix):
    sd = mod.state_dict()
    assert list(sd.keys()) == [str(k) for k in MAN[f"{prefix}.keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in MAN[f"{prefix}.shapes"]]


def test_restoration_net_state_dict_layout():
    with torch.device("meta"):
        net = Restoration_net(512, 512, 8, channel_multiplier=2)
    _check(net, "restoration_net")


def test_generator_state_dict_layout():
    with torch.device("meta"):
        gen = Generator(1024, 512, 8, channel_multiplier=2)
    _check(gen, "generator")


def test_discriminator_state_dict_layout():
    with torch.device("meta"):
        disc = Discriminator(512)
    _check(disc, "discriminator")


def test_seeded_init_matches_reference():
    size = int(NET["size"])
    torch.manual_seed(2024)
    net = Restoration_net(size, 512, 2, channel_multiplier=2)
    dec = Generator(size, 512, 2, channel_multiplier=2)
    for prefix, mod in (("net", net), ("dec", dec)):
        sd = mod.state_dict()
        assert list(sd.keys()) == [str(k) for k in NET[f"{prefix}.keys"]]
        sums = np.array([float(v.double().sum()) for v in sd.values()])
        np.testing.assert_allclose(sums, NET[f"{prefix}.sums"], rtol=1e-6, atol=1e-4)
