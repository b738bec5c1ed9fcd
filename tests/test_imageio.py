"""I/O either side of the hot path: device-side quantisation must produce the bytes torchvision's
save_image(normalize=True, range=(-1, 1)) writes (restoration_test.py:138-157), the asynchronous writer must put them on
disk under the reference's names, and the prefetching loader must decode like dataset.py:411-436."""
import os

import numpy as np
import pytest
import torch


def _reference_bytes(x, lo=-1.0, hi=1.0):
    """torchvision.utils.make_grid(normalize=True, value_range=(lo, hi)) + save_image's quantisation, per image."""
    t = x.clone().float().clamp_(min=lo, max=hi)
    t.sub_(lo).div_(max(hi - lo, 1e-5))
    return t.mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8)


def test_load_image_matches_the_reference_decode(tmp_path):
    from PIL import Image
    from torchvision import transforms
    from vspbfr_b200.imageio import load_image

    rng = np.random.default_rng(0)
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5), inplace=True)])
    for (h, w), size in (((64, 64), (64, 64)), ((50, 80), (32, 32)), ((90, 40), (48, 48))):
        p = str(tmp_path / f"{h}x{w}.png")
        Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(p)
        img = Image.open(p).convert("RGB")
        if (h, w) != size:                                   # dataset.py:418-432
            ratio = max(1.0 * size[0] / h, 1.0 * size[1] / w)
            nw, nh = int(ratio * w), int(ratio * h)
            img = img.resize((nw, nh), Image.Resampling.LANCZOS)
            hi = (nh - size[0]) // 2 if nh - size[0] > 0 else 0
            wi = (nw - size[1]) // 2 if nw - size[1] > 0 else 0
            img = img.crop((wi, hi, int(wi + size[1]), int(hi + size[0])))
        want = tf(img).numpy()
        got = load_image(p, size)
        assert got.shape == (3,) + size and got.dtype == np.float32
        np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 3, 64, 64), (2, 3, 33, 47), (1, 1, 16, 16), (5, 3, 512, 512)])
def test_quantize_u8_is_byte_identical_to_save_image(shape):
    from vspbfr_b200.imageio import quantize_u8
    g = torch.Generator().manual_seed(shape[-1])
    x = (torch.randn(shape, generator=g) * 0.8).cuda()
    x.view(-1)[:7] = torch.tensor([-1.0, 1.0, 0.0, -1.5, 2.0, 0.00392, -0.99999], device="cuda")
    got = quantize_u8(x)
    assert torch.equal(got.cpu(), _reference_bytes(x.cpu()))
    got2 = quantize_u8(x, value_range=(0.0, 1.0))
    assert torch.equal(got2.cpu(), _reference_bytes(x.cpu(), 0.0, 1.0))


@pytest.mark.gpu
def test_image_writer_and_prefetch_loader_round_trip(tmp_path):
    """restored batch -> PNG files named as restoration_test.py:141-157 -> PrefetchLoader -> the quantised images again."""
    from PIL import Image
    from vspbfr_b200.imageio import ImageWriter, PrefetchLoader, quantize_u8
    g = torch.Generator().manual_seed(3)
    batches = [(torch.rand(4, 3, 48, 48, generator=g) * 2.4 - 1.2).cuda() for _ in range(5)]
    out = str(tmp_path / "eval")
    with ImageWriter(out, rank=0, name="toy", workers=4, depth=2) as wr:
        for i, b in enumerate(batches):
            wr.save(i * 4, restored=b, low=-b)
    names = sorted(os.listdir(out))
    assert len(names) == 40 and names[0] == "000000_0_toy_low.png" and names[1] == "000000_0_toy_restored.png"
    for i, b in enumerate(batches):
        want = quantize_u8(b).cpu().numpy()
        for j in range(4):
            got = np.asarray(Image.open(os.path.join(out, f"{i * 4 + j:06d}_0_toy_restored.png")).convert("RGB"))
            np.testing.assert_array_equal(got, want[j])
    paths = [os.path.join(out, f"{k:06d}_0_toy_restored.png") for k in range(20)]
    seen = 0
    for first, x in PrefetchLoader(paths, batch=8, size=(48, 48), workers=4, depth=2):
        assert x.is_cuda and x.shape[1:] == (3, 48, 48) and first == seen
        want = torch.cat(batches)[first:first + x.shape[0]]
        ref = quantize_u8(want).permute(0, 3, 1, 2).float() / 255.0 * 2 - 1        # what the PNG holds, as [-1, 1]
        assert float((x - ref).abs().max()) < 1e-6
        seen += x.shape[0]
    assert seen == 20


@pytest.mark.gpu
def test_restore_folder_runs_the_reference_test_loop(tmp_path):
    """restoration_test.py:111-157 without ground truth on a folder of 10 images (one ragged batch), small random-init
    networks: every image produces its restore / low / sample PNG under the reference's names, the `low` file holds the
    decoded input, and the restored files equal a direct call of the pipeline on the same inputs (same seeded z)."""
    from PIL import Image
    from vspbfr_b200 import frontend
    from vspbfr_b200.imageio import list_images, load_image, quantize_u8, restore_folder
    from vspbfr_b200.restorenet import Restoration_net
    from vspbfr_b200.stylegan2 import Generator
    rng = np.random.default_rng(5)
    lq = tmp_path / "lq" / "sub"
    lq.mkdir(parents=True)
    for i in range(10):
        Image.fromarray(rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)).save(str(lq / f"{i:03d}.png"))
    (tmp_path / "lq" / "notes.txt").write_text("x")
    assert len(list_images(str(tmp_path / "lq"))) == 10
    size = 64
    torch.manual_seed(0)
    dev = torch.device("cuda")
    net = Restoration_net(size, 512, 2, channel_multiplier=2).to(dev).eval()
    dec = Generator(2 * size, 512, 2, channel_multiplier=2).to(dev).eval()
    n_latent = dec.n_latent
    front = frontend.WPlusFrontEnd(frontend.Encoder4Editing(50, "ir_se", stylegan_size=2 * size), n_latent=n_latent).to(dev).eval()
    ddpm = frontend.My_DDPM(frontend.Code_diffuser(timesteps=4), timesteps=4, linear_start=0.1, linear_end=0.99).to(dev).eval()
    out = str(tmp_path / "eval")
    n = restore_folder(front, ddpm, dec, net, str(tmp_path / "lq"), out, batch=4, size=size, name="toy", workers=2)
    assert n == 10
    names = sorted(os.listdir(out))
    assert len(names) == 30 and names[:3] == ["000000_0_toy_low.png", "000000_0_toy_restore.png", "000000_0_toy_sample.png"]
    paths = list_images(str(tmp_path / "lq"))
    for i in (0, 5, 9):
        low = torch.from_numpy(load_image(paths[i], (size, size)))[None].cuda()
        got = np.asarray(Image.open(os.path.join(out, f"{i:06d}_0_toy_low.png")).convert("RGB"))
        np.testing.assert_array_equal(got, quantize_u8(low)[0].cpu().numpy())
        r = np.asarray(Image.open(os.path.join(out, f"{i:06d}_0_toy_restore.png")).convert("RGB"))
        assert r.shape == (size, size, 3) and r.std() > 0


@pytest.mark.gpu
def test_restore_from_host_uint8_output_equals_quantised_fp32_output():
    """sharding.restore_from_host with a uint8 [N,S,S,3] host buffer returns exactly quantize_u8 of its fp32 result
    (eager micro-batches, a ragged tail, and the grouped-tail graph restorer)."""
    from vspbfr_b200 import fastpath as fp, sharding
    from vspbfr_b200.imageio import quantize_u8
    from vspbfr_b200.restorenet import Restoration_net
    from vspbfr_b200.stylegan2 import Generator
    torch.manual_seed(2)
    size, micro, n = 64, 4, 10
    net = Restoration_net(size, 512, 2, channel_multiplier=2).cuda().eval()
    dec = Generator(2 * size, 512, 2, channel_multiplier=2).cuda().eval()
    g = torch.Generator().manual_seed(4)
    low = (torch.rand(n, 3, size, size, generator=g) * 2 - 1).pin_memory()
    codes = torch.randn(n, dec.n_latent, 512, generator=g).pin_memory()
    z = torch.randn(n, 512, generator=g).pin_memory()
    out_f = torch.empty(n, 3, size, size).pin_memory()
    out_u = torch.empty(n, size, size, 3, dtype=torch.uint8).pin_memory()
    for restorer in (None, fp.GraphedRestorer(net, dec, micro, n_latent=dec.n_latent, device="cuda", tail_groups=2)):
        sharding.restore_from_host(net, dec, low, codes, z, out_f, micro=micro, device="cuda", restorer=restorer)
        sharding.restore_from_host(net, dec, low, codes, z, out_u, micro=micro, device="cuda", restorer=restorer)
        torch.cuda.synchronize()
        want = quantize_u8(out_f.cuda()).cpu()
        assert torch.equal(out_u, want)
