"""Drop-in boundary on the CPU (no compute): the helper contracts the reference scripts rely on, and — when the reference
tree is present (/root/reference exists only in the build container, never on the GPU box) — that ITS OWN model / dataset
files import cleanly on top of ``vspbfr_b200.dropin.install()`` and describe the same networks as this package."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_listdir_appends_in_place_like_the_reference(tmp_path):
    """op/utils_train.py:8-26: recursive, every file (no extension filter), caller's list filled and sorted, returns None."""
    from vspbfr_b200.op.utils_train import listdir

    (tmp_path / "b").mkdir()
    for rel in ("b/2.png", "b/1.txt", "a.jpg", "c.JPG"):
        (tmp_path / rel).write_bytes(b"x")
    names = ["zzz"]
    assert listdir(str(tmp_path), names) is None
    assert names == sorted([str(tmp_path / r) for r in ("a.jpg", "b/1.txt", "b/2.png", "c.JPG")] + ["zzz"])
    assert listdir(str(tmp_path)) == sorted(str(tmp_path / r) for r in ("a.jpg", "b/1.txt", "b/2.png", "c.JPG"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_files_import_over_the_dropin(tmp_path):
    """models/RestoreNet.py:6, e4e/models/stylegan2/model.py:7-12, dataset.py:7 import ``op`` — after install() that is this
    package.  Run in a subprocess so the aliased ``op`` modules do not leak into the rest of the suite."""
    for rel in ("lq/x/1.png", "lq/0.jpg", "hq/x/1.png", "hq/0.jpg", "lq/notes.txt"):
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(b"x")
    code = textwrap.dedent(f"""
        import sys, types
        sys.path.insert(0, {ROOT!r})
        from vspbfr_b200 import dropin
        dropin.install()
        sys.path.insert(1, {REF!r})
        m = types.ModuleType("matplotlib"); m.use = lambda *a, **k: None; sys.modules["matplotlib"] = m
        import op
        assert op.__name__ == "vspbfr_b200.op", op.__name__
        from models import RestoreNet as R                     # the reference's own file
        assert R.__file__.startswith({REF!r}), R.__file__
        assert R.conv2d_gradfix.__name__ == "vspbfr_b200.op.conv2d_gradfix"
        assert R.upfirdn2d.__module__ == "vspbfr_b200.op.upfirdn2d"
        from e4e.models.stylegan2 import model as S
        assert S.fused_leaky_relu.__module__ == "vspbfr_b200.op.fused_act"
        import torch
        from vspbfr_b200 import restorenet, stylegan2
        torch.manual_seed(0); a = R.Restoration_net(16, 512, 2)
        torch.manual_seed(0); b = restorenet.Restoration_net(16, 512, 2)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa)       # same construction order -> same seeded init
        torch.manual_seed(1); a = R.Discriminator(16)
        torch.manual_seed(1); b = restorenet.Discriminator(16)
        assert list(a.state_dict()) == list(b.state_dict())
        assert all(torch.equal(v, b.state_dict()[k]) for k, v in a.state_dict().items())
        torch.manual_seed(2); a = S.Generator(16, 512, 2)
        torch.manual_seed(2); b = stylegan2.Generator(16, 512, 2)
        assert list(a.state_dict()) == list(b.state_dict())
        # the reference's layers raise this package's loud error on CPU tensors instead of silently using another path
        try:
            R.Blur([1, 3, 3, 1], pad=(1, 1))(torch.zeros(1, 1, 8, 8))
            raise SystemExit("expected RuntimeError")
        except RuntimeError:
            pass
        import dataset                                          # dataset.py:7 -> op.utils_train.listdir(path, list)
        ds = dataset.ImageFolder_restore_test({str(tmp_path / 'lq')!r}, {str(tmp_path / 'hq')!r}, im_size=(16, 16))
        assert len(ds) == 2 and all(f.endswith(('.png', '.jpg')) for f in ds.lq_frame), ds.lq_frame
        print("ok")
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
