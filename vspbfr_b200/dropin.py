"""Make the reference's own scripts and model files import THIS implementation.

The reference does ``from op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d, conv2d_gradfix``
(models/RestoreNet.py:6), ``from op.fused_act import …`` / ``from op.upfirdn2d import …``
(e4e/models/stylegan2/model.py:8-9), ``from op import conv2d_gradfix`` (restoration_train.py:24) and a few
helpers from ``op.utils`` / ``op.utils_train``.  ``install()`` registers this package's modules under those
names in ``sys.modules`` (before the reference is imported), so its ``models/RestoreNet.py`` and the e4e
decoder pick up the sm_100a operators unchanged; ``install(models=True)`` additionally serves
``models.RestoreNet`` itself from ``vspbfr_b200.restorenet`` / ``layers``.

    python -m vspbfr_b200.dropin /path/to/VSPBFR/restoration_test.py --ckpt …      # runpy launcher
"""
from __future__ import annotations

import importlib
import runpy
import sys
import types


def install(models: bool = False) -> None:
    from . import op
    from .op import conv2d_gradfix, fused_act, utils, utils_train
    from .op import upfirdn2d as _  # noqa: F401  (the attribute is the function; the module lives in sys.modules)

    sys.modules["op"] = op
    sys.modules["op.fused_act"] = fused_act
    sys.modules["op.upfirdn2d"] = sys.modules["vspbfr_b200.op.upfirdn2d"]
    sys.modules["op.conv2d_gradfix"] = conv2d_gradfix
    sys.modules["op.utils"] = utils
    sys.modules["op.utils_train"] = utils_train
    # the e4e decoder imports these names when no GPU is visible (e4e/models/stylegan2/model.py:10-12)
    sys.modules["op.fused_act_cpu"] = fused_act
    sys.modules["op.upfirdn2d_cpu"] = sys.modules["vspbfr_b200.op.upfirdn2d"]
    if models:
        from . import layers, restorenet

        pkg = sys.modules.get("models") or types.ModuleType("models")
        pkg.__path__ = getattr(pkg, "__path__", [])
        mod = types.ModuleType("models.RestoreNet")
        for src in (layers, restorenet):
            for name in dir(src):
                if not name.startswith("_"):
                    setattr(mod, name, getattr(src, name))
        sys.modules["models"] = pkg
        sys.modules["models.RestoreNet"] = mod
        pkg.RestoreNet = mod


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m vspbfr_b200.dropin <reference script.py> [args…]")
    install()
    script = argv[0]
    import os

    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
