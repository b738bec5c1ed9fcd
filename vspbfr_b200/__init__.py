"""vspbfr_b200 — B200-native (sm_100a) implementation of VSPBFR's StyleGAN2-style synthesis hot path.

Host side: Python/PyTorch mirroring the reference's operator interface (``vspbfr_b200.op``
== the reference's ``op`` package; ``vspbfr_b200.restorenet`` / ``.stylegan2`` == its layer
classes).  Device side: hand-written CUDA behind the C ABI of ``include/vsp_b200.h``.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"


def build(force: bool = False, verbose: bool = False) -> str:
    return _lib.build(force=force, verbose=verbose)
