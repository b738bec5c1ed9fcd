"""Drop-in for the reference's ``op`` package (/root/reference/op/__init__.py:2-7).

Same public names: ``FusedLeakyReLU``, ``fused_leaky_relu``, ``upfirdn2d`` (the attribute
is the FUNCTION, shadowing the submodule exactly like the reference),
``conv2d_gradfix``, plus the small helpers the reference scripts import.
"""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d
from . import conv2d_gradfix
from .utils import mkdirs, delete_dirs, set_random_seed

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d", "conv2d_gradfix",
           "mkdirs", "delete_dirs", "set_random_seed"]
