"""Oracle for the fused bias + activation op (test infrastructure).

Follows /root/reference/op/fused_bias_act_kernel.cu:18-65 (element formula and the
``act*10+grad`` switch), op/fused_act.py:126-196 (first/second-order backward)
and op/fused_act.py:216-233 (public function; note its CPU branch ignores
``negative_slope`` and hard-codes 0.2, :222/:228 — reproduced by
``fused_leaky_relu_ref(..., cpu_quirk=True)``).
"""
from __future__ import annotations

import numpy as np


def bias_act_ref(x, b=None, ref=None, act=3, grad=0, alpha=0.2, scale=2 ** 0.5):
    """y[i] = act(x[i] + b[(i // step_b) % size_b]) * scale; bias broadcasts on dim 1."""
    x = np.asarray(x)
    v = x.astype(x.dtype, copy=True)
    if b is not None and np.size(b):
        b = np.asarray(b, dtype=x.dtype)
        shape = [1] * x.ndim
        shape[1] = b.shape[0]
        v = v + b.reshape(shape)
    mode = act * 10 + grad
    if mode == 30:
        y = np.where(v > 0, v, v * x.dtype.type(alpha))
    elif mode == 31:
        y = np.where(np.asarray(ref) > 0, v, v * x.dtype.type(alpha))
    elif mode in (32, 12):
        y = np.zeros_like(v)
    else:  # 10, 11, default
        y = v
    return (y * x.dtype.type(scale)).astype(x.dtype)


def fused_leaky_relu_ref(x, bias=None, negative_slope=0.2, scale=2 ** 0.5, cpu_quirk=False):
    """op/fused_act.py:216-233. ``cpu_quirk`` reproduces the hard-coded 0.2 of the CPU branch."""
    slope = 0.2 if cpu_quirk else negative_slope
    return bias_act_ref(x, bias, None, 3, 0, slope, scale)


def fused_leaky_relu_grads_ref(grad_out, out, has_bias, negative_slope=0.2, scale=2 ** 0.5):
    """First-order backward, op/fused_act.py:126-150: gate on the sign of the saved OUTPUT."""
    grad_out = np.asarray(grad_out)
    dx = bias_act_ref(grad_out, None, out, 3, 1, negative_slope, scale)
    dbias = None
    if has_bias:
        axes = (0,) + tuple(range(2, dx.ndim))
        dbias = dx.sum(axis=axes)
    return dx, dbias
