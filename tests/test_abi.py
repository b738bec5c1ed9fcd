"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol
include/vsp_b200.h declares; no compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest

from vspbfr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vsp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsp_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_table_agree():
    assert _declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    path = _lib.build()
    lib = ctypes.CDLL(path)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.vsp_version() == 1


def test_out_size_helper_matches_reference_formula():
    lib = _lib.load()
    for (n, k, up, down, p0, p1) in [(64, 4, 2, 1, 2, 1), (65, 4, 1, 1, 1, 1), (64, 4, 1, 2, 1, 1),
                                      (10, 12, 2, 1, 6, 5), (24, 12, 1, 2, -1, -1), (3, 4, 1, 1, 0, 0)]:
        assert lib.vsp_upfirdn2d_out_size(n, k, up, down, p0, p1) == (n * up + p0 + p1 - k + down) // down


def test_cpu_tensors_are_rejected_loudly():
    import torch

    from vspbfr_b200 import op

    x = torch.zeros(1, 1, 4, 4)
    k = torch.ones(2, 2)
    with pytest.raises(RuntimeError):
        op.upfirdn2d(x, k)
    with pytest.raises(RuntimeError):
        op.fused_leaky_relu(x, None)
    with pytest.raises(RuntimeError):
        op.conv2d_gradfix.conv2d(x, torch.zeros(1, 1, 3, 3))


def test_op_surface_names():
    from vspbfr_b200 import op

    for name in ("FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d", "conv2d_gradfix"):
        assert hasattr(op, name)
    assert callable(op.upfirdn2d)  # attribute is the function, like the reference (op/__init__.py:4)
    import sys

    assert "vspbfr_b200.op.upfirdn2d" in sys.modules and "vspbfr_b200.op.fused_act" in sys.modules
    for name in ("conv2d", "conv_transpose2d", "no_weight_gradients", "enabled", "weight_gradients_disabled"):
        assert hasattr(op.conv2d_gradfix, name)
    m = op.FusedLeakyReLU(8)
    assert list(m.state_dict().keys()) == ["bias"]


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the ABI structs (vsp_linear_desc, vsp_conv_epilogue) must have the size and field offsets a C
    compiler gives the declarations in include/vsp_b200.h — a field added on one side only would silently shift every
    later field (the header is what a cgo / JNI / ctypes binding of the reference would compile against)."""
    import ctypes
    import shutil
    import subprocess

    from vspbfr_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    structs = {"vsp_linear_desc": _lib.LinearDesc, "vsp_conv_epilogue": _lib.ConvEpilogue}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vsp_b200.h"', "int main(void) {"]
    for cname, mirror in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    include = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run([gcc, "-I", include, str(src), "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.splitlines():
        cname, field, value = line.split()
        mirror = structs[cname]
        want = ctypes.sizeof(mirror) if field == "size" else getattr(mirror, field).offset
        assert int(value) == want, f"{cname}.{field}: header {value} vs ctypes {want}"
