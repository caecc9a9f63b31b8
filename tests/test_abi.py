"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/glowk.h declares (no compute without a GPU), argument validation returns error codes instead
of aborting, and the ctypes table mirrors the header."""
import ctypes
import os
import re

import pytest

from pytorch_glow_b200 import _C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_decls():
    src = open(os.path.join(ROOT, "include", "glowk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:const\s+char\s*\*|int64_t|int)\s+(glowk_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_library_exports_every_declared_symbol():
    decls = header_decls()
    assert len(decls) >= 20
    L = _C.lib()
    for name, nargs in decls.items():
        assert hasattr(L, name), "libglowk.so does not export %s" % name
        assert name in _C.SIGNATURES, "%s is missing from the ctypes table" % name
        assert len(_C.SIGNATURES[name]) == nargs, "%s: header has %d args, ctypes table %d" % (
            name, nargs, len(_C.SIGNATURES[name]))
    for name in _C.SIGNATURES:
        assert name in decls, "%s is bound but not declared in include/glowk.h" % name


def test_version_and_device_probe_without_gpu():
    L = _C.lib()
    assert L.glowk_version() >= 100
    assert L.glowk_has_tcgen05() in (0, 1)
    assert L.glowk_coupling_nblk(1024) == 4 and L.glowk_coupling_nblk(16) == 1 and L.glowk_coupling_nblk(64) == 1


def test_argument_validation_returns_codes():
    """Bad arguments come back as GLOWK_EINVAL + message (-> ValueError in Python), never an abort.
    None of these reaches a kernel launch, so they run without a GPU."""
    L = _C.lib()
    one = ctypes.c_void_p(16)
    # odd channel count (network/model.py:169)
    rc = L.glowk_coupling(one, 64, one, one, 3.0, one, one, None, 1, 3, 2, 2, 1, 0, None)
    assert rc == 1 and b"even" in L.glowk_last_error()
    # squeeze: H, W not divisible by the factor (network/module.py:588)
    rc = L.glowk_squeeze2d(one, ctypes.c_void_p(32), 1, 3, 3, 4, 36, 2, 0, None)
    assert rc == 1 and b"divisible" in L.glowk_last_error()
    # unsqueeze: C not divisible by factor^2 (network/module.py:566)
    rc = L.glowk_squeeze2d(one, ctypes.c_void_p(32), 1, 6, 2, 2, 24, 2, 1, None)
    assert rc == 1
    # mix needs exactly one of weight / indices
    rc = L.glowk_actnorm_mix(one, ctypes.c_void_p(32), None, None, None, None, 3.0, 1, 4, 4, 0, None)
    assert rc == 1
    # im2col: only 1x1 / 3x3
    rc = L.glowk_im2col(one, 1, 64, 0, 4, 4, 4, 5, 0, one, 0, 128, None)
    assert rc == 1
    rc = L.glowk_pack_conv_weight(one, 4, 4, 3, 7, one, 0, 4, 36, None)       # bad layout id
    assert rc == 1


def test_empty_batches_are_no_ops():
    L = _C.lib()
    assert L.glowk_actnorm(None, None, None, None, 3.0, 0, 4, 16, 0, None) == 0
    assert L.glowk_gemm(None, 8, None, 8, 0, 0, 8, 8, 0, None, None, 3.0, None, 0, None, None, None, 0, 8, None) == 0
    assert L.glowk_squeeze2d(None, None, 0, 3, 4, 4, 48, 2, 0, None) == 0
