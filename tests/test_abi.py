"""CPU: the C-ABI library builds, loads and exports every symbol include/basi_b200.h declares."""
import ctypes
import os
import re

from basi_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "basi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(basi_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_header_symbols():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    assert lib.basi_version() >= 100


def test_fp16_build_exports_the_same_symbols():
    lib = _lib.load("f16")
    assert lib.basi_half_format() == 1 and _lib.load("bf16").basi_half_format() == 0
    for s in header_symbols():
        assert hasattr(lib, s), "missing export %s in the fp16 build" % s


def test_binding_table_matches_header():
    assert sorted(_lib.exported_symbols()) == header_symbols()


def test_bad_arguments_fail_loudly_without_gpu():
    lib = _lib.load()
    rc = lib.basi_bn_finalize(None, None, None, ctypes.c_double(1.0), ctypes.c_float(1e-5), None, 0, None)
    assert rc == -1
    assert b"bn_finalize" in lib.basi_last_error()


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.Tensor) == 8 + 6 * 4
    assert ctypes.sizeof(_lib.ConvDesc) == 7 * 4
