"""CPU: libp2gpu.so loads and exports every symbol include/p2gpu.h declares (no compute)."""
import ctypes
import os
import re

from plonky2_aes_b200.host import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "p2gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p2g_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(ffi.lib_path())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(ffi.SIGNATURES) == names
    assert lib.p2g_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    from plonky2_aes_b200.host.polynomial_batch import Context
    try:
        Context(0)
    except ffi.P2GError as e:
        assert e.code == -1
    else:
        raise AssertionError("context creation must fail without a GPU")


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "plonky2_aes_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "oracle/" not in txt, f
