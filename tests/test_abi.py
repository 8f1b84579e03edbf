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


def test_rust_bindings_match_header():
    """rust/p2gpu-sys declares exactly what include/p2gpu.h declares: the extern block is generated from the header
    (tools/gen_rust_bindings.py) and the #[repr(C)] structs list the same fields in the same order."""
    from tools import gen_rust_bindings as g
    text, names = g.render()
    assert open(g.OUT).read() == text, "rust/p2gpu-sys/src/ffi.rs is stale: run python tools/gen_rust_bindings.py"
    assert names == [n for n in sorted(set(names), key=names.index)] and sorted(names) == _declared()
    rs = open(os.path.join(ROOT, "rust", "p2gpu-sys", "src", "types.rs")).read()
    hdr = "".join(re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", h)).read(), flags=re.S) for h in ("p2gpu.h", "p2witness.h"))
    for name in ("p2g_gate", "p2g_circuit_desc", "p2g_timings", "p2g_transcript", "p2w_program_desc"):
        body = re.search(r"typedef struct\s*\{([^}]*)\}\s*" + name + r"\s*;", hdr).group(1)
        c_fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *rest = decl.split(",")
            c_fields.append(re.search(r"(\w+)\s*(\[\d+\])?$", first.strip()).group(1))
            c_fields += [re.search(r"(\w+)\s*(\[\d+\])?$", r.strip()).group(1) for r in rest]
        rbody = re.search(r"pub struct " + name + r"\s*\{([^}]*)\}", rs).group(1)
        r_fields = re.findall(r"pub (\w+)\s*:", rbody)
        assert r_fields == c_fields, (name, r_fields, c_fields)
