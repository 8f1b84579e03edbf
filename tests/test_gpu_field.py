"""GPU parity of the device field arithmetic (csrc/gl64.cuh: carry-chain PTX add / sub / canon, the
32-bit-half multiply with its 2^64 / 2^96 folds, the shift-multiply of the last NTT pass) against
Python integers, on every pair of carry / borrow boundary values plus random ones.  Restates
plonky2 field/src/goldilocks_field.rs (Add, Sub, Mul, to_canonical_u64) of the dependency pinned at
/root/reference/Cargo.toml:12."""
import ctypes as C
import itertools

import numpy as np
import pytest

P = 0xFFFFFFFF00000001
EPS = 0xFFFFFFFF
M64 = (1 << 64) - 1
pytestmark = pytest.mark.gpu

CANON_EDGES = [0, 1, 2, 7, EPS - 1, EPS, EPS + 1, 1 << 32, (1 << 32) + 1, (1 << 33) - 1, 1 << 63, (1 << 63) - 1,
               0xFFFFFFFE00000000, 0xFFFFFFFEFFFFFFFF, 0xFFFFFFFF00000000, P - 2, P - 1,
               0x00000001FFFFFFFF, 0x7FFFFFFF80000000, 0x80000000FFFFFFFF, 0xFFFFFFFE00000001, 0xFFFFFFFE00000002]
LAZY_EDGES = CANON_EDGES + [P, P + 1, P + 2, M64, M64 - 1, 0xFFFFFFFF80000000, 0xFFFFFFFFFFFF0000, 0xFFFFFFFF00000002]


def _run(gpu_ctx, a, b, la, lb):
    n = len(a)
    arrs = [np.ascontiguousarray(np.array(x, dtype=np.uint64)) for x in (a, b, la, lb)]
    out = np.empty((6, n), dtype=np.uint64)
    gpu_ctx.check(gpu_ctx.lib.p2g_field_ops(gpu_ctx.handle, *[x.ctypes.data_as(C.c_void_p) for x in arrs], n,
                                            out.ctypes.data_as(C.c_void_p)))
    return out


def _check(gpu_ctx, a, b, la, lb):
    out = _run(gpu_ctx, a, b, la, lb)
    n = len(a)
    exp = np.empty((6, n), dtype=np.uint64)
    for i in range(n):
        exp[0, i] = (a[i] + b[i]) % P
        exp[1, i] = (a[i] - b[i]) % P
        exp[2, i] = (a[i] * b[i]) % P
        exp[3, i] = la[i] % P
        exp[4, i] = (la[i] * lb[i]) % P
        exp[5, i] = (a[i] << (i % 96)) % P
    names = ["add", "sub", "mul", "canon", "mul_lazy", "mul_pow2"]
    for k in range(6):
        bad = np.nonzero(out[k] != exp[k])[0]
        assert bad.size == 0, (names[k], [(hex(a[i]), hex(b[i]), hex(la[i]), hex(lb[i]), int(i % 96), hex(int(out[k, i])),
                                            hex(int(exp[k, i]))) for i in bad[:4]])


def test_field_ops_boundaries(gpu_ctx):
    pairs = list(itertools.product(CANON_EDGES, CANON_EDGES))
    lpairs = list(itertools.product(LAZY_EDGES, LAZY_EDGES))
    n1 = max(len(pairs), len(lpairs))
    n1 = ((n1 + 95) // 96) * 96
    n2 = 96 * len(CANON_EDGES)                      # every boundary value against every shift (index mod 96)
    n = n1 + n2
    a = [pairs[i % len(pairs)][0] if i < n1 else CANON_EDGES[(i - n1) // 96] for i in range(n)]
    b = [pairs[i % len(pairs)][1] for i in range(n)]
    la = [lpairs[i % len(lpairs)][0] for i in range(n)]
    lb = [lpairs[i % len(lpairs)][1] for i in range(n)]
    _check(gpu_ctx, a, b, la, lb)


def test_field_ops_random(gpu_ctx):
    rng = np.random.default_rng(7)
    n = 96 * 2048
    a = [int(x) % P for x in rng.integers(0, 1 << 64, size=n, dtype=np.uint64)]
    b = [int(x) % P for x in rng.integers(0, 1 << 64, size=n, dtype=np.uint64)]
    # lazy inputs biased towards the non-canonical band [p, 2^64)
    la = [int(x) for x in rng.integers(0, 1 << 64, size=n, dtype=np.uint64)]
    lb = [int(x) | (0xFFFFFFFF00000000 if i % 3 == 0 else 0) for i, x in enumerate(rng.integers(0, 1 << 64, size=n, dtype=np.uint64))]
    # small differences / sums around the modulus
    for i in range(0, n, 5):
        b[i] = (a[i] + int(rng.integers(-3, 4))) % P
    for i in range(1, n, 7):
        b[i] = (P - a[i] + int(rng.integers(-3, 4))) % P
    _check(gpu_ctx, a, b, la, lb)
