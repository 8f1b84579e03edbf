"""Vectorised Goldilocks arithmetic on numpy uint64 arrays (host-side preprocessing only:
sigma polynomials, subgroup tables).  Canonical inputs, canonical outputs."""
import numpy as np

P = 0xFFFFFFFF00000001
_M32 = np.uint64(0xFFFFFFFF)
_EPS = np.uint64(0xFFFFFFFF)
_P = np.uint64(P)
_S32 = np.uint64(32)


def gl_mul(a, b):
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    with np.errstate(over="ignore"):
        a0, a1, b0, b1 = a & _M32, a >> _S32, b & _M32, b >> _S32
        p00, p01, p10, p11 = a0 * b0, a0 * b1, a1 * b0, a1 * b1
        mid = (p01 & _M32) + (p10 & _M32) + (p00 >> _S32)
        lo = (p00 & _M32) | ((mid & _M32) << _S32)
        hi = p11 + (p01 >> _S32) + (p10 >> _S32) + (mid >> _S32)
        hi_hi, hi_lo = hi >> _S32, hi & _M32
        t0 = lo - hi_hi
        t0 = np.where(lo < hi_hi, t0 - _EPS, t0)
        t1 = hi_lo * _EPS
        t2 = t0 + t1
        t2 = np.where(t2 < t1, t2 + _EPS, t2)
        return np.where(t2 >= _P, t2 - _P, t2)


def gl_add(a, b):
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    with np.errstate(over="ignore"):
        s = a + b
        return np.where((s < a) | (s >= _P), s - _P, s)


def powers(base, count):
    """[base^0, ..., base^(count-1)] by doubling (log2(count) vector multiplies)."""
    out = np.ones(count, dtype=np.uint64)
    if count > 1:
        out[1] = base
    have = 2
    while have < count:
        step = int(out[have - 1]) * base % P        # base^have
        take = min(have, count - have)
        out[have:have + take] = gl_mul(out[:take], np.uint64(step))
        have += take
    return out
