#!/usr/bin/env python3
"""Chained-permutation rate of the thread-per-permutation Poseidon in its three scheduling variants
(p2g_poseidon_peak_mode): plain, phase-paired warps (w, w+4) and the control pairing (w, w+1).
Prints one JSON line per mode; the checksums must agree."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonky2_aes_b200.host.polynomial_batch import Context

ctx = Context(0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sums = []
for mode, name in ((0, "plain"), (1, "paired (w, w+4): same SM sub-partition"), (2, "paired (w, w+1): control"), (0, "plain again")):
    best = 0.0
    for _ in range(3):
        v, cs = C.c_double(), C.c_uint64()
        ctx.check(ctx.lib.p2g_poseidon_peak_mode(ctx.handle, iters, mode, C.byref(v), C.byref(cs)))
        best = max(best, v.value)
    sums.append(cs.value)
    print(json.dumps({"mode": mode, "name": name, "iters": iters, "perms_per_s": best, "checksum": hex(cs.value)}))
assert len(set(sums)) == 1, "results differ between modes"
ctx.close()
