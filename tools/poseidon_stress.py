#!/usr/bin/env python3
"""One-off stress of the FP64-pipe Poseidon against the CPU oracle: random rows plus rows whose 32-bit halves are forced to
0 / 2^32 - 1 / small values at random (the patterns that move the exact-double accumulators to their extremes), through
the thread-per-permutation kernel (hash_no_pad_many) and, as 8-word leaves of a tree with cap height 0, through the
12-lane kernels.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonky2_aes_b200.host.polynomial_batch import Context
from tests import oracle_lib

P = 0xFFFFFFFF00000001
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = np.random.default_rng(20261018)
rows = rng.integers(0, P, size=(n, 8), dtype=np.uint64)
lo = rows & np.uint64(0xFFFFFFFF)
hi = rows >> np.uint64(32)
sel_lo = rng.integers(0, 6, size=rows.shape)
sel_hi = rng.integers(0, 6, size=rows.shape)
lo = np.where(sel_lo == 0, 0, np.where(sel_lo == 1, 0xFFFFFFFF, np.where(sel_lo == 2, lo & np.uint64(0xFF), lo))).astype(np.uint64)
hi = np.where(sel_hi == 0, 0, np.where(sel_hi == 1, 0xFFFFFFFF, np.where(sel_hi == 2, hi & np.uint64(0xFF), hi))).astype(np.uint64)
forced = (hi << np.uint64(32)) | lo
forced = np.where(forced >= np.uint64(P), forced - np.uint64(P), forced)          # keep canonical inputs
rows[n // 2:] = forced[n // 2:]
orc = oracle_lib.load()
ctx = Context(0)
got = ctx.hash_no_pad_many(rows)
bad = 0
for i in range(n):
    if list(got[i]) != list(orc.hash_no_pad(rows[i])):
        bad += 1
m = 1 << 14
cap, dig = ctx.merkle_cap(rows[:m], 0)
t = orc.merkle(rows[:m], 0)
tree_ok = bool(np.array_equal(cap, t.cap) and np.array_equal(dig, t.level(0)))
t.free()
ctx.close()
print(json.dumps({"rows": n, "forced_half_rows": n - n // 2, "mismatches": bad, "tree_of_16384_leaves_equal": tree_ok}))
sys.exit(0 if bad == 0 and tree_ok else 1)
