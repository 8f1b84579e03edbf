#!/usr/bin/env python3
"""Chained-permutation rate (p2g_poseidon_peak) and a digest of the permutation of fixed inputs, for the
library selected by P2G_LIB_PATH (A/B of kernel variants, tools/build_variant.sh).  One JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonky2_aes_b200.host import ffi
from plonky2_aes_b200.host.polynomial_batch import Context

ctx = Context(0)
rows = np.random.default_rng(1).integers(0, 0xFFFFFFFF00000001, size=(4096, 24), dtype=np.uint64)
rows[0, :] = 0xFFFFFFFF00000000
h = ctx.hash_no_pad_many(rows)
digest = int(np.bitwise_xor.reduce(h.ravel() * np.arange(1, h.size + 1, dtype=np.uint64)))
best = max(ctx.poseidon_peak(64) for _ in range(3))
print(json.dumps({"lib": os.path.basename(ffi.lib_path()), "perms_per_s": best, "hash_digest": hex(digest)}))
ctx.close()
