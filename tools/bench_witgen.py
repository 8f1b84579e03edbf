#!/usr/bin/env python3
"""Witness generation rate for BASELINE config 2 (AES-GCM-128, 256 B): host generators (libp2witness, OpenMP)
vs the level-scheduled device program (p2g_wprog_generate), and one proof from input values
(p2g_prove_inputs).  One JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonky2_aes_b200.host.polynomial_batch import Context
from tests import circuits

ctx = Context(0)
data, _, tg = circuits.aes_gcm(256, True)
data.load(ctx)
targets = tg.input_targets()
wp = data.load_witness_program(ctx, targets)
out = {"levels": ctx.lib.p2g_wprog_levels(wp), "ext_slots": data.ext_slots, "inputs": len(targets)}
for count in (1, 8, 64, 256):
    vals = circuits.gcm_inputs(tg, 20261018, count)
    data.generate_slots_device(ctx, wp, vals[:1])
    t0 = time.perf_counter(); dev = data.generate_slots_device(ctx, wp, vals); dt = time.perf_counter() - t0
    out[f"device_{count}"] = {"s": dt, "witnesses_per_s": count / dt, "includes": "H2D of inputs, D2H of the slot vectors"}
    if count == 64:
        for th in (1, 8, 16):
            data._wlib.p2w_set_num_threads(th)
            t0 = time.perf_counter(); host = data.generate_slots_many(targets, vals); dh = time.perf_counter() - t0
            out[f"host_{th}_threads"] = {"witnesses_per_s": count / dh}
        out["equal"] = bool(np.array_equal(dev, host))
vals = circuits.gcm_inputs(tg, 5, 2)
data.prove_inputs(vals[0], wp)
t0 = time.perf_counter(); p = data.prove_inputs(vals[1], wp); out["prove_inputs_one_proof_ms"] = (time.perf_counter() - t0) * 1e3
s = data.generate_slots_many(targets, vals[1:2])[0]
t0 = time.perf_counter(); p2 = data.prove_slots(s); out["prove_slots_one_proof_ms"] = (time.perf_counter() - t0) * 1e3
out["proofs_equal"] = bool(np.array_equal(p, p2))
print(json.dumps(out))
