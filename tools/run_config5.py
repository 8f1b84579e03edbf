#!/usr/bin/env python3
"""BASELINE config 5: a batch of independent AES-GCM 16-block proofs on the GPU(s) of this process
(one process per GPU under torchrun; here: single GPU), including witness generation on the host
and the restated verifier on a sample.  Also records device memory before/after to show the
prover does not leak across proofs.  Usage: python tools/run_config5.py [num_proofs] [in_flight]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from plonky2_aes_b200.host.polynomial_batch import Context
from plonky2_aes_b200.host.sharding import BatchProver
from tests import circuits, oracle_lib

total = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
USE_SLOTS = os.environ.get("P2G_CONFIG5_WIRES", "0") != "1"   # default: device-side wire fill (p2g_prove_slots)
CHUNK = 32
ctxs = [Context(0) for _ in range(T)]
data, _, tg = circuits.aes_gcm(256, True)
data.load(ctxs[0])
bp = BatchProver(data, ctxs)
orc = oracle_lib.load()
oc = oracle_lib.OracleCircuit(orc, data)
import threading
shape = (CHUNK, data.ext_slots) if USE_SLOTS else (CHUNK, 135, data.n)
bufs = [torch.empty(shape, dtype=torch.int64).pin_memory() for _ in range(2)]
hvs = [b.numpy().view(np.uint64) for b in bufs]
hv = hvs[0]

t_wit = t_prove = 0.0
checked = 0
wit_time = [0.0]


def make_chunk(base, buf):
    t0 = time.perf_counter()
    cnt = min(CHUNK, total - base)
    vals = circuits.gcm_inputs(tg, 20261018 + base, cnt)        # key/nonce/pt from default_rng(20261018 + i)
    if USE_SLOTS:
        data.generate_slots_many(tg.input_targets(), vals, out=buf[:cnt])
    else:
        data.generate_witnesses(tg.input_targets(), vals, out=buf[:cnt])
    wit_time[0] += time.perf_counter() - t0


make_chunk(0, hvs[0])
bp.prove_many([hvs[0][i % min(CHUNK, total)] for i in range(T)], slots=USE_SLOTS)     # warm-up (tables, pools)
torch.cuda.synchronize()
free0, _ = torch.cuda.mem_get_info()
wit_time[0] = 0.0
t_all = time.perf_counter()
make_chunk(0, hvs[0])
for k, base in enumerate(range(0, total, CHUNK)):
    cnt = min(CHUNK, total - base)
    cur = hvs[k % 2]
    th = None
    if base + CHUNK < total:                     # host generates the next chunk's witnesses while the GPU proves
        th = threading.Thread(target=make_chunk, args=(base + CHUNK, hvs[(k + 1) % 2]))
        th.start()
    t1 = time.perf_counter()
    proofs = bp.prove_many([cur[i] for i in range(cnt)], slots=USE_SLOTS)
    t_prove += time.perf_counter() - t1
    if th:
        th.join()
    if base % (8 * CHUNK) == 0:
        assert oc.verify(proofs[0]) == 0
        checked += 1
t_wit = wit_time[0]
wall = time.perf_counter() - t_all
free1, _ = torch.cuda.mem_get_info()
print(json.dumps({"config": "BASELINE configs[4]: batch of independent AES-GCM-128 16-block proofs", "proofs": total, "gpus": 1,
                  "in_flight": T, "wire_fill": "device (p2g_prove_slots)" if USE_SLOTS else "host", "prove_s": t_prove, "proofs_per_s_prove_only": total / t_prove,
                  "witness_generation_s_host": t_wit, "wall_s_witness_overlapped": wall, "proofs_per_s_wall": total / wall,
                  "verified_samples": checked, "device_free_bytes_before": free0, "device_free_bytes_after": free1,
                  "host_threads": os.cpu_count()}))
