"""iNTT / coset-LDE / Merkle times of PolynomialBatch::from_values by size (p2g_last_commit_timings), for A/B runs of
the NTT shape knobs (P2G_NTT_PREFOLD, P2G_NTT_LOG_M).  One JSON line per size."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from plonky2_aes_b200.host.polynomial_batch import Context, PolynomialBatch

P = 0xFFFFFFFF00000001
ctx = Context(0)
cases = [(135, 15), (135, 16), (64, 17), (32, 18), (8, 20)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for ncols, log_n in cases:
    n = 1 << log_n
    host = np.random.default_rng(1).integers(0, P, size=(ncols, n), dtype=np.uint64)
    dev = torch.from_numpy(host.view(np.int64)).cuda()
    torch.cuda.synchronize()
    ctx.check(ctx.lib.p2g_set_timing(ctx.handle, 1))
    times = []
    for it in range(5):
        b = PolynomialBatch.from_values_device(ctx, dev.data_ptr(), ncols, log_n)
        cm = (C.c_float * 3)()
        ctx.check(ctx.lib.p2g_last_commit_timings(ctx.handle, C.byref(cm)))
        times.append(list(cm))
        b.free()
    ctx.check(ctx.lib.p2g_set_timing(ctx.handle, 0))
    intt, lde, merkle = [min(t[i] for t in times[1:]) for i in range(3)]
    print(json.dumps({"ncols": ncols, "log_n": log_n, "prefold": os.environ.get("P2G_NTT_PREFOLD", ""), "log_m": os.environ.get("P2G_NTT_LOG_M", ""),
                      "intt_ms": round(intt, 4), "lde_ms": round(lde, 4), "merkle_ms": round(merkle, 4),
                      "lde_GBs": round(72 * ncols * n / lde / 1e6, 1)}))
    del dev
ctx.close()
