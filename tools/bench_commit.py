"""Kernel-level timing of PolynomialBatch::from_values on device-resident columns."""
import ctypes as C
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from plonky2_aes_b200.host.polynomial_batch import Context, PolynomialBatch

P = 0xFFFFFFFF00000001
ctx = Context(0)
print("poseidon peak perms/s", ctx.poseidon_peak(64))
for ncols, log_n in [(135, 13), (135, 15), (34, 15), (16, 15), (2, 15)]:
    n = 1 << log_n
    host = np.random.default_rng(1).integers(0, P, size=(ncols, n), dtype=np.uint64)
    dev = torch.from_numpy(host.view(np.int64)).cuda()
    torch.cuda.synchronize()
    st = torch.cuda.ExternalStream(ctx.stream)
    times = []
    for it in range(6):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        b = PolynomialBatch.from_values_device(ctx, dev, ncols, log_n)
        e1.record(st)
        ctx.sync(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        b.free()
    t = min(times[2:])
    bytes_ = (136 * ncols + 768) * n
    perms = (8 * n) * ((ncols + 7) // 8) + 8 * n - 16
    print(json.dumps({"ncols": ncols, "log_n": log_n, "ms": t, "lde_merkle_GBs": bytes_ / t / 1e6, "perms_per_s": perms / t * 1e3}))
