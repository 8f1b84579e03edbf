#!/usr/bin/env python3
"""LDE+Merkle of ONE wires-shaped commitment (135 columns, n = 2^15) split by coset over the ranks
of a torchrun job (NCCL all-gather of cap entries).  Reports aggregate LDE+Merkle GB/s.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_sharded_commit.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from plonky2_aes_b200.host.polynomial_batch import Context
from plonky2_aes_b200.host.sharding import sharded_commit

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
P = 0xFFFFFFFF00000001
ncols, log_n = 135, 15
cols = np.random.default_rng(5).integers(0, P, size=(ncols, 1 << log_n), dtype=np.uint64)   # same on every rank
dev = torch.from_numpy(cols.view(np.int64)).cuda()
st = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
caps, times = None, []
for it in range(6):
    torch.cuda.synchronize(); ctx.sync()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    caps, h = sharded_commit(ctx, dev, ncols, log_n, rank, world, dist if world > 1 else None)
    e1.record(st)
    torch.cuda.synchronize(); ctx.sync()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    times.append(ms)
    ctx.check(ctx.lib.p2g_batch_free(ctx.handle, h))
if rank == 0:
    ms = min(times[2:])
    n = 1 << log_n
    full = None
    if world > 1:      # the gathered cap must equal the single-GPU commitment
        from plonky2_aes_b200.host.polynomial_batch import PolynomialBatch
        b = PolynomialBatch.from_values_device(ctx, dev.data_ptr(), ncols, log_n)
        full = bool(np.array_equal(b.cap, caps)); b.free()
    print(json.dumps({"what": "one 135 x 2^15 commitment, coset-sharded", "n_gpus": world, "ms": ms,
                      "lde_merkle_GBs_aggregate": (136 * ncols + 768) * n / ms / 1e6, "cap_equals_single_gpu": full}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
