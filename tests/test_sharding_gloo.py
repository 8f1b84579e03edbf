"""CPU, world_size 2 over gloo: the N>1 host path (proof sharding, gather to rank 0, max timing).
The GPU prover is replaced by the CPU oracle here — only the distribution logic is under test."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from plonky2_aes_b200.host.sharding import gather_to_rank0, proof_checksum, shard_indices
    from tests import circuits, oracle_lib
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.load()
    data, wires = circuits.tiny_arith()
    oc = oracle_lib.OracleCircuit(orc, data)
    mine = shard_indices(total, rank, world)
    local = []
    for i in mine:
        w = wires.copy()
        local.append((i, oc.prove(w)))
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the max-over-ranks timing reduction
    gathered = gather_to_rank0(local, dist, rank, world)
    if rank == 0:
        ok = sorted(gathered) == list(range(total)) and all(oc.verify(p) == 0 for p in gathered.values())
        sums = {proof_checksum(p) for p in gathered.values()}
        ret["ok"] = bool(ok and float(t.item()) == world and len(sums) == 1)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_proof_sharding():
    from plonky2_aes_b200.host.sharding import shard_indices
    assert shard_indices(5, 0, 2) == [0, 2, 4] and shard_indices(5, 1, 2) == [1, 3]
    assert sorted(shard_indices(1024, 3, 8) + shard_indices(1024, 5, 8))[:4] == [3, 5, 11, 13]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29517, 5, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
