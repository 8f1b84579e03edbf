"""Multi-GPU work split for batches of independent proofs (BASELINE config 5).

The path shards by proof: rank r proves proofs r, r + world, r + 2*world, ... of the batch with
its own GPU context(s); there is no data-path collective.  The only communication is the
gather of results (proof words or their digests) to rank 0 and the max-over-ranks timing, both
through torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import threading

import numpy as np


def shard_indices(total, rank, world):
    """proof indices owned by `rank` (round-robin so every rank gets the same count +-1)"""
    return list(range(rank, total, world))


def proof_checksum(proof_words):
    """order-independent 64-bit fingerprint used to cross-check gathered proofs"""
    a = np.ascontiguousarray(proof_words, dtype=np.uint64)
    w = np.arange(1, a.size + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        return int(np.bitwise_xor.reduce(a * w))


def gather_to_rank0(local, dist, rank, world):
    """gather a list of (index, uint64 array) pairs on rank 0 -> dict index -> array"""
    import torch
    if world == 1:
        return dict(local)
    n_local = len(local)
    width = max((len(p) for _, p in local), default=0)
    meta = torch.tensor([n_local, width], dtype=torch.int64)
    metas = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(metas, meta)
    n_max = max(int(m[0]) for m in metas)
    w_max = max(int(m[1]) for m in metas)
    buf = torch.zeros((n_max, w_max + 2), dtype=torch.int64)
    for k, (idx, p) in enumerate(local):
        buf[k, 0], buf[k, 1] = idx, len(p)
        buf[k, 2:2 + len(p)] = torch.from_numpy(np.ascontiguousarray(p, dtype=np.uint64).view(np.int64))
    bufs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, bufs, dst=0)
    if rank != 0:
        return None
    out = {}
    for r in range(world):
        for k in range(int(metas[r][0])):
            idx, ln = int(bufs[r][k, 0]), int(bufs[r][k, 1])
            out[idx] = bufs[r][k, 2:2 + ln].numpy().view(np.uint64).copy()
    return out


class BatchProver:
    """Proves many witnesses of one circuit with several proofs in flight on one GPU: one p2g
    context (stream + host thread) per in-flight proof, so the latency-bound tails of one proof
    (tree tops, Fiat-Shamir round trips) overlap the heavy kernels of another."""

    def __init__(self, data, ctxs):
        self.data, self.ctxs = data, ctxs
        self.handles = [data._gpu_circuit if c is data.ctx else data.load_handle(c) for c in ctxs]

    def prove_many(self, wires_list, device_resident=False, slots=False):
        """wires_list: wire matrices [135][n] (host, or device with device_resident=True), or -- with
        slots=True -- slot vectors from CircuitData.generate_slots*, gathered into wires on the GPU."""
        import ctypes as C
        lib = self.ctxs[0].lib
        words = self.data.proof_words
        out = [None] * len(wires_list)
        if slots:
            if getattr(self, "wmap", None) is None:      # device-global, shared by all contexts of this GPU
                self.wmap = self.data.load_wire_map(self.ctxs[0], self.handles[0])
            wmap = self.wmap

            def fn(ctx_h, circ_h, ptr, pi, buf, cap, got):
                return lib.p2g_prove_slots(ctx_h, circ_h, wmap, ptr, pi, buf, cap, got)
        else:
            fn = lib.p2g_prove_dev if device_resident else lib.p2g_prove
        errors = []

        def worker(t):
            got = C.c_size_t()
            try:
                for i in range(t, len(wires_list), len(self.ctxs)):
                    w = wires_list[i]
                    ptr = w.data_ptr() if hasattr(w, "data_ptr") else np.ascontiguousarray(w, dtype=np.uint64).ctypes.data
                    buf = np.empty(words, dtype=np.uint64)
                    self.ctxs[t].check(fn(self.ctxs[t].handle, self.handles[t], ptr, None, buf.ctypes.data, words, C.byref(got)))
                    out[i] = buf[:got.value]
            except Exception as e:       # surfaced to the caller below
                errors.append(e)
        ths = [threading.Thread(target=worker, args=(t,)) for t in range(len(self.ctxs))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        if errors:
            raise errors[0]
        return out


def sharded_commit(ctx, dev_cols, ncols, log_n, rank, world, dist=None, rate_bits=3, cap_height=4):
    """One PolynomialBatch commitment split over `world` GPUs by coset.

    Every rank holds the column values (they are inputs of the proof on every rank), runs the
    cheap inverse NTT redundantly, then extends and Merkle-hashes only its 2^rate_bits / world
    leaf blocks.  The only exchange is the all-gather of each rank's cap entries (64 B .. 512 B),
    after which every rank knows MerkleTree::new(...).cap.  Returns (cap, handle)."""
    import ctypes as C
    import numpy as np
    nblk = 1 << rate_bits
    assert nblk % world == 0, "world size must divide the number of cosets"
    per = nblk // world
    h = C.c_void_p()
    part = np.empty(((1 << cap_height) * per // nblk, 4), dtype=np.uint64)
    ptr = dev_cols.data_ptr() if hasattr(dev_cols, "data_ptr") else int(dev_cols)
    ctx.check(ctx.lib.p2g_commit_blocks_from_values_dev(ctx.handle, ptr, ncols, log_n, rate_bits, cap_height,
                                                        rank * per, per, C.byref(h), part.ctypes.data))
    if world == 1 or dist is None:
        return part, h
    import torch
    dev = torch.device("cuda", ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(part.view(np.int64)).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    cap = torch.cat(parts).cpu().numpy().view(np.uint64)
    return cap, h


EXCHANGE_FN = None


def _exchange_type():
    import ctypes as C
    global EXCHANGE_FN
    if EXCHANGE_FN is None:
        EXCHANGE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_uint64)
    return EXCHANGE_FN


def sharded_prove(ctx, circuit, proof_words, wires, rank, world, all_gather, public_inputs=None):
    """One proof split over `world` GPUs by coset (p2g_prove_sharded, SURVEY.md section 8(e) split 2).

    `all_gather(recv, send, nbytes)` gathers the first nbytes of every rank's `send` (a uint8 CUDA tensor) into
    `recv` ([world * nbytes], rank order) and returns when recv is complete -- torch.distributed.all_gather_into_tensor
    on NCCL in production (nccl_all_gather below), a copy between host threads in the single-GPU tests.  Every
    rank passes the same wire matrix and gets the same proof back."""
    import ctypes as C
    import torch
    lib = ctx.lib
    nbytes = lib.p2g_shard_buffer_bytes(circuit, world)
    dev = torch.device("cuda", ctx.device)
    send = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    recv = torch.empty(nbytes * world, dtype=torch.uint8, device=dev)
    errors = []

    def cb(_user, stage, nb):
        try:
            all_gather(recv, send, int(nb))
            return 0
        except Exception as e:          # an exception must not cross the C frames
            errors.append(e)
            return 1
    fn = _exchange_type()(cb)
    out = np.empty(proof_words, dtype=np.uint64)
    got = C.c_size_t()
    wires = np.ascontiguousarray(wires, dtype=np.uint64)
    pi = np.ascontiguousarray(public_inputs, dtype=np.uint64) if public_inputs is not None and len(public_inputs) else None
    rc = lib.p2g_prove_sharded(ctx.handle, circuit, wires.ctypes.data, pi.ctypes.data if pi is not None else None, rank, world,
                               send.data_ptr(), recv.data_ptr(), nbytes, C.cast(fn, C.c_void_p), None, out.ctypes.data, proof_words,
                               C.byref(got))
    if errors:
        raise errors[0]
    ctx.check(rc)
    return out[:got.value]


def nccl_all_gather(dist, device):
    """the exchange of sharded_prove over torch.distributed (NCCL): one all_gather_into_tensor per stage"""
    import torch

    def gather(recv, send, nbytes):
        world = dist.get_world_size()
        dist.all_gather_into_tensor(recv[:world * nbytes], send[:nbytes])
        torch.cuda.synchronize(device)
    return gather


class ThreadedShards:
    """`world` coset shards of one proof as host threads on ONE GPU (tests, and the single-GPU emulation the
    profiling recipe asks for: no kernel ever waits for another rank, the rendezvous is on the host)."""

    def __init__(self, world):
        self.world, self.barrier = world, threading.Barrier(world)
        self.sends = [None] * world

    def gather_fn(self, rank):
        import torch

        def gather(recv, send, nbytes):
            self.sends[rank] = send
            self.barrier.wait()
            for r in range(self.world):
                recv[r * nbytes:(r + 1) * nbytes].copy_(self.sends[r][:nbytes])
            torch.cuda.synchronize()
            self.barrier.wait()
        return gather

    def prove(self, ctxs, circuits, proof_words, wires, public_inputs=None):
        out, errors = [None] * self.world, []

        def run(r):
            try:
                out[r] = sharded_prove(ctxs[r], circuits[r], proof_words, wires, r, self.world, self.gather_fn(r), public_inputs)
            except Exception as e:
                errors.append(e)
                self.barrier.abort()
        ths = [threading.Thread(target=run, args=(r,)) for r in range(self.world)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errors:
            raise errors[0]
        return out
