#!/usr/bin/env python3
"""bench.py — AES-GCM proofs/sec on B200 (BASELINE.json metric), with the LDE+Merkle roofline
and the CPU prover timed beside it.

A "step" = proving one batch of PROOFS_PER_STEP independent proofs of BASELINE config 2
(AES-GCM-128, 256-byte plaintext, GHASH tag; n = 2^15 rows, 135 wire columns) per GPU, each with
a different witness.  `value` is measured with the witnesses already resident in HBM
(p2g_prove_dev); `e2e` goes through the host API a user calls (CircuitData.prove_wires ->
p2g_prove) with pinned host witnesses, so it includes the 35 MB H2D copy per proof and the D2H
of every proof.  N > 1 shards independent proofs over ranks (no data-path collective; NCCL only
carries the barrier and the max-over-ranks time): weak scaling.

`--impl reference` times the CPU restatement of plonky2's prover (oracle/, OpenMP on all host
cores) on the same circuit and witnesses; one step = one proof.  The Rust reference itself cannot
be built in this image (no cargo/rustc; the prover lives in an un-vendored git dependency).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

L_BYTES = 256
PROOFS_PER_STEP = int(os.environ.get('P2G_BENCH_PROOFS_PER_STEP', '8'))
METRIC = "AES-128 block proofs/sec at 1/2/4/8 B200; LDE+Merkle GB/s vs HBM peak"
UNIT = "proofs/s"
WORKLOAD = ("AES-GCM-128 16-block (256 B) plaintext with GHASH tag, standard_recursion_config, "
            "n=2^15, 135 wires (BASELINE configs[1])")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_begin = index, [], None, 0.0

    def mark_begin(self):
        """samples received before this point (start-up, warm-up) are dropped"""
        self.t_begin = time.monotonic()

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.rows = [r for t, r in self.rows if t >= self.t_begin]
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_workload(count, seed=20261018):
    from tests import circuits
    data, _, tg = circuits.aes_gcm(L_BYTES, True)
    vals = circuits.gcm_inputs(tg, seed, count)
    return data, tg, vals


def run_reference(args, rank, world):
    """CPU arm: the oracle prover (port of the reference's CPU algorithm) on all host cores."""
    if rank != 0:
        return
    from tests import oracle_lib
    orc = oracle_lib.load()
    # all the host threads: torchrun exports OMP_NUM_THREADS=1 to its workers
    orc.lib.orc_set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    data, tg, vals = build_workload(2)
    wires = data.generate_witnesses(tg.input_targets(), vals)
    oc = oracle_lib.OracleCircuit(orc, data)
    for i in range(args.warmup):
        oc.prove(wires[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        proof = oc.prove(wires[i % 2])
    dt = time.perf_counter() - t0
    assert oc.verify(proof) == 0
    v = args.steps / dt
    cores = orc.lib.orc_num_threads()
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64 (Goldilocks)", "data": "synthetic",
           "config": {"workload": WORKLOAD, "proofs_per_step": 1},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{args.steps} proofs of the same circuit, one proof per step, OpenMP on {cores} threads"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=8, help="proofs in flight per GPU (one p2g context + host thread each)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from plonky2_aes_b200.host.polynomial_batch import Context, PolynomialBatch
    from plonky2_aes_b200.host import ffi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    T = max(1, args.streams)
    # host threads waiting for their proof's stream poll and yield (csrc/ctx.h, ctx_wait): robust when the
    # node has fewer cores than ranks x proofs in flight; P2G_SYNC=spin|block selects the other modes
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ.setdefault("P2G_SYNC", "yield")
    ctx = Context(local_rank)
    lib = ctx.lib
    ctxs = [ctx] + [Context(local_rank) for _ in range(T - 1)]
    streams = [torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", local_rank)) for c in ctxs]

    # ---- workload: distinct witnesses per proof and per rank ----
    B = PROOFS_PER_STEP
    data, tg, vals = build_workload(B, seed=20261018 + 1000 * rank)
    data.load(ctx)
    n, W = data.n, 135
    host_wires = torch.empty((B, W, n), dtype=torch.int64).pin_memory()
    data.generate_witnesses(tg.input_targets(), vals, out=host_wires.numpy().view(np.uint64))
    dev_wires = host_wires.cuda(non_blocking=False)
    words = data.proof_words
    proofs = torch.empty((B, words), dtype=torch.int64).pin_memory()
    got = C.c_size_t()
    handles = [data._gpu_circuit] + [data.load_handle(c) for c in ctxs[1:]]

    def run_worker(t, fn, src):
        g = C.c_size_t()
        for i in range(t, B, T):      # proofs dealt round-robin to the in-flight contexts
            ctxs[t].check(fn(ctxs[t].handle, handles[t], src[i].data_ptr(), None, proofs[i].data_ptr(), words, C.byref(g)))

    def step(fn, src):
        if T == 1:
            return run_worker(0, fn, src)
        ths = [threading.Thread(target=run_worker, args=(t, fn, src)) for t in range(T)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    def step_device():
        step(lib.p2g_prove_dev, dev_wires)

    def step_e2e():
        step(lib.p2g_prove, host_wires)

    def barrier():
        torch.cuda.synchronize()
        for c in ctxs:
            c.sync()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in streams]
        e0.record(streams[0])
        for _ in range(steps):
            fn()
        for e, st_ in zip(ends, streams):
            e.record(st_)
        barrier()
        ms = max(e0.elapsed_time(e) for e in ends)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # nvidia-smi needs a few hundred ms to produce its first line: start it before the warm-up and keep
    # the samples that arrive between the start of the device-timed leg and the end of the end-to-end
    # leg (both legs and the warm-up between them run the same proofs back to back)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    launches0 = lib.p2g_launch_count()
    sampler.mark_begin()
    ms_dev = timed(step_device, args.steps)
    launches = lib.p2g_launch_count() - launches0
    for _ in range(args.warmup):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    total_proofs = B * args.steps * world
    value = total_proofs / (ms_dev * 1e-3)
    e2e = total_proofs / (ms_e2e * 1e-3)

    # ---- parity spot check of what was just timed (rank 0): the restated verifier accepts ----
    verified = None
    stages = None
    roof = None
    cpu = None
    if rank == 0:
        from tests import oracle_lib
        orc = oracle_lib.load()
        oc = oracle_lib.OracleCircuit(orc, data)
        verified = oc.verify(proofs[B - 1].numpy().view(np.uint64)) == 0
        # per-stage device times of one proof
        ctx.check(lib.p2g_set_timing(ctx.handle, 1))
        g1 = C.c_size_t()       # one proof alone on the GPU, so the stage times are not stretched by the others
        ctx.check(lib.p2g_prove_dev(ctx.handle, handles[0], dev_wires[0].data_ptr(), None, proofs[0].data_ptr(), words, C.byref(g1)))
        tms = ffi.Timings()
        ctx.check(lib.p2g_last_timings(ctx.handle, C.byref(tms)))
        stages = {k: round(getattr(tms, k), 3) for k, _ in ffi.Timings._fields_}
        # ---- roofline of the dominant kernel group: LDE + Merkle of the 135-column wires batch ----
        peak, peak_kind = measured_peaks()
        times = []
        for _ in range(4):
            b = PolynomialBatch.from_values_device(ctx, dev_wires[0].data_ptr(), W, data.degree_bits)
            cm = (C.c_float * 3)()
            ctx.check(lib.p2g_last_commit_timings(ctx.handle, C.byref(cm)))
            times.append(list(cm))
            b.free()
        ctx.check(lib.p2g_set_timing(ctx.handle, 0))
        intt_ms, lde_ms, merkle_ms = [min(t[i] for t in times[1:]) for i in range(3)]
        N = 8 * n
        lde_bytes = 72 * W * n                    # read coeffs once, write the LDE once
        merkle_bytes = 8 * W * N + 96 * N         # read leaves, write digests + levels
        perms = N * ((W + 7) // 8) + N - 16
        poseidon_peak = ctx.poseidon_peak(32)
        roof = {"bound": "hbm", "kernel": "merkle_leaves_kernel (Poseidon leaf sponge + in-block tree levels)",
                "achieved": merkle_bytes / merkle_ms / 1e6, "peak": peak, "unit": "GB/s",
                "frac": merkle_bytes / merkle_ms / 1e6 / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the ncu --set full capture
                # profiles/r1_kernels_final_ncu_summary.txt (283.3 MB + 12.2 MB; algorithmic 308 MB)
                "traffic": 295.5e6, "algorithmic_bytes": merkle_bytes, "peak_kind": peak_kind + " copy bandwidth",
                "int_pipe": {"perms_per_s": perms / merkle_ms * 1e3, "peak_perms_per_s": poseidon_peak,
                             "frac": perms / merkle_ms * 1e3 / poseidon_peak,
                             "note": "Poseidon is INT-pipe bound (63 B of input per permutation); peak = chained "
                                     "permutations without memory traffic (p2g_poseidon_peak)",
                             # hardware-derived ceiling: the 30 x 204 FP64 instructions of the split-circulant MDS
                             # layers issue at one warp-instruction per 2.07 cycles per SM sub-partition
                             # (profiles/r1_int_pipes_microbench.jsonl), the S-box multiplies overlap on the IMAD pipe
                             "pipe_roofline_perms_per_s": 148 * 4 * 1.965e9 / 2.07 * 32 / 6120,
                             "pipe_roofline_frac": perms / merkle_ms * 1e3 / (148 * 4 * 1.965e9 / 2.07 * 32 / 6120)},
                "lde": {"ms": lde_ms, "achieved": lde_bytes / lde_ms / 1e6, "frac": lde_bytes / lde_ms / 1e6 / peak},
                "intt_ms": intt_ms, "merkle_ms": merkle_ms,
                "lde_merkle_gbs": (136 * W + 768) * n / (intt_ms + lde_ms + merkle_ms) / 1e6}
        # ---- CPU baseline: the oracle prover on the host cores, bounded sample ----
        if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only
            w0 = host_wires[0].numpy().view(np.uint64)
            t0 = time.perf_counter()
            cnt = 0
            while cnt < 2 and (cnt == 0 or time.perf_counter() - t0 < 12):
                p = oc.prove(w0)
                cnt += 1
            dt = time.perf_counter() - t0
            ctx.check(lib.p2g_prove(ctx.handle, data._gpu_circuit, host_wires[0].data_ptr(), None, proofs[0].data_ptr(), words, C.byref(got)))
            same = bool(np.array_equal(p, proofs[0].numpy().view(np.uint64)))
            cpu_threads = orc.lib.orc_num_threads()
            cpu = {"value": cnt / dt, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                   "sample": f"{cnt} proof(s) of the same circuit/witness by the CPU restatement (oracle/), OpenMP {cpu_threads} threads",
                   "proof_bit_identical_to_gpu": same}
        oc.free()
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u64 (Goldilocks p=2^64-2^32+1, exact)", "data": "synthetic",
               "config": {"workload": WORKLOAD, "proofs_per_step_per_gpu": B, "aes_blocks_per_proof": L_BYTES // 16,
                          "aes_block_proofs_per_s": value * (L_BYTES // 16), "l2": "per-proof working set 1.0 GB > 126 MB L2",
                          "proofs_in_flight_per_gpu": T, "host_wait": os.environ["P2G_SYNC"], "host_cores": cores,
                          "parallelism": f"independent proofs sharded over {world} GPU(s)"},
               "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * W * n * 8, "d2h_bytes_per_step": B * words * 8,
                       "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
               "stages_ms_one_proof": stages, "verifier_accepts": verified}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    main()
