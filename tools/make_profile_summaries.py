#!/usr/bin/env python3
"""Turns the artifacts of one profiling gpurun (gpurun_out/bench_final.json, launches_final.csv,
kernels_final.ncu-rep) into the tracked summaries under profiles/.
Usage: python tools/make_profile_summaries.py [tag]   (tag defaults to r1)"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

bench = json.loads(open(os.path.join(G, "bench_final.json")).read().strip().splitlines()[-1])
open(os.path.join(P, f"{tag}_bench_final.json"), "w").write(json.dumps(bench) + "\n")

rows = list(csv.reader(open(os.path.join(G, "launches_final.csv"))))
open(os.path.join(P, f"{tag}_launches_final.csv"), "w").write(open(os.path.join(G, "launches_final.csv")).read())
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ik, iv = H.index("Kernel Name"), H.index("Metric Value")
d = defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > iv:
        d[r[ik].split("(")[0]].append(float(r[iv].replace(",", "")) / 1000)
tot = sum(sum(v) for v in d.values())
nl = sum(len(v) for v in d.values())
L = [f"# Round {tag[1:]}, end state: ncu launch list summary", "",
     f"Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c {nl} --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1`",
     f"({nl} consecutive launches ~ 3 proofs of BASELINE config 2; cold-cache serialised times: compare shares; raw list: {tag}_launches_final.csv)", "",
     "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
pos = 0.0
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    L.append(f"| {k} | {len(v)} | {sum(v) / 1000:.3f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v):.1f} |")
    if "merkle" in k or "pow_grind" in k:
        pos += sum(v)
rf = bench["roofline"]
commit_ms = rf["merkle_ms"] + rf["lde"]["ms"] + rf["intt_ms"]
L += ["", f"Total {tot / 1000:.1f} ms.  Poseidon kernels (merkle_*, pow_grind) = {100 * pos / tot:.1f}% of device time "
          f"(86.4% in the first version, {tag}_launches_baseline_summary.md).", "",
      f"bench.py on the same build (profiles/{tag}_bench_final.json): value {bench['value']:.1f} proofs/s, e2e {bench['e2e']['value']:.1f} proofs/s, "
      f"cpu port {bench['cpu_baseline']['value']:.2f} proofs/s on {bench['cpu_baseline']['cores']} cores.",
      f"Stages of one proof alone on the GPU (ms, CUDA events): {bench['stages_ms_one_proof']}",
      f"Share check: the Merkle kernels are {100 * rf['merkle_ms'] / commit_ms:.0f}% of the wires commitment by CUDA events "
      f"({rf['merkle_ms']:.2f} of {commit_ms:.2f} ms) and the launch list gives merkle_leaves<1> the same dominant share of the step."]
open(os.path.join(P, f"{tag}_launches_final_summary.md"), "w").write("\n".join(L) + "\n")

raw = subprocess.run(["ncu", "-i", os.path.join(G, "kernels_final.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
want = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"] + \
       [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in
        ("long_scoreboard", "math_pipe_throttle", "no_instruction", "barrier", "wait", "not_selected", "dispatch_stall")]
idx = {w: H.index(w) for w in want if w in H}
ikn = H.index("Kernel Name")
out = ["# ncu --set full --clock-control none --import-source on -k regex:'merkle_leaves_kernel|ntt_dif_kernel|quotient_kernel' -s 6 -c 8",
       "# python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1   (BASELINE config 2, n = 2^15)",
       "# launches in order: iNTT wires, LDE wires, Merkle leaves wires (135 cols), iNTT zs, LDE zs (34 cols), Merkle leaves zs, quotient, iNTT of the quotient cosets", ""]
for r in rows[2:]:
    if len(r) < len(H):
        continue
    out.append(r[ikn].split("(")[0])
    out += [f"    {w} = {r[idx[w]]} {U[idx[w]]}" for w in want if w in idx]
open(os.path.join(P, f"{tag}_kernels_final_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(L[-6:]))
