#!/usr/bin/env python3
"""Turns the artifacts of one profiling gpurun (gpurun_out/bench_final.json, launches_final.csv,
kernels_final.ncu-rep) into the tracked summaries under profiles/: <tag>_bench_final.json,
<tag>_launches_final.csv + _summary.md, <tag>_kernels_final_ncu_summary.txt and <tag>_ncu_traffic.json
(DRAM bytes per launch of the dominant kernels, which bench.py reports as roofline.traffic).
Usage: python tools/make_profile_summaries.py [tag]   (tag defaults to r2)"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
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
     f"Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c {nl} --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 --config5-proofs 0`",
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
          f"(73.1% at the end of round 1, r1_launches_final_summary.md).", "",
      f"bench.py on the same build (profiles/{tag}_bench_final.json): value {bench['value']:.1f} proofs/s, e2e {bench['e2e']['value']:.1f} proofs/s, "
      f"config 5 (1024 distinct proofs) {bench['config5']['proofs_per_s_wall']:.1f} proofs/s wall." if bench.get('config5') else "",
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
       "# python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 1 --config5-proofs 0   (BASELINE config 2, n = 2^15)",
       "# launches in order: iNTT wires, LDE wires, Merkle leaves wires (135 cols), iNTT zs, LDE zs (34 cols), Merkle leaves zs, quotient, iNTT of the quotient cosets", ""]
for r in rows[2:]:
    if len(r) < len(H):
        continue
    out.append(r[ikn].split("(")[0])
    out += [f"    {w} = {r[idx[w]]} {U[idx[w]]}" for w in want if w in idx]
open(os.path.join(P, f"{tag}_kernels_final_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
# DRAM traffic per launch of the largest launch of each dominant kernel (wires batch: 135 columns)
traffic = {"source": f"profiles/{tag}_kernels_final_ncu_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                     "largest launch of each kernel = the 135-column wires batch)"}
def num(r, w):
    v, u = float(r[idx[w]].replace(",", "")), U[idx[w]].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1}.get(u, 1)
for key in ("merkle_leaves_kernel", "ntt_dif_kernel", "quotient_kernel"):
    best = None
    for r in rows[2:]:
        if len(r) >= len(H) and key in r[ikn] and (best is None or num(r, "gpu__time_duration.sum") > num(best, "gpu__time_duration.sum")):
            best = r
    if best is not None:
        traffic[key] = {"dram_bytes": num(best, "dram__bytes_read.sum") + num(best, "dram__bytes_write.sum"),
                        "duration_s_under_ncu": num(best, "gpu__time_duration.sum"),
                        "inst_executed": float(best[idx["smsp__inst_executed.sum"]].replace(",", "")),
                        "registers": int(float(best[idx["launch__registers_per_thread"]]))}
open(os.path.join(P, f"{tag}_ncu_traffic.json"), "w").write(json.dumps(traffic, indent=1) + "\n")
print("\n".join(L[-6:]))
