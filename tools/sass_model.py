#!/usr/bin/env python3
"""Instruction-count model of the hot kernels, read from the SASS of the built libp2gpu.so
(cuobjdump), written to profiles/r2_sass_model.json + SASS excerpts under profiles/.

bench.py turns these counts into the instruction-issue rooflines of its `roofline` block:
  * Poseidon: instructions per permutation = 8 x (full-round loop body) + 11 x (partial-pair loop body)
    + the straight-line rest; ceiling = SM sub-partitions x clock x 32 lanes / instructions.
  * NTT: SASS instructions of one general radix-2 butterfly (modular add + sub + multiply by a twiddle)
    and of one shift-butterfly of the last pass, counted on the device functions as inlined.
  * field multiply / add / sub instruction counts (quotient kernel floor = multiplies x mul count).
Run on the build box (no GPU needed):  python tools/sass_model.py
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "plonky2_aes_b200", "libp2gpu.so")
OUT = os.path.join(ROOT, "profiles")


def sass(fun_regex):
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs, cur, name = {}, None, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            cur = funcs.setdefault(name, [])
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur is not None:
            ins = re.sub(r"^@!?U?P[0-9T]+\s+", "", m.group(2))
            cur.append((int(m.group(1), 16), ins))
    return {k: v for k, v in funcs.items() if re.search(fun_regex, k)}


def loops(code):
    """innermost loops = backward branches with no other backward branch inside: [(start, end)]"""
    back = []
    for a, ins in code:
        m = re.match(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) < a:
            back.append((int(m.group(1), 16), a))
    inner = [l for l in back if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in back)]
    return sorted(inner), sorted(back)


def hist(code, lo=None, hi=None):
    c = collections.Counter()
    for a, ins in code:
        if (lo is None or a >= lo) and (hi is None or a <= hi):
            c[ins.split()[0]] += 1
    return c


PIPE = {"fp64": ("DFMA", "DADD", "DMUL"), "imad_wide": ("IMAD.WIDE.U32", "IMAD.WIDE"),
        "fma_int": ("IMAD.X", "IMAD.MOV", "IMAD.MOV.U32", "IMAD", "IMAD.U32", "IMAD.SHL.U32", "IMAD.IADD"),
        "alu": ("IADD3", "IADD3.X", "LOP3.LUT", "SEL", "ISETP.GE.U32.AND", "ISETP.GE.U32.AND.EX", "SHF.R.U32.HI", "SHF.L.U32", "VIADD"),
        "convert": ("I2F.F64.U32",)}


def by_pipe(c):
    out = {k: sum(c.get(i, 0) for i in v) for k, v in PIPE.items()}
    out["total"] = sum(c.values())
    out["other"] = out["total"] - sum(v for k, v in out.items() if k != "total")
    return out


def main():
    model = {}
    f = sass(r"poseidon_bench_kernel")
    (name, code), = f.items()
    inner, _ = loops(code)
    assert len(inner) == 2, inner
    full, pair = hist(code, *inner[0]), hist(code, *inner[1])
    pf, pp = by_pipe(full), by_pipe(pair)
    per_perm = {k: 8 * pf[k] + 11 * pp[k] for k in pf}
    # straight-line part per permutation (initial constants, loop control): everything of the permutation
    # body outside the two inner loops, measured as outer-loop body minus the inner loops
    model["poseidon"] = {"kernel": name, "full_round_loop": pf, "partial_pair_loop": pp, "per_permutation": per_perm,
                         "note": "per_permutation = 8 x full-round loop body + 11 x partial-pair loop body (96 + 22 S-boxes)"}
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "r2_sass_poseidon_full_round.txt"), "w") as o:
        o.write(f"# {name}: full-round loop body [{inner[0][0]:#x}, {inner[0][1]:#x}] ({pf['total']} instructions: 12 S-boxes + split-circulant MDS + 12 read-outs)\n")
        o.writelines(f"/*{a:04x}*/ {i}\n" for a, i in code if inner[0][0] <= a <= inner[0][1])
    with open(os.path.join(OUT, "r2_sass_poseidon_partial_pair.txt"), "w") as o:
        o.write(f"# {name}: partial-pair loop body [{inner[1][0]:#x}, {inner[1][1]:#x}] ({pp['total']} instructions: 2 S-boxes + dense 12x12 + 13 read-outs)\n")
        o.writelines(f"/*{a:04x}*/ {i}\n" for a, i in code if inner[1][0] <= a <= inner[1][1])
    # field ops: the p2g_field_ops test kernel holds one add, sub, mul, canon ... ; count through a tiny probe
    f = sass(r"field_ops_kernel")
    (name, code), = f.items()
    model["field_ops_kernel_total"] = by_pipe(hist(code))
    # NTT: the kernel's radix-16 general pass loop and the last (shift) pass loop
    f = sass(r"ntt_dif_kernel")
    (name, code), = f.items()
    inner, allb = loops(code)
    ntt = []
    for lo, hi in inner:
        h = by_pipe(hist(code, lo, hi))
        h["range"] = [hex(lo), hex(hi)]
        h["lds"] = sum(v for k, v in hist(code, lo, hi).items() if k.startswith("LDS"))
        h["sts"] = sum(v for k, v in hist(code, lo, hi).items() if k.startswith("STS"))
        ntt.append(h)
    model["ntt_dif_kernel"] = {"kernel": name, "inner_loops": ntt,
                               "note": "a radix-16 pass body holds 32 butterflies (16 points x 4 stages / 2); the pass with 16 LDS + 16 STS "
                                       "and the most IMAD.WIDE is the general-twiddle pass, the one without twiddle LDS the last (shift) pass"}
    gen = max((l for l in ntt if l["lds"] >= 16), key=lambda l: l["imad_wide"], default=None)
    if gen:
        model["ntt_dif_kernel"]["general_pass_instr_per_butterfly"] = gen["total"] / 32.0
        lo, hi = int(gen["range"][0], 16), int(gen["range"][1], 16)
        with open(os.path.join(OUT, "r2_sass_ntt_radix16_pass.txt"), "w") as o:
            o.write(f"# {name}: general radix-16 pass loop body ({gen['total']} instructions = 32 butterflies with twiddle multiplies)\n")
            o.writelines(f"/*{a:04x}*/ {i}\n" for a, i in code if lo <= a <= hi)
    f = sass(r"quotient_kernelILb0")
    (name, code), = f.items()
    model["quotient_kernel"] = {"kernel": name, "static": by_pipe(hist(code))}
    with open(os.path.join(OUT, "r2_sass_model.json"), "w") as o:
        json.dump(model, o, indent=1)
    print(json.dumps({"poseidon_instr_per_perm": per_perm, "ntt_general_pass_instr_per_butterfly": model["ntt_dif_kernel"].get("general_pass_instr_per_butterfly")}))


if __name__ == "__main__":
    sys.exit(main())
