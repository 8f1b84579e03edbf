#!/usr/bin/env python3
"""Emit plonky2_aes_b200/csrc/poseidon_rc_f64.inc from poseidon_rc.inc.

Two families of exact-double tables for the FP64-pipe MDS layer of csrc/poseidon.cuh:
  POSEIDON_RCD_LO/HI[31*12]  2^52 + low/high 32 bits of the round constants (plain 12x12 form);
  POSEIDON_RCS_LO/HI[31*12]  biases of the two-level split-circulant form (see split2): per row
        [0..2]  pp_r = 2^52 + (beta_r + beta_{r+3}) / 2,   [3..5]  pm_r = (beta_r - beta_{r+3}) / 2,
        [6..11] acc-_j = (c_j - c_{j+6}) / 2,              beta_j = (c_j + c_{j+6}) / 2,
     so that acc+_r = pp_r + pm_r, acc+_{r+3} = pp_r - pm_r, acc+ + acc- = 2^52 + c_j + y_j and
     acc+ - acc- = 2^52 + c_{j+6} + y_{j+6}.
Row r holds the constants added BEFORE round r (rows 0..29); row 30 is zero.
"""
import os, re, sys
from fractions import Fraction

here = os.path.dirname(os.path.abspath(__file__))
csrc = os.path.join(here, "..", "plonky2_aes_b200", "csrc")
rc = [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]+)ULL", open(os.path.join(csrc, "poseidon_rc.inc")).read())]
assert len(rc) == 360
rc += [0] * 12
rc_true = list(rc)


def fmt(v):
    """exact decimal literal of a multiple of 1/2"""
    v = Fraction(v)
    assert v.denominator in (1, 2)
    assert float(v) == v          # representable
    return ("%d.0" % v.numerator) if v.denominator == 1 else ("%d.5" % (v.numerator // 2))


def table(name, vals):
    out = ["__constant__ double %s[%d] = {" % (name, len(vals))]
    for i in range(0, len(vals), 4):
        out.append("  " + ", ".join(fmt(v) for v in vals[i:i + 4]) + ",")
    out.append("};")
    return "\n".join(out)


# The accumulators are read back as raw mantissa words WITHOUT masking off the exponent: the bit pattern
# of the double 2^52 + v is 0x4330_0000_0000_0000 + v, so combining the raw words of the (low, high)
# pair as  lo_bits + hi_bits * 2^32  adds  E = 0x43300000 * 2^32 * (1 + 2^32)  to every state word.
# All bias tables therefore carry (c - E) mod p instead of the round constant c.
PRIME = 0xFFFFFFFF00000001
E_READ = (0x43300000 << 32) * (1 + (1 << 32)) % PRIME
rc = [(c - E_READ) % PRIME for c in rc]


def halves(sel):
    return [sel(c) for c in rc]


lo = halves(lambda c: c & 0xFFFFFFFF)
hi = halves(lambda c: c >> 32)


# Split-circulant biases.  (C_j +- C_{j+6}) / 2 are integers for this MDS matrix (and for C cyclically convolved
# with itself), so the two accumulators of a row pair hold integers:  acc+ = 2^52 + (k_j + k_{j+6}) / 2 + ...,
# acc- = (k_j - k_{j+6}) / 2 + ...  (unbiased, signed), and  acc+ + acc- = 2^52 + k_j + y_j,
# acc+ - acc- = 2^52 + k_{j+6} + y_{j+6}  need no further bias.  k_j + k_{j+6} is made even, per 32-bit half, by
# adding one of three (low, high) pairs that are 0 mod p to (k_{j+6} low, k_{j+6} high).
DL, DH = 2**48 - 2**32 + 2**16 + 1, 2**48 + 2**32 - 2**17           # odd, even
PRIME_ = 0xFFFFFFFF00000001
assert (DL + (DH << 32)) % PRIME_ == 0 and (1 + ((2**32 - 1) << 32)) % PRIME_ == 0
PARITY_FIX = {(0, 0): (0, 0), (1, 0): (DL, DH), (0, 1): (DL + 1, DH + 2**32 - 1), (1, 1): (1, 2**32 - 1)}


def split2(lo_rows, hi_rows):
    """lo_rows / hi_rows: lists of 12-entry rows -> (table_lo, table_hi) in the [acc+ x6, acc- x6] layout"""
    out_lo, out_hi = [], []
    for lo_r, hi_r in zip(lo_rows, hi_rows):
        lo_r, hi_r = list(lo_r), list(hi_r)
        for j in range(6):
            a, b = PARITY_FIX[((lo_r[j] + lo_r[j + 6]) & 1, (hi_r[j] + hi_r[j + 6]) & 1)]
            lo_r[j + 6] += a; hi_r[j + 6] += b
        # second level: acc+ itself is a 6 x 6 cyclic product and splits again ((P_k +- P_{k+3}) / 2 are integers too):
        # acc+_r = pp_r + pm_r, acc+_{r+3} = pp_r - pm_r, so beta_r + beta_{r+3} (beta_r = (k_r + k_{r+6}) / 2) must be
        # even: fixed on k_{r+9} with twice the pairs above (even, so the first-level parity stays)
        for r in range(3):
            need = tuple(((row[r] + row[r + 6] + row[r + 3] + row[r + 9]) // 2) & 1 for row in (lo_r, hi_r))
            a, b = PARITY_FIX[need]
            lo_r[r + 9] += 2 * a; hi_r[r + 9] += 2 * b
        for row, out in ((lo_r, out_lo), (hi_r, out_hi)):
            beta = [(row[j] + row[j + 6]) // 2 for j in range(6)]
            assert all((row[j] + row[j + 6]) % 2 == 0 for j in range(6)) and all((beta[r] + beta[r + 3]) % 2 == 0 for r in range(3))
            out += [2**52 + (beta[r] + beta[r + 3]) // 2 for r in range(3)]
            out += [(beta[r] - beta[r + 3]) // 2 for r in range(3)]
            out += [(row[j] - row[j + 6]) // 2 for j in range(6)]
    return out_lo, out_hi


def split_product(xs, Cx, pp, pm, am):
    """the device's accumulation order for one 32-bit half: groups (j, j+3), j = 1, 2, 0"""
    for j in (1, 2, 0):
        xp = [None] * 2
        for u, jj in enumerate((j, j + 3)):
            xp[u] = dbl(xs[jj] + xs[jj + 6])
            xm = dbl(xs[jj] - xs[jj + 6])
            for r in range(6):
                a, b = Cx[(jj - r) % 12], Cx[(jj + 6 - r) % 12]
                assert (a - b) % 2 == 0
                am[r] = dbl(am[r] + xm * ((a - b) // 2))
        up, um = dbl(xp[0] + xp[1]), dbl(xp[0] - xp[1])
        for r in range(3):
            a = (Cx[(j - r) % 12] + Cx[(j + 6 - r) % 12]) // 2
            b = (Cx[(j + 3 - r) % 12] + Cx[(j + 9 - r) % 12]) // 2
            assert (a + b) % 2 == 0
            pp[r] = dbl(pp[r] + up * ((a + b) // 2))
            pm[r] = dbl(pm[r] + um * ((a - b) // 2))


def split_recombine(pp, pm, am):
    ap = [dbl(pp[r] + pm[r]) for r in range(3)] + [dbl(pp[r] - pm[r]) for r in range(3)]
    return [dbl(ap[r] + am[r]) for r in range(6)] + [dbl(ap[r] - am[r]) for r in range(6)]


text = """/* generated by tools/gen_poseidon_f64.py from poseidon_rc.inc (constants: tools/gen_poseidon_constants.py).
 * POSEIDON_RCD_*: 2^52 + low / high 32 bits of every round constant, as exact doubles;
 * POSEIDON_RCS_*: biases of the split-circulant MDS (2^52 + (c_j + c_{j+6})/2 and (c_j - c_{j+6})/2, see the generator);
 * row r holds the constants ADDED BEFORE round r (rows 0..29) and row 30 is zero. */
"""
text += table("POSEIDON_RCD_LO", [2**52 + v for v in lo]) + "\n"
text += table("POSEIDON_RCD_HI", [2**52 + v for v in hi]) + "\n"
# + 2 (DL, DH) on every word: the full rounds feed the S-box outputs as UNFOLDED products (low limb l0 - h0 - h1 in
# (-2^33, 2^32), high limb l1 + h0 < 2^33), so the variable part of an accumulator can be as low as -264 * 2^33
RCS_LO, RCS_HI = split2([[v + 2 * DL for v in lo[12 * r:12 * r + 12]] for r in range(31)],
                        [[v + 2 * DH for v in hi[12 * r:12 * r + 12]] for r in range(31)])
text += table("POSEIDON_RCS_LO", RCS_LO) + "\n"
text += table("POSEIDON_RCS_HI", RCS_HI) + "\n"
open(os.path.join(csrc, "poseidon_rc_f64.inc"), "w").write(text)
print("wrote poseidon_rc_f64.inc", len(text), "bytes")

# ---- merged pairs of partial rounds (csrc/poseidon.cuh, poseidon_partial_pair) -------------------
# Two consecutive partial rounds  s' = (sbox(s0), s1..s11); t = M s' + cA;  t' = (sbox(t0), t1..);
# u = M t' + cB  collapse to   u = A s' + col0(M) sbox(t0) + K,   t0 = row0(M) s' + cA_0,
# with A = M~ M (M~ = M with column 0 zeroed, entries < 2^15: still exact in FP64 on 32-bit halves)
# and K = M~ cA + cB.  Tables: POSEIDON_PAIR_A (144 literals, row-major) and the 2^52-biased
# halves of K for the 11 pairs.
P = 0xFFFFFFFF00000001
CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
M = [[CIRC[(i - r) % 12] + (8 if r == 0 and i == 0 else 0) for i in range(12)] for r in range(12)]
A = [[sum(M[r][k] * M[k][j] for k in range(1, 12)) for j in range(12)] for r in range(12)]
assert max(max(row) for row in A) < 2**15 and max(sum(row) for row in A) < 2**17
K = []
for p in range(11):
    kA = 4 + 2 * p
    cA = rc_true[12 * (kA + 1):12 * (kA + 2)]
    cB = rc_true[12 * (kA + 2):12 * (kA + 3)]
    K += [(sum(M[r][k] * cA[k] for k in range(1, 12)) + cB[r]) % P for r in range(12)]
text = "#define POSEIDON_PAIR_A_INIT { " + ", ".join("%d." % v for row in A for v in row) + " }\n"
KE = [(v - E_READ) % P for v in K]            # raw-word read-out offset, see E_READ above
text += table("POSEIDON_PAIRK_LO", [2**52 + (v & 0xFFFFFFFF) for v in KE]) + "\n"
text += table("POSEIDON_PAIRK_HI", [2**52 + (v >> 32) for v in KE]) + "\n"
open(os.path.join(csrc, "poseidon_rc_f64.inc"), "a").write(text)


# self-check of the algebra against the plain round function
def sbox(x): return pow(x, 7, P)
def mds(s): return [sum(M[r][i] * s[i] for i in range(12)) % P for r in range(12)]
import random
random.seed(1)
s = [random.randrange(P) for _ in range(12)]
for p in range(11):
    kA = 4 + 2 * p
    ref = list(s)
    for k in (kA, kA + 1):
        ref[0] = sbox(ref[0])
        ref = [(a + b) % P for a, b in zip(mds(ref), rc_true[12 * (k + 1):12 * (k + 2)])]
    sp = [sbox(s[0])] + s[1:]
    t0 = (sum(M[0][j] * sp[j] for j in range(12)) + rc_true[12 * (kA + 1)]) % P
    z0 = sbox(t0)
    u = [(sum(A[r][j] * sp[j] for j in range(12)) + M[r][0] * z0 + K[12 * p + r]) % P for r in range(12)]
    assert u == ref, p
    s = ref
print("pair tables ok")

# ---- split form of the merged pair (csrc/poseidon.cuh, poseidon_partial_pair with P2G_PAIR_SPLIT) -----------
# u = M t' + cB with t = M s' + cA and t' = t + (z - t0) e0, z = sbox(t0), gives
#     u = M^2 s' + (z - t0) col0(M) + (M cA + cB),
# and with M = circ(C) + 8 E00:   M^2 = circ(C)^2 + 8 col0(circ) e0^T + 8 e0 row0(circ) + 64 E00, so
#     u = circ(C2) s' + 8 cc y0 + cm (z - t0) + 8 (t0 - cA_0) e0 + K'
#       = circ(C2) s' + cc (8 y0 + z - t0) + 8 (z - cA_0) e0 + K',
# C2 = C (*) C (cyclic), cc = col0(circ), cm = col0(M) = cc + 8 e0, y0 = s'_0, K' = M cA + cB.  circ(C2) is
# circulant, so it splits like the full-round layer (72 products per 32-bit half instead of 144); the rank-one
# term costs 24 products and word 0 two more.  260 FP64 instructions per pair instead of 336.
# On the device t0 sits in accumulators biased by 2^52 whose raw words carry the read-out offset E_READ:
#     dd = (2^52 + z) - t_b = z - t0 + E_READ  (per half, exact),     gg = dd + 8 y0,
# so  cc (8 y0 + z - t0) = cc gg - cc E_READ;  the constants go into KS below, together with an offset
# DL + 2^32 DH = 0 (mod p) that keeps every accumulator positive although cc gg may be negative.
C2 = [sum(CIRC[i] * CIRC[(k - i) % 12] for i in range(12)) for k in range(12)]
cc = [CIRC[(12 - r) % 12] for r in range(12)]
cm = [cc[r] + (8 if r == 0 else 0) for r in range(12)]
KL_ROWS, KH_ROWS = [], []
for p in range(11):
    kA = 4 + 2 * p
    cA = rc_true[12 * (kA + 1):12 * (kA + 2)]
    cB = rc_true[12 * (kA + 2):12 * (kA + 3)]
    kf = [(sum(M[r][k] * cA[k] for k in range(12)) + cB[r] - cc[r] * E_READ - (8 * cA[0] if r == 0 else 0) - E_READ) % P
          for r in range(12)]
    KL_ROWS.append([(v & 0xFFFFFFFF) + 2 * DL for v in kf])           # + 2 (DL, DH): positive and parity-neutral
    KH_ROWS.append([(v >> 32) + 2 * DH for v in kf])
KS_LO, KS_HI = split2(KL_ROWS, KH_ROWS)
assert all((C2[k] + C2[(k + 6) % 12]) % 2 == 0 for k in range(12)) and all((CIRC[k] + CIRC[(k + 6) % 12]) % 2 == 0 for k in range(12))
text = "#define POSEIDON_C2_INIT { " + ", ".join("%d." % v for v in C2) + " }\n"
text += table("POSEIDON_PAIRS_LO", KS_LO) + "\n"
text += table("POSEIDON_PAIRS_HI", KS_HI) + "\n"
open(os.path.join(csrc, "poseidon_rc_f64.inc"), "a").write(text)


# exact emulation of the device arithmetic of the split pair: every intermediate must be a double
def dbl(v):
    v = Fraction(v)
    a = abs(v)
    if a == 0: return v
    e = 0
    while a >= 2**53: a /= 2; e += 1
    while a < 2**52: a *= 2; e -= 1
    assert a.denominator == 1, ("not representable", float(v))
    return v


def raw_words(d):
    """raw (low, high) 32-bit words of a double in [2^52, 2^53)"""
    assert 2**52 <= d < 2**53 and d.denominator == 1
    bits = (0x433 << 52) + (int(d) - 2**52)
    return bits & 0xFFFFFFFF, bits >> 32


def readout(l, h):
    al0, al1 = raw_words(l)
    ah0, ah1 = raw_words(h)
    return (al0 + ((al1 + ah0) << 32) + (ah1 << 64)) % P


def pair_split_emulated(s, p):
    kA = 4 + 2 * p
    y0 = sbox(s[0])
    x = [y0] + s[1:]
    xl = [Fraction(v & 0xFFFFFFFF) for v in x]
    xh = [Fraction(v >> 32) for v in x]
    cA0e = (rc_true[12 * (kA + 1)] - E_READ) % P
    out = []
    accs = []
    tb = []
    for half, (xs, tab, t_init) in enumerate(((xl, KS_LO, 2**52 + (cA0e & 0xFFFFFFFF)), (xh, KS_HI, 2**52 + (cA0e >> 32)))):
        pp = [Fraction(tab[12 * p + r]) for r in range(3)]
        pm = [Fraction(tab[12 * p + 3 + r]) for r in range(3)]
        am = [Fraction(tab[12 * p + 6 + r]) for r in range(6)]
        t = Fraction(t_init)
        split_product(xs, C2, pp, pm, am)
        for j in range(12):
            t = dbl(t + xs[j] * M[0][j])
        accs.append(split_recombine(pp, pm, am)); tb.append(t)
    t0 = readout(tb[0], tb[1])
    assert t0 == (sum(M[0][j] * x[j] for j in range(12)) + rc_true[12 * (kA + 1)]) % P
    z = sbox(t0)
    for half, (zz, xs) in enumerate(((z & 0xFFFFFFFF, xl), (z >> 32, xh))):
        mz = Fraction(2**52 + zz)
        gg = dbl(dbl(mz - tb[half]) + 8 * xs[0])
        for r in range(12):
            accs[half][r] = dbl(accs[half][r] + gg * cc[r])
        accs[half][0] = dbl(accs[half][0] + 8 * dbl(mz - 2**52))
    return [readout(accs[0][r], accs[1][r]) for r in range(12)]


for trial in range(20):
    if trial == 0: s = [P - 1] * 12
    elif trial == 1: s = [0] * 12
    elif trial == 2: s = [0xFFFFFFFF] * 12
    elif trial == 3: s = [0xFFFFFFFF00000000] * 12
    else: s = [random.randrange(P) for _ in range(12)]
    for p in range(11):
        kA = 4 + 2 * p
        ref = list(s)
        for k in (kA, kA + 1):
            ref[0] = sbox(ref[0])
            ref = [(a + b) % P for a, b in zip(mds(ref), rc_true[12 * (k + 1):12 * (k + 2)])]
        got = pair_split_emulated(s, p)
        assert got == ref, (trial, p)
        s = ref
# worst-case magnitudes: the inputs are lazy residues, any u64 (halves up to 2^32 - 1), in the patterns that
# maximise X+ / X- and the negative rank-one term
U = 0xFFFFFFFFFFFFFFFF
for pat in ([U] * 12, [U] * 6 + [0] * 6, [0] * 6 + [U] * 6, [0] + [U] * 11, [U] + [0] * 11,
            [U if j % 2 else 0 for j in range(12)], [0xFFFFFFFF00000000] * 6 + [0xFFFFFFFF] * 6):
    for p in range(11):
        kA = 4 + 2 * p
        ref = [v % P for v in pat]
        for k in (kA, kA + 1):
            ref[0] = sbox(ref[0])
            ref = [(a + b) % P for a, b in zip(mds(ref), rc_true[12 * (k + 1):12 * (k + 2)])]
        assert pair_split_emulated(list(pat), p) == ref
print("split pair tables ok")


# exact emulation of the split full round (poseidon_round<true>): s <- M sbox(s) + rc[next_row].  The S-box outputs
# arrive as limb pairs (X_lo, X_hi) with value X_lo + 2^32 X_hi (mod p): the unfolded 128-bit product x^3 * x^4 gives
# X_lo = l0 - h0 - h1, X_hi = l1 + h0
def full_round_emulated(limbs, next_row):
    halves = []
    for xs, tab in (([Fraction(a) for a, _ in limbs], RCS_LO), ([Fraction(b) for _, b in limbs], RCS_HI)):
        pp = [Fraction(tab[12 * next_row + r]) for r in range(3)]
        pm = [Fraction(tab[12 * next_row + 3 + r]) for r in range(3)]
        am = [Fraction(tab[12 * next_row + 6 + r]) for r in range(6)]
        split_product(xs, CIRC, pp, pm, am)
        pp[0] = dbl(pp[0] + 2 * xs[0]); pm[0] = dbl(pm[0] + 2 * xs[0]); am[0] = dbl(am[0] + 4 * xs[0])
        halves.append(split_recombine(pp, pm, am))
    return [readout(halves[0][r], halves[1][r]) for r in range(12)]


LIM = 2**33 - 2
pats = [[(random.randrange(-LIM, 2**32), random.randrange(0, LIM)) for _ in range(12)] for _ in range(4)]
pats += [[(-LIM, LIM)] * 12, [(2**32 - 1, LIM)] * 12, [(-LIM, 0)] * 12, [(0, 0)] * 12,
         [(-LIM, LIM) if j < 6 else (2**32 - 1, 0) for j in range(12)], [(2**32 - 1, 0) if j < 6 else (-LIM, LIM) for j in range(12)],
         [(-LIM, LIM) if j % 2 else (2**32 - 1, 0) for j in range(12)], [(-LIM, LIM) if (j // 3) % 2 else (2**32 - 1, 0) for j in range(12)]]
for pat in pats:
    vals = [(a + (b << 32)) % P for a, b in pat]
    for row in range(1, 31):
        ref = [(a + b) % P for a, b in zip(mds(vals), rc_true[12 * row:12 * row + 12])]
        assert full_round_emulated(pat, row) == ref, row
print("split full round ok")


# ---- the pair in the (E, F) basis (csrc/poseidon.cuh, poseidon_partial_pair_basis) ---------------------------------
# Between two pairs the state is kept as E_j = u_j + u_{j+6}, F_j = u_j - u_{j+6} (j < 6): a pair's split accumulators
# give exactly these (acc+ = E / 2, acc- = F / 2; all coefficients doubled here, which is free), so neither the
# recombination at the end of a pair nor the X+- formation at the start of the next one is needed -- 198 instead of 236
# FP64 instructions per pair.  Word 0, the only one that meets an S-box, is (E_0 + F_0) / 2, exact on the accumulators.
# Per pair and 32-bit half the table holds [KPP_0..2, KPM_0..2, KF_0..5]: E'_r = pp_r + pm_r, E'_{r+3} = pp_r - pm_r.
# Every KE_r / KF_r is even, so (E'_0 + F'_0) / 2 and, in the last pair, (E'_r +- F'_r) / 2 are integers.
def even_rep(v, mult):
    """(lo, hi) of v (mod p) with both parts even and offset by mult * (DL, DH)"""
    lo, hi = (v & 0xFFFFFFFF) + mult * DL, (v >> 32) + mult * DH
    a, b = PARITY_FIX[(lo & 1, hi & 1)]
    return lo + a, hi + b


PB_LO, PB_HI = [], []
for p in range(11):
    kA = 4 + 2 * p
    cA = rc_true[12 * (kA + 1):12 * (kA + 2)]
    cB = rc_true[12 * (kA + 2):12 * (kA + 3)]
    Kp = [(sum(M[r][k] * cA[k] for k in range(12)) + cB[r]) % P for r in range(12)]
    KE, KF = [], []
    for r in range(6):
        e0 = -8 * cA[0] if r == 0 else 0
        KE.append(even_rep((Kp[r] + Kp[r + 6] + e0 - (cc[r] + cc[r + 6]) * E_READ - E_READ) % P, 4))
        KF.append(even_rep((Kp[r] - Kp[r + 6] + e0 - (cc[r] - cc[r + 6]) * E_READ - E_READ) % P, 4))
    for half, out in ((0, PB_LO), (1, PB_HI)):
        out += [(KE[r][half] + KE[r + 3][half]) // 2 for r in range(3)]
        out += [(KE[r][half] - KE[r + 3][half]) // 2 for r in range(3)]
        out += [2**52 + KF[r][half] for r in range(6)]
        assert all((KE[r][half] + KE[r + 3][half]) % 2 == 0 for r in range(3))
    # pp carries the 2^52 of E' (pp + pm and pp - pm are both 2^52 + ...)
    for out in (PB_LO, PB_HI):
        for r in range(3):
            out[12 * p + r] += 2**52
EXIT_LO, EXIT_HI = even_rep((-E_READ) % P, 4)
# the t accumulator (row 0 of M on the (E, F) inputs) has negative coefficients on F: a small offset that is 0 mod p
# (E_0 = y0 + x6 and F_0 = y0 - x6 can be as low as -2^32 per half, F_3 / F_4 meet -1 / -16: the sum is above -42 * 2^32)
TOFF = (44 * 2**32 + 45, 45 * 2**32 - 89)
assert (TOFF[0] + (TOFF[1] << 32)) % P == 0
text = table("POSEIDON_PAIRB_LO", PB_LO) + "\n" + table("POSEIDON_PAIRB_HI", PB_HI) + "\n"
text += "#define POSEIDON_PAIRB_EXIT_LO %s\n#define POSEIDON_PAIRB_EXIT_HI %s\n" % (fmt(2**52 + EXIT_LO), fmt(2**52 + EXIT_HI))
PBT_LO = [2**52 + ((rc_true[12 * (5 + 2 * p)] - E_READ) % P & 0xFFFFFFFF) + TOFF[0] for p in range(11)] + [0]
PBT_HI = [2**52 + ((rc_true[12 * (5 + 2 * p)] - E_READ) % P >> 32) + TOFF[1] for p in range(11)] + [0]
text += table("POSEIDON_PAIRB_T_LO", PBT_LO) + "\n" + table("POSEIDON_PAIRB_T_HI", PBT_HI) + "\n"
open(os.path.join(csrc, "poseidon_rc_f64.inc"), "a").write(text)


def pair_basis_emulated(st, p, first, last, force_y0=None, force_z=None):
    """st: natural state (first) or [e_0..e_5, x0, f_1..f_5] as field elements (any representative < 2^64);
    force_y0 / force_z replace the two S-box outputs by given lazy words (range checks only)"""
    kA = 4 + 2 * p
    cA0e = (rc_true[12 * (kA + 1)] - E_READ) % P
    halves = lambda v: (Fraction(v & 0xFFFFFFFF), Fraction(v >> 32))
    x0 = st[0] if first else st[6]
    y0 = sbox(x0) if force_y0 is None else force_y0
    Eo, Fo, tb, y0h, x0h = [], [], [], halves(y0), halves(x0)
    for h, (tabp, t_init) in enumerate(((PB_LO, 2**52 + (cA0e & 0xFFFFFFFF) + TOFF[0]), (PB_HI, 2**52 + (cA0e >> 32) + TOFF[1]))):
        if first:
            xs = [halves(v)[h] for v in st]
            E = [dbl(xs[j] + xs[j + 6]) for j in range(6)]
            F = [dbl(xs[j] - xs[j + 6]) for j in range(6)]
            x6 = xs[6]
        else:
            E = [halves(st[j])[h] for j in range(6)]
            F = [None] + [halves(st[6 + j])[h] for j in range(1, 6)]
            x6 = dbl(E[0] - x0h[h])
        E[0] = dbl(y0h[h] + x6); F[0] = dbl(y0h[h] - x6)
        pp = [Fraction(tabp[12 * p + r]) for r in range(3)]
        pm = [Fraction(tabp[12 * p + 3 + r]) for r in range(3)]
        Fa = [Fraction(tabp[12 * p + 6 + r]) for r in range(6)]
        t = Fraction(t_init)
        for j in (1, 2, 0):
            for jj in (j, j + 3):
                for r in range(6):
                    a, b = C2[(jj - r) % 12], C2[(jj + 6 - r) % 12]
                    Fa[r] = dbl(Fa[r] + F[jj] * (a - b))
                m0a = M[0][jj]; m0b = M[0][jj + 6]
                assert (m0a + m0b) % 2 == 0
                t = dbl(t + E[jj] * ((m0a + m0b) // 2)); t = dbl(t + F[jj] * ((m0a - m0b) // 2))
            up, um = dbl(E[j] + E[j + 3]), dbl(E[j] - E[j + 3])
            for r in range(3):
                a = (C2[(j - r) % 12] + C2[(j + 6 - r) % 12]) // 2
                b = (C2[(j + 3 - r) % 12] + C2[(j + 9 - r) % 12]) // 2
                pp[r] = dbl(pp[r] + up * (a + b)); pm[r] = dbl(pm[r] + um * (a - b))
        Eo.append([dbl(pp[r] + pm[r]) for r in range(3)] + [dbl(pp[r] - pm[r]) for r in range(3)])
        Fo.append(Fa); tb.append(t)
    t0 = readout(tb[0], tb[1])
    z = sbox(t0) if force_z is None else force_z
    zh = halves(z)
    for h in range(2):
        mz = Fraction(2**52 + zh[h])
        gg = dbl(dbl(mz - tb[h]) + 8 * y0h[h])
        for r in range(6):
            Eo[h][r] = dbl(Eo[h][r] + gg * (cc[r] + cc[r + 6]))
            Fo[h][r] = dbl(Fo[h][r] + gg * (cc[r] - cc[r + 6]))
        z8 = dbl(8 * dbl(mz - 2**52))
        Eo[h][0] = dbl(Eo[h][0] + z8); Fo[h][0] = dbl(Fo[h][0] + z8)
    if last:
        ex = (2**52 + EXIT_LO, 2**52 + EXIT_HI)
        nat = [[None] * 12 for _ in range(2)]
        for h in range(2):
            for r in range(6):
                he = dbl(Eo[h][r] / 2)
                nat[h][r] = dbl(he + Fo[h][r] / 2)
                nat[h][r + 6] = dbl(dbl(he - Fo[h][r] / 2) + ex[h])
        return [readout(nat[0][r], nat[1][r]) for r in range(12)]
    y0acc = [dbl(dbl(Eo[h][0] / 2) + Fo[h][0] / 2) for h in range(2)]
    return ([readout(Eo[0][r], Eo[1][r]) for r in range(6)] + [readout(y0acc[0], y0acc[1])] +
            [readout(Fo[0][r], Fo[1][r]) for r in range(1, 6)])


for trial in range(12):
    if trial == 0: s = [P - 1] * 12
    elif trial == 1: s = [0] * 12
    elif trial == 2: s = [0xFFFFFFFFFFFFFFFF] * 6 + [0] * 6
    elif trial == 3: s = [0] * 6 + [0xFFFFFFFFFFFFFFFF] * 6
    elif trial == 4: s = [0xFFFFFFFF] * 12
    else: s = [random.randrange(P) for _ in range(12)]
    ref = [v % P for v in s]
    # poseidon_to_pair_basis: the natural state enters the basis once, in integer arithmetic
    st = [(s[j] + s[j + 6]) % P for j in range(6)] + [s[0] % P] + [(s[j] - s[j + 6]) % P for j in range(1, 6)]
    for p in range(11):
        kA = 4 + 2 * p
        for k in (kA, kA + 1):
            ref[0] = sbox(ref[0])
            ref = [(a + b) % P for a, b in zip(mds(ref), rc_true[12 * (k + 1):12 * (k + 2)])]
        st = pair_basis_emulated(st, p, False, p == 10)
        if p < 10:
            e, x0, f = st[:6], st[6], [None] + st[7:]
            assert x0 % P == ref[0] and all(e[j] % P == (ref[j] + ref[j + 6]) % P for j in range(6))
            assert all(f[j] % P == (ref[j] - ref[j + 6]) % P for j in range(1, 6)), (trial, p)
    assert [v % P for v in st] == ref, trial
# worst-case magnitudes inside the chain: basis words forced to extreme representatives
U = 0xFFFFFFFFFFFFFFFF
for pat in ([U] * 12, [0] * 12, [U] * 6 + [0] * 6, [0] * 7 + [U] * 5, [U if j % 2 else 0 for j in range(12)], [U, 0, 0, U, 0, 0, U] + [0, U, 0, U, 0]):
    for p in range(1, 10):
        pair_basis_emulated(list(pat), p, False, False)
    pair_basis_emulated(list(pat), 10, False, True)
    pair_basis_emulated(list(pat), 0, False, False)
print("basis pair tables ok")
# random extreme representatives (each basis word 0, 2^64 - 1, low half only, high half only) through every pair
EXT = [0, U, 0xFFFFFFFF, 0xFFFFFFFF00000000]
for trial in range(int(os.environ.get("P2G_GEN_TRIALS", "3000"))):
    pat = [random.choice(EXT) for _ in range(12)]
    p = random.randrange(0, 11)
    try:
        pair_basis_emulated(list(pat), p, False, p == 10, random.choice(EXT + [None]), random.choice(EXT + [None]))
        if 0 < p < 10:
            pair_basis_emulated(list(pat), p, False, False, random.choice(EXT), random.choice(EXT))
    except AssertionError:
        print("range violation", p, [hex(v) for v in pat])
        raise
print("basis pair extremes ok")
