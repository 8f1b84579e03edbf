// Poseidon-Goldilocks width-12 permutation (x^7, 4+22+4 rounds) for host and device.
// Replaces plonky2/src/hash/{poseidon.rs,poseidon_goldilocks.rs,hashing.rs} of the pinned
// dependency (/root/reference/Cargo.toml:12; entered from e.g.
// /root/reference/aes-gcm/src/circuit_gcm.rs:781 `data.prove(pw)` and directly from
// /root/reference/feistel/src/lib.rs:106 `hash_n_to_hash_no_pad`).
//
// Device design (one permutation per thread, everything in registers; DESIGN.md section 3 has the measurements):
//  * the permutation is bound by instruction issue: on B200 an FP64 instruction and an IMAD.WIDE hold the issue port
//    of a sub-partition for two cycles, IMAD.WIDE occupies the fmaheavy pipe for four
//    (profiles/r2_issue_rate_microbench.jsonl), so the design minimises issue slots;
//  * S-box: x^2 and x^4 as three-product squarings, x^3 and x^7 on the compiler's 128-bit product (four IMAD.WIDE
//    whose 64-bit addend and carry absorb the partial-product additions), each followed by the Goldilocks fold
//    hi*2^64 = hi_lo*(2^32-1) - hi_hi  as a 12-instruction carry chain;
//  * linear layers on the otherwise idle FP64 pipe: the 32-bit halves of the state are exact doubles, the circulant
//    coefficients are small, every partial sum stays below 2^53, and the accumulators start at 2^52 + round
//    constant so the integer result is read straight from the mantissa words.  The circulant splits twice
//    (12 -> 6 + 6 -> 3 + 3 + 6): 184 FP64 instructions per full round instead of 290;
//  * the 22 partial rounds run as 11 merged pairs: u = circ(C*C) s' + cc (8 y0 + z - t0) + 8 (z - cA0) e0 + K, the same
//    split product plus one rank-one term, 236 FP64 instructions and 13 accumulator read-outs per TWO rounds;
//  * intermediate values are lazy residues (any u64); only outputs are canonicalised;
//  * the 4+4 full rounds share one loop body and the 11 pairs another, keeping the hot code inside the
//    instruction cache.
#pragma once
#include "gl64.cuh"

static const gl_t POSEIDON_RC_HOST[360] = {
#include "poseidon_rc.inc"
};

#if defined(__CUDACC__)
__constant__ gl_t POSEIDON_RC_DEV[372] = {      // 30 rounds + one all-zero row ("next" of the last round)
#include "poseidon_rc.inc"
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0
};
__device__ gl_t POSEIDON_RC_GLOBAL[372] = {     // same table in global memory for per-lane reads
#include "poseidon_rc.inc"
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0
};
#include "poseidon_rc_f64.inc"
#endif

#if defined(__CUDA_ARCH__)
// ------------------------------------------------------------------------------------------
// device implementation
// ------------------------------------------------------------------------------------------
// a^2: the cross product a0*a1 is computed once and added twice (one IMAD.WIDE fewer)
__device__ __forceinline__ gl_t psqr(gl_t a) {
    uint32_t a0, a1; gl_unpack(a, a0, a1);
    uint32_t l0, c0, ml, mh, p11l, p11h;
    gl_unpack(gl_mulw(a0, a0), l0, c0); gl_unpack(gl_mulw(a0, a1), ml, mh); gl_unpack(gl_mulw(a1, a1), p11l, p11h);
    uint32_t l1, h0, h1;
    asm("add.cc.u32 %0, %3, %4;\n\taddc.cc.u32 %1, %5, %6;\n\taddc.u32 %2, %7, 0;\n\t"
        "add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.u32 %2, %2, 0;"
        : "=&r"(l1), "=&r"(h0), "=&r"(h1)
        : "r"(c0), "r"(ml), "r"(p11l), "r"(mh), "r"(p11h));
    return gl_fold4(l0, l1, h0, h1);
}
// any u64 * any u64 -> lazy residue (about 21 SASS instructions, 5 of them IMAD.WIDE)
__device__ __forceinline__ gl_t pmul(gl_t a, gl_t b) {
    uint32_t l0, l1, h0, h1; pmul128(a, b, l0, l1, h0, h1);
    return gl_fold4(l0, l1, h0, h1);
}
// a * b + c (all lazy) -> lazy residue; the sum still fits 128 bits
__device__ __forceinline__ gl_t pmul_add(gl_t a, gl_t b, gl_t c) {
    uint32_t l0, l1, h0, h1, c0, c1; pmul128(a, b, l0, l1, h0, h1); gl_unpack(c, c0, c1);
    asm("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.u32 %3, %3, 0;"
        : "+r"(l0), "+r"(l1), "+r"(h0), "+r"(h1) : "r"(c0), "r"(c1));
    return gl_fold4(l0, l1, h0, h1);
}
// 160-bit accumulator for sums of up to 2^32 128-bit products
struct Acc160 { uint32_t w[5]; };
__device__ __forceinline__ void acc_mul(Acc160& A, gl_t a, gl_t b) {
    uint32_t l0, l1, h0, h1; pmul128(a, b, l0, l1, h0, h1);
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, %4, 0;"
        : "+r"(A.w[0]), "+r"(A.w[1]), "+r"(A.w[2]), "+r"(A.w[3]), "+r"(A.w[4]) : "r"(l0), "r"(l1), "r"(h0), "r"(h1));
}
__device__ __forceinline__ gl_t acc_fold(const Acc160& A) { return gl_fold5(A.w[0], A.w[1], A.w[2], A.w[3], A.w[4]); }
// P2G_DIAG_NO_SBOX / P2G_DIAG_NO_MDS: timing diagnostics only (tools/poseidon_overlap.sh) -- they remove the
// integer or the FP64 half of every round to measure how much the two halves overlap; results are wrong.
// S-box multiply, variant 2 (P2G_SBOX_V): the 128-bit product is left to the compiler -- ptxas turns
// mul.lo.u64 + mul.hi.u64 into four IMAD.WIDE whose 64-bit addend and carry (IMAD.WIDE.U32.X) absorb the
// partial-product additions, 7 instructions instead of 4 + 6 -- and the fold runs through one more wide
// MAD:  t = lo + h0 * (2^32 - 1)  (65 bits),  t -= h1,  then the top word k in {-1, 0, 1} is folded back as
// k * (2^32 - 1).  73 instead of 83 instructions per S-box.  Lazy in, lazy out.
__device__ __forceinline__ gl_t pmul_v2(gl_t a, gl_t b) {
    typedef unsigned __int128 u128;
    const u128 p = (u128)a * b;
    const uint64_t lo = (uint64_t)p, hi = (uint64_t)(p >> 64);
    u128 t = (u128)(uint32_t)hi * 0xFFFFFFFFull + lo;
    t -= (uint32_t)(hi >> 32);
    const uint64_t k = (uint64_t)(t >> 64);              // 0, 1 or 2^64 - 1
    return (uint64_t)t + (k << 32) - k;                   // cannot wrap: see the bounds in DESIGN.md
}
__device__ __forceinline__ gl_t pmul_v1(gl_t a, gl_t b) {
    typedef unsigned __int128 u128;
    const u128 p = (u128)a * b;
    uint32_t l0, l1, h0, h1; gl_unpack((uint64_t)p, l0, l1); gl_unpack((uint64_t)(p >> 64), h0, h1);
    return gl_fold4(l0, l1, h0, h1);
}
#ifndef P2G_SBOX_V
#define P2G_SBOX_V 3
#endif
__device__ __forceinline__ gl_t poseidon_sbox(gl_t x) {
#ifdef P2G_DIAG_NO_SBOX
    return x + 1;
#elif P2G_SBOX_V == 2
    gl_t x2 = pmul_v2(x, x), x4 = pmul_v2(x2, x2), x3 = pmul_v2(x, x2);
    return pmul_v2(x3, x4);
#elif P2G_SBOX_V == 1
    gl_t x2 = pmul_v1(x, x), x4 = pmul_v1(x2, x2), x3 = pmul_v1(x, x2);
    return pmul_v1(x3, x4);
#elif P2G_SBOX_V == 3
    gl_t x2 = psqr(x), x4 = psqr(x2), x3 = pmul_v1(x, x2);   // three-product squarings, compiler product elsewhere
    return pmul_v1(x3, x4);
#else
    gl_t x2 = psqr(x), x4 = psqr(x2), x3 = pmul(x, x2);
    return pmul(x3, x4);
#endif
}
// x^7 with the last product left unfolded: (h1:h0:l1:l0) = x^3 * x^4, value = (l0 - h0 - h1) + 2^32 (l1 + h0) (mod p).
// P2G_SBOX_UNFOLDED=1 lets the full rounds hand these two limbs to the FP64 layer directly (four I2F and three FP64
// additions instead of the 12-instruction fold and two I2F).  Measured: the full-round loop alone 1893 -> 1860 cycles
// per warp, but the whole permutation 1.516 -> 1.497 G/s (register pressure: spills in the round loop), so it is off;
// the bias tables carry the offset that makes it exact either way (tools/gen_poseidon_f64.py).
#ifndef P2G_SBOX_UNFOLDED
#define P2G_SBOX_UNFOLDED 0
#endif
__device__ __forceinline__ void poseidon_sbox_limbs(gl_t x, double& xl, double& xh) {
#if P2G_SBOX_UNFOLDED && !defined(P2G_DIAG_NO_SBOX)
    const gl_t x2 = psqr(x), x4 = psqr(x2), x3 = pmul_v1(x, x2);
    uint32_t l0, l1, h0, h1; pmul128(x3, x4, l0, l1, h0, h1);
    xl = __dsub_rn(__dsub_rn((double)l0, (double)h0), (double)h1);
    xh = __dadd_rn((double)l1, (double)h0);
#else
    const gl_t v = poseidon_sbox(x);
    xl = (double)(uint32_t)v; xh = (double)(uint32_t)(v >> 32);
#endif
}
// s <- MDS * s + RC[next_row]   (lazy in, lazy out) on the FP64 pipe.
// The circulant coefficients are <= 41, so every partial sum of  c_i * (32-bit half)  plus a 32-bit
// constant stays below 2^43: DFMA on integer-valued doubles is exact, and the FP64 pipe (16
// lanes/clk/SMSP on B200, idle otherwise) co-issues with the IADD3 carry chains of the S-boxes,
// which IMAD.WIDE does not (profiles/r1_int_pipes_microbench.jsonl: dfma+iadd3 2.08 cyc/pair,
// imad.wide+iadd3 5.3).  Accumulators start at 2^52 + round constant, so the integer result can
// be read straight out of the mantissa with no conversion instruction.
// SBOX_ALL: every word goes through the S-box before it enters the accumulation (full rounds).
// Otherwise (single partial rounds, used only when P2G_PARTIAL_PAIRS is 0) only word 0 does and its group is
// accumulated LAST, so the DFMAs of the other words hide the latency of that dependent multiply chain.
// accumulators start at 2^52 + constant: the integer sits in the low mantissa bits (an F2I readout
// from plain-constant accumulators was measured slower: 0.96 vs 0.99 G perm/s)
// Raw words of the double, exponent bits included: the offset they add to the state word
// (0x43300000 * 2^32 * (1 + 2^32) once the low and high accumulators are combined) is already
// subtracted from every bias table (tools/gen_poseidon_f64.py, E_READ), which saves the two masks.
#define POS_READ(d, w0, w1) do { w0 = (uint32_t)__double2loint(d); w1 = (uint32_t)__double2hiint(d); } while (0)
#ifndef P2G_MDS_SPLIT
#define P2G_MDS_SPLIT 1
#endif

#if P2G_MDS_SPLIT
// Split-circulant form of the same layer.  The MDS matrix is circ(C) (+ 8 on entry [0][0]), i.e.
// [[A, B], [B, A]] in 6x6 blocks, so with X+ = x_lo + x_hi and X- = x_lo - x_hi (word halves j, j+6)
//     y_lo = ((A+B)/2) X+  +  ((A-B)/2) X-,      y_hi = ((A+B)/2) X+  -  ((A-B)/2) X-
// which is 72 products per 32-bit half instead of 144 (204 FP64-pipe instructions per round
// instead of 290).  For this matrix (A+-B)/2 are integers, so acc+ is biased by 2^52 and acc- is a plain signed
// integer (POSEIDON_RCS_*): acc+ + acc- and acc+ - acc- both land in [2^52, 2^53) as 2^52 + word, two FP64
// instructions per pair of words (a third one re-biased the difference before: -12 per round).  I2F.F64.U32
// stays: building 2^52 + w from the raw words and subtracting 2^52 trades one I2F for a DADD plus 1.5 moves and
// measured slower (profiles/r2_poseidon_magic_lag_experiments.jsonl).
// Second level: acc+ is a 6 x 6 cyclic product of X+ and splits once more, acc+_r = pp_r + pm_r, acc+_{r+3} = pp_r - pm_r
// with U+- = X+_j +- X+_{j+3} (18 + 6 + 6 instead of 36 products and sums per half); acc- is negacyclic and stays.
// One accumulator set per 32-bit half: pp biased by 2^52, pm and am plain signed integers.
__device__ __forceinline__ gl_t pos_readout(double l, double h);
struct PosSplitAcc { double pp[3], pm[3], am[6]; };
__device__ __forceinline__ void pos_split_load(PosSplitAcc& A, const double* __restrict__ tab) {
#pragma unroll
    for (int r = 0; r < 3; r++) { A.pp[r] = tab[r]; A.pm[r] = tab[3 + r]; }
#pragma unroll
    for (int r = 0; r < 6; r++) A.am[r] = tab[6 + r];
}
// words j, j + 6, j + 3, j + 9 (one half each) of the vector that circ(CX) multiplies; j < 3
#define POS_SPLIT_GROUP(A, CX, j, xa, xb, xc, xd) do {                                                         \
        const double xp0 = __dadd_rn(xa, xb), xm0 = __dsub_rn(xa, xb), xp1 = __dadd_rn(xc, xd), xm1 = __dsub_rn(xc, xd); \
        _Pragma("unroll")                                                                                         \
        for (int r = 0; r < 6; r++) {                                                                             \
            (A).am[r] = __fma_rn(xm0, 0.5 * (CX[((j) - r + 12) % 12] - CX[((j) + 6 - r + 12) % 12]), (A).am[r]);   \
            (A).am[r] = __fma_rn(xm1, 0.5 * (CX[((j) + 3 - r + 12) % 12] - CX[((j) + 9 - r + 12) % 12]), (A).am[r]); \
        }                                                                                                         \
        const double up = __dadd_rn(xp0, xp1), um = __dsub_rn(xp0, xp1);                                          \
        _Pragma("unroll")                                                                                         \
        for (int r = 0; r < 3; r++) {                                                                             \
            const double pa = 0.5 * (CX[((j) - r + 12) % 12] + CX[((j) + 6 - r + 12) % 12]);                      \
            const double pb = 0.5 * (CX[((j) + 3 - r + 12) % 12] + CX[((j) + 9 - r + 12) % 12]);                  \
            (A).pp[r] = __fma_rn(up, 0.5 * (pa + pb), (A).pp[r]); (A).pm[r] = __fma_rn(um, 0.5 * (pa - pb), (A).pm[r]); \
        }                                                                                                         \
    } while (0)
// y[r], y[r + 6] = acc+ +- acc-, each 2^52 + word
__device__ __forceinline__ void pos_split_recombine(const PosSplitAcc& A, double y[12]) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const double a0 = __dadd_rn(A.pp[r], A.pm[r]), a1 = __dsub_rn(A.pp[r], A.pm[r]);
        y[r] = __dadd_rn(a0, A.am[r]); y[r + 6] = __dsub_rn(a0, A.am[r]);
        y[r + 3] = __dadd_rn(a1, A.am[r + 3]); y[r + 9] = __dsub_rn(a1, A.am[r + 3]);
    }
}
template <bool SBOX_ALL>
__device__ __forceinline__ void poseidon_round(gl_t s[12], int next_row) {
    const double C[12] = {17., 15., 41., 16., 2., 28., 13., 13., 39., 18., 34., 20.};
    PosSplitAcc L, H;
    pos_split_load(L, POSEIDON_RCS_LO + 12 * next_row); pos_split_load(H, POSEIDON_RCS_HI + 12 * next_row);
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
        const int j = SBOX_ALL ? jj : (jj + 1) % 3;             // words (0, 6, 3, 9) last in partial rounds
        double xl[4], xh[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int w = j + 6 * (u & 1) + 3 * (u >> 1);        // j, j + 6, j + 3, j + 9
            if (SBOX_ALL || w == 0) poseidon_sbox_limbs(s[w], xl[u], xh[u]);
            else { xl[u] = (double)(uint32_t)s[w]; xh[u] = (double)(uint32_t)(s[w] >> 32); }
        }
#ifdef P2G_DIAG_NO_MDS
        L.am[j] = __dadd_rn(L.am[j], __dadd_rn(__dadd_rn(xl[0], xl[1]), __dadd_rn(xl[2], xl[3])));
        H.am[j] = __dadd_rn(H.am[j], __dadd_rn(__dadd_rn(xh[0], xh[1]), __dadd_rn(xh[2], xh[3])));
#else
        POS_SPLIT_GROUP(L, C, j, xl[0], xl[1], xl[2], xl[3]);
        POS_SPLIT_GROUP(H, C, j, xh[0], xh[1], xh[2], xh[3]);
#endif
        if (j == 0) {                                            // + 8 x_0 on row 0 only: 4 x_0 to acc+_0 (2 + 2) and to acc-_0
            L.pp[0] = __fma_rn(xl[0], 2., L.pp[0]); L.pm[0] = __fma_rn(xl[0], 2., L.pm[0]); L.am[0] = __fma_rn(xl[0], 4., L.am[0]);
            H.pp[0] = __fma_rn(xh[0], 2., H.pp[0]); H.pm[0] = __fma_rn(xh[0], 2., H.pm[0]); H.am[0] = __fma_rn(xh[0], 4., H.am[0]);
        }
    }
    // all coefficients are integers for this matrix, so pp (biased by 2^52), pm and acc- (signed) hold integers and
    // acc+ + acc-, acc+ - acc- are both 2^52 + word (POSEIDON_RCS_*, tools/gen_poseidon_f64.py)
    double yl[12], yh[12];
    pos_split_recombine(L, yl); pos_split_recombine(H, yh);
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = pos_readout(yl[r], yh[r]);
}
#else
template <bool SBOX_ALL>
__device__ __forceinline__ void poseidon_round(gl_t s[12], int next_row) {
    const double C[12] = {17., 15., 41., 16., 2., 28., 13., 13., 39., 18., 34., 20.};
    double al[12], ah[12];
#pragma unroll
    for (int r = 0; r < 12; r++) { al[r] = POSEIDON_RCD_LO[12 * next_row + r]; ah[r] = POSEIDON_RCD_HI[12 * next_row + r]; }
#pragma unroll
    for (int ii = 0; ii < 12; ii++) {
        const int i = SBOX_ALL ? ii : (ii + 1) % 12;            // 1, 2, ..., 11, 0
        const gl_t v = (SBOX_ALL || i == 0) ? poseidon_sbox(s[i]) : s[i];
        const double xl = (double)(uint32_t)v, xh = (double)(uint32_t)(v >> 32);   // I2F: conversion pipe
#pragma unroll
        for (int r = 0; r < 12; r++) {
            al[r] = __fma_rn(xl, C[(i - r + 12) % 12], al[r]);
            ah[r] = __fma_rn(xh, C[(i - r + 12) % 12], ah[r]);
        }
        if (i == 0) { al[0] = __fma_rn(xl, 8., al[0]); ah[0] = __fma_rn(xh, 8., ah[0]); }
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        // accumulators are 2^52 + integer (< 2^43): the integer sits in the low mantissa bits
        uint32_t al0, al1, ah0, ah1;
        POS_READ(al[r], al0, al1); POS_READ(ah[r], ah0, ah1);
        uint32_t m, t;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=&r"(m), "=&r"(t) : "r"(al1), "r"(ah0), "r"(ah1));
        s[r] = gl_fold3(al0, m, t);
    }
}
#endif
// MDS only (used by the PoseidonGate constraint evaluator, where the S-box inputs are wires)
__device__ __forceinline__ void poseidon_mds_rc(gl_t s[12], int next_row) {
    const double C[12] = {17., 15., 41., 16., 2., 28., 13., 13., 39., 18., 34., 20.};
    double al[12], ah[12];
#pragma unroll
    for (int r = 0; r < 12; r++) { al[r] = POSEIDON_RCD_LO[12 * next_row + r]; ah[r] = POSEIDON_RCD_HI[12 * next_row + r]; }
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const double xl = (double)(uint32_t)s[i], xh = (double)(uint32_t)(s[i] >> 32);
#pragma unroll
        for (int r = 0; r < 12; r++) {
            al[r] = __fma_rn(xl, C[(i - r + 12) % 12], al[r]);
            ah[r] = __fma_rn(xh, C[(i - r + 12) % 12], ah[r]);
        }
        if (i == 0) { al[0] = __fma_rn(xl, 8., al[0]); ah[0] = __fma_rn(xh, 8., ah[0]); }
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        uint32_t al0, al1, ah0, ah1;
        POS_READ(al[r], al0, al1); POS_READ(ah[r], ah0, ah1);
        uint32_t m, t;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=&r"(m), "=&r"(t) : "r"(al1), "r"(ah0), "r"(ah1));
        s[r] = gl_fold3(al0, m, t);
    }
}
// 2^52-biased (low, high) accumulator pair -> lazy residue
__device__ __forceinline__ gl_t pos_readout(double l, double h) {
    uint32_t al0, al1, ah0, ah1;
    POS_READ(l, al0, al1); POS_READ(h, ah0, ah1);
    uint32_t m, t;
    asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=&r"(m), "=&r"(t) : "r"(al1), "r"(ah0), "r"(ah1));
    return gl_fold3(al0, m, t);
}
#ifndef P2G_PARTIAL_PAIRS
#define P2G_PARTIAL_PAIRS 1
#endif
// always 0, but opaque to the compiler's uniformity analysis
__device__ __forceinline__ int pos_lane_zero() {
    long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
    return (int)((unsigned long long)c >> 63);
}
// Two consecutive partial rounds in one pass over the state.  With s' = (sbox(s0), s1..s11),
//     t = M s' + cA,   u = M (sbox(t0), t1..t11) + cB
// collapses to   u = A s' + col0(M) sbox(t0) + K,   t0 = row0(M) s' + cA_0,
// A = M~ M (M~: M with column 0 zeroed) and K = M~ cA + cB precomputed (tools/gen_poseidon_f64.py,
// which also checks the identity against the plain rounds).  The entries of A stay below 2^15, so
// the products with 32-bit halves are still exact on the FP64 pipe (row sums < 2^49), and the 11
// words that skip the S-box are read out of the accumulators and converted back once per TWO rounds:
// 336 FP64 instructions and 13 readouts per pair instead of 408 and 24.
#ifndef P2G_PAIR_SPLIT
#define P2G_PAIR_SPLIT 1
#endif
#if P2G_PAIR_SPLIT
// Split form of the pair.  From  u = M t' + cB,  t' = t + (z - t0) e0,  t = M s' + cA:
//     u = M^2 s' + (z - t0) col0(M) + (M cA + cB),
// and with M = circ(C) + 8 E00:  M^2 = circ(C2) + 8 col0(circ) e0^T + 8 e0 row0(circ) + 64 E00  (C2 = C cyclically
// convolved with itself), so
//     u = circ(C2) s' + 8 cc y0 + cm (z - t0) + 8 (t0 - cA_0) e0 + K'
//       = circ(C2) s' + cc (8 y0 + z - t0) + 8 (z - cA_0) e0 + K',          cc = col0(circ), cm = col0(M), y0 = s'_0.
// circ(C2) is circulant and splits like the full-round layer (72 products per 32-bit half instead of 144); the
// rank-one term costs 24 products and word 0 two more: 260 FP64 instructions per pair instead of 336.  The
// constants, the read-out offset and an offset = 0 (mod p) that keeps the accumulators positive although
// 8 y0 + z - t0 can be negative are in POSEIDON_PAIRS_* (tools/gen_poseidon_f64.py, which emulates this function
// exactly).  z needs no I2F: (2^52 + z_half) is built from the raw words and (2^52 + z_half) - (biased t
// accumulator) is exact.
__device__ __forceinline__ void poseidon_partial_pair(gl_t s[12], int pair, int zero) {
    const double C[12] = {17., 15., 41., 16., 2., 28., 13., 13., 39., 18., 34., 20.};
    const double C2[12] = POSEIDON_C2_INIT;
    const double BIAS = 4503599627370496.0;               // 2^52
    PosSplitAcc L, H;
    const int pz = 12 * pair + zero;                      // per-thread looking index: LDC instead of LDCU + moves
    pos_split_load(L, POSEIDON_PAIRS_LO + pz); pos_split_load(H, POSEIDON_PAIRS_HI + pz);
    const int row_a = 5 + 2 * pair;                       // constants between the two rounds
    double tl = POSEIDON_RCD_LO[12 * row_a + zero], th = POSEIDON_RCD_HI[12 * row_a + zero];
    const gl_t y0 = poseidon_sbox(s[0]);
    double y0l8 = 0., y0h8 = 0.;
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
        const int j = (jj + 1) % 3;                       // words (0, 6, 3, 9) last: the S-box chain of word 0 hides behind the others
        double xl[4], xh[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int w = j + 6 * (u & 1) + 3 * (u >> 1);  // j, j + 6, j + 3, j + 9
            const gl_t v = w == 0 ? y0 : s[w];
            xl[u] = (double)(uint32_t)v; xh[u] = (double)(uint32_t)(v >> 32);
            const double m0 = C[w] + (w == 0 ? 8. : 0.);  // row 0 of M
            tl = __fma_rn(xl[u], m0, tl); th = __fma_rn(xh[u], m0, th);
        }
#ifdef P2G_DIAG_NO_MDS
        L.am[j] = __dadd_rn(L.am[j], __dadd_rn(__dadd_rn(xl[0], xl[1]), __dadd_rn(xl[2], xl[3])));
        H.am[j] = __dadd_rn(H.am[j], __dadd_rn(__dadd_rn(xh[0], xh[1]), __dadd_rn(xh[2], xh[3])));
#else
        POS_SPLIT_GROUP(L, C2, j, xl[0], xl[1], xl[2], xl[3]);
        POS_SPLIT_GROUP(H, C2, j, xh[0], xh[1], xh[2], xh[3]);
#endif
        if (j == 0) { y0l8 = xl[0]; y0h8 = xh[0]; }
    }
    const gl_t z0 = poseidon_sbox(pos_readout(tl, th));
    // while that S-box runs: recombine the split accumulators
    double yl[12], yh[12];
    pos_split_recombine(L, yl); pos_split_recombine(H, yh);
    const double mzl = __hiloint2double(0x43300000, (int)(uint32_t)z0), mzh = __hiloint2double(0x43300000, (int)(uint32_t)(z0 >> 32));
    const double gl = __fma_rn(y0l8, 8., __dsub_rn(mzl, tl)), gh = __fma_rn(y0h8, 8., __dsub_rn(mzh, th));   // 8 y0 + z - t0 (+ E)
#pragma unroll
    for (int r = 0; r < 12; r++) {
        const double cr = C[(12 - r) % 12];                                   // cc[r]
        yl[r] = __fma_rn(gl, cr, yl[r]); yh[r] = __fma_rn(gh, cr, yh[r]);
    }
    yl[0] = __fma_rn(__dsub_rn(mzl, BIAS), 8., yl[0]); yh[0] = __fma_rn(__dsub_rn(mzh, BIAS), 8., yh[0]);
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = pos_readout(yl[r], yh[r]);
}
#else
__device__ __forceinline__ void poseidon_partial_pair(gl_t s[12], int pair, int zero) {
    const double C[12] = {17., 15., 41., 16., 2., 28., 13., 13., 39., 18., 34., 20.};
    const double A[144] = POSEIDON_PAIR_A_INIT;
    double al[12], ah[12];
    // pos_lane_zero(): the table index is made to look per-thread so the biases arrive by LDC straight in
    // vector registers; a uniform index is loaded to uniform registers and then costs two moves per
    // double (52 of the 753 instructions of this loop body)
    const int pz = 12 * pair + zero;
#pragma unroll
    for (int r = 0; r < 12; r++) { al[r] = POSEIDON_PAIRK_LO[pz + r]; ah[r] = POSEIDON_PAIRK_HI[pz + r]; }
    const int row_a = 5 + 2 * pair;                       // constants between the two rounds
    double tl = POSEIDON_RCD_LO[12 * row_a + zero], th = POSEIDON_RCD_HI[12 * row_a + zero];
    const gl_t y0 = poseidon_sbox(s[0]);
#pragma unroll
    for (int jj = 0; jj < 12; jj++) {
        const int j = (jj + 1) % 12;                      // word 0 last: its S-box chain hides behind the others
        const gl_t v = j == 0 ? y0 : s[j];
        const double xl = (double)(uint32_t)v, xh = (double)(uint32_t)(v >> 32);
#ifdef P2G_DIAG_NO_MDS
        al[j] = __dadd_rn(al[j], xl); ah[j] = __dadd_rn(ah[j], xh);
#else
#pragma unroll
        for (int r = 0; r < 12; r++) {
            al[r] = __fma_rn(xl, A[12 * r + j], al[r]);
            ah[r] = __fma_rn(xh, A[12 * r + j], ah[r]);
        }
#endif
        const double m0j = C[j] + (j == 0 ? 8. : 0.);
        tl = __fma_rn(xl, m0j, tl); th = __fma_rn(xh, m0j, th);
    }
    const gl_t z0 = poseidon_sbox(pos_readout(tl, th));
    const double zl = (double)(uint32_t)z0, zh = (double)(uint32_t)(z0 >> 32);
#pragma unroll
    for (int r = 0; r < 12; r++) {
        const double mr0 = C[(12 - r) % 12] + (r == 0 ? 8. : 0.);
        al[r] = __fma_rn(zl, mr0, al[r]); ah[r] = __fma_rn(zh, mr0, ah[r]);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = pos_readout(al[r], ah[r]);
}
#endif  // P2G_PAIR_SPLIT
// P2G_PAIR_BASIS=1: experiment, off.  Measured (profiles/r2_poseidon_pair_basis_experiment.txt): chained permutations
// 1.512 -> 1.553 G/s with the basis entered in FP64 under a `first` flag (but 68 bytes of spills in the Merkle leaf
// kernel: its time unchanged, 3.40 ms), 1.520 G/s in this form (integer entry, no spills: leaf kernel 3.40 ms as well),
// 1.447 G/s with the last pair as a second instantiation of the body (instruction fetch).  The proof rate did not move.
#ifndef P2G_PAIR_BASIS
#define P2G_PAIR_BASIS 0
#endif
#if P2G_PAIR_BASIS && P2G_PAIR_SPLIT
// The pair with its state kept in the (E, F) basis between pairs: E_j = u_j + u_{j+6}, F_j = u_j - u_{j+6} (j < 6).
// The split accumulators of a pair are exactly E / 2 and F / 2 of its output, so with all coefficients doubled
// (free) a pair reads (E, F) and writes (E', F'): neither the recombination acc+ +- acc- at the end of a pair nor the
// X+- formation at the start of the next is needed -- 198 instead of 236 FP64 instructions.  Word 0, the only one that
// meets an S-box, is (E'_0 + F'_0) / 2, exact on the accumulators (every bias is even); inside a pair it becomes
// E_0 = y0 + x6, F_0 = y0 - x6 with x6 = E_0 - x0.  Between pairs s[] holds [e_0..e_5, x0, f_1..f_5]
// (poseidon_to_pair_basis converts the natural state once, in integer arithmetic); the last pair writes the natural
// state (a uniform branch: two instantiations of this body measured 1.45 instead of 1.55 G perm/s, instruction fetch).  Tables and an exact emulation of this function: tools/gen_poseidon_f64.py (POSEIDON_PAIRB_*).
// natural state -> [e_0..e_5, x0, f_1..f_5], e_j = x_j + x_{j+6}, f_j = x_j - x_{j+6} (lazy residues)
__device__ __forceinline__ gl_t gl_add_lazy_dev(gl_t a, gl_t c);
__device__ __forceinline__ void poseidon_to_pair_basis(gl_t s[12]) {
    const gl_t x0 = s[0];
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const gl_t b = gl_canon(s[j + 6]);
        const gl_t e = gl_add_lazy_dev(s[j], b), f = gl_sub(s[j], b);
        s[j] = e; s[j + 6] = f;
    }
    s[6] = x0;
}
__device__ __forceinline__ void poseidon_partial_pair_basis(gl_t s[12], int pair, int zero, bool last) {
    const double C[12] = {17., 15., 41., 16., 2., 28., 13., 13., 39., 18., 34., 20.};
    const double C2[12] = POSEIDON_C2_INIT;
    const double BIAS = 4503599627370496.0;               // 2^52
    double ppl[3], pml[3], fal[6], pph[3], pmh[3], fah[6];
    const int pz = 12 * pair + zero;                      // per-thread looking index: LDC instead of LDCU + moves
#pragma unroll
    for (int r = 0; r < 3; r++) {
        ppl[r] = POSEIDON_PAIRB_LO[pz + r]; pml[r] = POSEIDON_PAIRB_LO[pz + 3 + r];
        pph[r] = POSEIDON_PAIRB_HI[pz + r]; pmh[r] = POSEIDON_PAIRB_HI[pz + 3 + r];
    }
#pragma unroll
    for (int r = 0; r < 6; r++) { fal[r] = POSEIDON_PAIRB_LO[pz + 6 + r]; fah[r] = POSEIDON_PAIRB_HI[pz + 6 + r]; }
    double tl = POSEIDON_PAIRB_T_LO[pair + zero], th = POSEIDON_PAIRB_T_HI[pair + zero];
    const gl_t y0 = poseidon_sbox(s[6]);
    const double y0l = (double)(uint32_t)y0, y0h = (double)(uint32_t)(y0 >> 32);
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
        const int j = (jj + 1) % 3;                       // group (0, 3) last: the S-box chain of word 0 hides behind the others
        // e_j, f_j (x0 for j = 0), e_{j+3}, f_{j+3}
        double e0l = (double)(uint32_t)s[j], e0h = (double)(uint32_t)(s[j] >> 32);
        double f0l = (double)(uint32_t)s[j + 6], f0h = (double)(uint32_t)(s[j + 6] >> 32);
        const double e1l = (double)(uint32_t)s[j + 3], e1h = (double)(uint32_t)(s[j + 3] >> 32);
        const double f1l = (double)(uint32_t)s[j + 9], f1h = (double)(uint32_t)(s[j + 9] >> 32);
        if (j == 0) {                                      // e0 = e_0, f0 = x0: x6 = e_0 - x0, then E_0 = y0 + x6, F_0 = y0 - x6
            const double x6l = __dsub_rn(e0l, f0l), x6h = __dsub_rn(e0h, f0h);
            e0l = __dadd_rn(y0l, x6l); f0l = __dsub_rn(y0l, x6l); e0h = __dadd_rn(y0h, x6h); f0h = __dsub_rn(y0h, x6h);
        }
#ifdef P2G_DIAG_NO_MDS
        fal[j] = __dadd_rn(fal[j], __dadd_rn(__dadd_rn(e0l, f0l), __dadd_rn(e1l, f1l)));
        fah[j] = __dadd_rn(fah[j], __dadd_rn(__dadd_rn(e0h, f0h), __dadd_rn(e1h, f1h)));
#else
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const double n0 = C2[(j - r + 12) % 12] - C2[(j + 6 - r + 12) % 12], n1 = C2[(j + 3 - r + 12) % 12] - C2[(j + 9 - r + 12) % 12];
            fal[r] = __fma_rn(f0l, n0, fal[r]); fal[r] = __fma_rn(f1l, n1, fal[r]);
            fah[r] = __fma_rn(f0h, n0, fah[r]); fah[r] = __fma_rn(f1h, n1, fah[r]);
        }
        const double upl = __dadd_rn(e0l, e1l), uml = __dsub_rn(e0l, e1l), uph = __dadd_rn(e0h, e1h), umh = __dsub_rn(e0h, e1h);
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double pa = 0.5 * (C2[(j - r + 12) % 12] + C2[(j + 6 - r + 12) % 12]);
            const double pb = 0.5 * (C2[(j + 3 - r + 12) % 12] + C2[(j + 9 - r + 12) % 12]);
            ppl[r] = __fma_rn(upl, pa + pb, ppl[r]); pml[r] = __fma_rn(uml, pa - pb, pml[r]);
            pph[r] = __fma_rn(uph, pa + pb, pph[r]); pmh[r] = __fma_rn(umh, pa - pb, pmh[r]);
        }
#endif
        // row 0 of M on the (E, F) inputs
        const double m0a = C[j] + (j == 0 ? 8. : 0.), m0b = C[j + 6], m1a = C[j + 3], m1b = C[j + 9];
        tl = __fma_rn(e0l, 0.5 * (m0a + m0b), tl); tl = __fma_rn(f0l, 0.5 * (m0a - m0b), tl);
        tl = __fma_rn(e1l, 0.5 * (m1a + m1b), tl); tl = __fma_rn(f1l, 0.5 * (m1a - m1b), tl);
        th = __fma_rn(e0h, 0.5 * (m0a + m0b), th); th = __fma_rn(f0h, 0.5 * (m0a - m0b), th);
        th = __fma_rn(e1h, 0.5 * (m1a + m1b), th); th = __fma_rn(f1h, 0.5 * (m1a - m1b), th);
    }
    const gl_t z0 = poseidon_sbox(pos_readout(tl, th));
    // while that S-box runs: E'_r = pp_r + pm_r, E'_{r+3} = pp_r - pm_r
    double el[6], eh[6];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        el[r] = __dadd_rn(ppl[r], pml[r]); el[r + 3] = __dsub_rn(ppl[r], pml[r]);
        eh[r] = __dadd_rn(pph[r], pmh[r]); eh[r + 3] = __dsub_rn(pph[r], pmh[r]);
    }
    const double mzl = __hiloint2double(0x43300000, (int)(uint32_t)z0), mzh = __hiloint2double(0x43300000, (int)(uint32_t)(z0 >> 32));
    const double gl = __fma_rn(y0l, 8., __dsub_rn(mzl, tl)), gh = __fma_rn(y0h, 8., __dsub_rn(mzh, th));   // 8 y0 + z - t0 (+ E)
#pragma unroll
    for (int r = 0; r < 6; r++) {
        const double cp = C[(12 - r) % 12] + C[(6 - r) % 12], cn = C[(12 - r) % 12] - C[(6 - r) % 12];       // cc[r] +- cc[r + 6]
        el[r] = __fma_rn(gl, cp, el[r]); fal[r] = __fma_rn(gl, cn, fal[r]);
        eh[r] = __fma_rn(gh, cp, eh[r]); fah[r] = __fma_rn(gh, cn, fah[r]);
    }
    const double zpl = __dsub_rn(mzl, BIAS), zph = __dsub_rn(mzh, BIAS);
    el[0] = __fma_rn(zpl, 8., el[0]); fal[0] = __fma_rn(zpl, 8., fal[0]);
    eh[0] = __fma_rn(zph, 8., eh[0]); fah[0] = __fma_rn(zph, 8., fah[0]);
    double ol[12], oh[12];
    if (last) {                                           // natural state: u_r = (E'_r + F'_r) / 2, u_{r+6} = (E'_r - F'_r) / 2
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const double hl = 0.5 * el[r], hh = 0.5 * eh[r];
            ol[r] = __fma_rn(fal[r], 0.5, hl); ol[r + 6] = __dadd_rn(__fma_rn(fal[r], -0.5, hl), POSEIDON_PAIRB_EXIT_LO);
            oh[r] = __fma_rn(fah[r], 0.5, hh); oh[r + 6] = __dadd_rn(__fma_rn(fah[r], -0.5, hh), POSEIDON_PAIRB_EXIT_HI);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 6; r++) { ol[r] = el[r]; oh[r] = eh[r]; ol[r + 6] = fal[r]; oh[r + 6] = fah[r]; }
        ol[6] = __fma_rn(fal[0], 0.5, 0.5 * el[0]); oh[6] = __fma_rn(fah[0], 0.5, 0.5 * eh[0]);   // word 0 for the next pair's S-box
    }
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = pos_readout(ol[r], oh[r]);
}
#endif
__device__ __forceinline__ gl_t gl_add_lazy_dev(gl_t a, gl_t c) {   // c canonical
    gl_t s = a + c;
    return s < a ? s + GL_EPS : s;
}
// lazy in (any u64), lazy out.  Dense partial rounds: the sparse ("fast") form was measured at
// the same throughput (profiles/r1_poseidon_kernel_ncu_summary.txt: 677 vs 673 M perm/s) and
// costs an 11x11 initial matrix plus 5 KB of constants, so the simpler form is kept.
__device__ __forceinline__ void poseidon_permute_lazy(gl_t s[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_lazy_dev(s[i], POSEIDON_RC_DEV[i]);
    int k = 0;
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
        for (int r = 0; r < 4; r++, k++) {
            poseidon_round<true>(s, k + 1);
        }
        if (phase == 0) {
#if P2G_PARTIAL_PAIRS
            const int zero = pos_lane_zero();
#if P2G_PAIR_BASIS && P2G_PAIR_SPLIT
            poseidon_to_pair_basis(s);
#pragma unroll 1
            for (int p = 0; p < 11; p++, k += 2) poseidon_partial_pair_basis(s, p, zero, p == 10);   // ONE copy of the body: a second one costs more in instruction fetch than the branch
#else
#pragma unroll 1
            for (int p = 0; p < 11; p++, k += 2) poseidon_partial_pair(s, p, zero);
#endif
#else
#pragma unroll 1
            for (int r = 0; r < 22; r++, k++) {
                poseidon_round<false>(s, k + 1);
            }
#endif
        }
    }
}
__device__ __forceinline__ void poseidon_permute(gl_t s[12]) {
    poseidon_permute_lazy(s);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}
// ------------------------------------------------------------------------------------------
// Low-latency form: one permutation spread over 12 lanes of a 16-lane group (2 per warp).
// Lane l holds state word l.  Every round costs one S-box + 22 shuffles + 24 IMAD.WIDE per lane
// instead of 12 S-boxes + a 144-term MDS per thread, so the dependent chain of a permutation is
// ~6x shorter.  Used where a Merkle level has too few nodes to fill the GPU (tree tops, FRI
// layers): there latency, not throughput, is the cost.  rc: global-memory copy of the constants
// (per-lane addresses would serialise in the constant cache).
__device__ __forceinline__ gl_t poseidon_coop(gl_t x, uint32_t l, uint32_t group_base, const gl_t* __restrict__ rc) {
    const uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    const uint32_t lc = l < 12 ? l : 0;
    x = gl_add_lazy_dev(x, __ldg(rc + lc));
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        const bool full = r < 4 || r >= 26;
        gl_t y = poseidon_sbox(x);
        if (full || l == 0) x = y;
        uint32_t lo, hi; gl_unpack(x, lo, hi);
        uint32_t nl, nh; gl_unpack(__ldg(rc + 12 * (r + 1) + lc), nl, nh);     // row 30 is zero
        uint64_t al = nl, ah = nh;
        al += (uint64_t)lo * C[0]; ah += (uint64_t)hi * C[0];
        if (l == 0) { al += (uint64_t)lo * 8u; ah += (uint64_t)hi * 8u; }
#pragma unroll
        for (int i = 1; i < 12; i++) {
            const uint32_t src = group_base + (lc + i >= 12 ? lc + i - 12 : lc + i);
            const uint32_t vl = __shfl_sync(0xffffffffu, lo, src), vh = __shfl_sync(0xffffffffu, hi, src);
            al += (uint64_t)vl * C[i]; ah += (uint64_t)vh * C[i];
        }
        uint32_t al0, al1, ah0, ah1; gl_unpack(al, al0, al1); gl_unpack(ah, ah0, ah1);
        uint32_t m, t;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=&r"(m), "=&r"(t) : "r"(al1), "r"(ah0), "r"(ah1));
        x = gl_fold3(al0, m, t);
    }
    return x;
}
#else
// ------------------------------------------------------------------------------------------
// host implementation (transcript only: a few hundred permutations per proof)
// ------------------------------------------------------------------------------------------
// Host permutation (Fiat-Shamir transcript: ~170 permutations per proof, on the critical path
// between kernels).  Branch-free reductions, MDS on 32-bit halves, sparse partial rounds.
namespace p2g_host {
#define PFAST_QUAL static const
#include "poseidon_fast.inc"
#undef PFAST_QUAL
inline gl_t red128(unsigned __int128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t0 = lo - hh; t0 -= GL_EPS & (0 - (uint64_t)(lo < hh));
    uint64_t t1 = hl * GL_EPS, t2 = t0 + t1;
    t2 += GL_EPS & (0 - (uint64_t)(t2 < t1));
    t2 -= GL_P & (0 - (uint64_t)(t2 >= GL_P));
    return t2;
}
inline gl_t mul(gl_t a, gl_t b) { return red128((unsigned __int128)a * b); }
inline gl_t add(gl_t a, gl_t b) { uint64_t s = a + b; return s - (GL_P & (0 - ((uint64_t)(s < a) | (uint64_t)(s >= GL_P)))); }
inline gl_t sbox(gl_t x) { gl_t x2 = mul(x, x), x4 = mul(x2, x2), x3 = mul(x, x2); return mul(x3, x4); }
inline void mds(gl_t s[12]) {
    static const uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[24], hi[24];
    for (int i = 0; i < 12; i++) { lo[i] = lo[i + 12] = (uint32_t)s[i]; hi[i] = hi[i + 12] = (uint32_t)(s[i] >> 32); }
    uint64_t al[12] = {0}, ah[12] = {0};
    for (int i = 0; i < 12; i++)
        for (int r = 0; r < 12; r++) { al[r] += (uint64_t)lo[i + r] * C[i]; ah[r] += (uint64_t)hi[i + r] * C[i]; }
    al[0] += (uint64_t)lo[0] * 8; ah[0] += (uint64_t)hi[0] * 8;
    for (int r = 0; r < 12; r++) s[r] = red128((unsigned __int128)al[r] + ((unsigned __int128)ah[r] << 32));
}
inline gl_t dot11(const gl_t* a, const gl_t* b, unsigned __int128 init) {
    unsigned __int128 accl = init, acch = 0;
    for (int j = 0; j < 11; j++) {
        accl += (unsigned __int128)a[j] * (uint32_t)b[j];
        acch += (unsigned __int128)a[j] * (uint32_t)(b[j] >> 32);
    }
    return add(red128(accl), mul(red128(acch), (gl_t)1 << 32));
}
}  // namespace p2g_host
inline void poseidon_permute(gl_t s[12]) {
    using namespace p2g_host;
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) s[i] = sbox(add(s[i], POSEIDON_RC_HOST[12 * r + i]));
        mds(s);
    }
    for (int i = 0; i < 12; i++) s[i] = add(s[i], PFAST_FIRST_C[i]);
    {
        gl_t t[11];
        for (int r = 0; r < 11; r++) t[r] = dot11(PFAST_INIT + r * 11, s + 1, 0);
        for (int r = 0; r < 11; r++) s[r + 1] = t[r];
    }
    for (int r = 0; r < 22; r++) {
        gl_t t = add(sbox(s[0]), PFAST_K[r]);
        gl_t s0 = dot11(PFAST_VROW + r * 11, s + 1, (unsigned __int128)t * 25);
        for (int j = 0; j < 11; j++) s[j + 1] = add(s[j + 1], mul(PFAST_WCOL[r * 11 + j], t));
        s[0] = s0;
    }
    for (int r = 26; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = sbox(add(s[i], POSEIDON_RC_HOST[12 * r + i]));
        mds(s);
    }
}
inline void poseidon_permute_lazy(gl_t s[12]) { poseidon_permute(s); }
#endif

// two_to_one(l, r): Poseidon([l, r, 0, 0, 0, 0])[0..4]
GL_HD void poseidon_two_to_one(const gl_t l[4], const gl_t r[4], gl_t out[4]) {
    gl_t s[12];
#pragma unroll
    for (int i = 0; i < 4; i++) { s[i] = l[i]; s[4 + i] = r[i]; s[8 + i] = 0; }
    poseidon_permute_lazy(s);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl_canon(s[i]);
}
