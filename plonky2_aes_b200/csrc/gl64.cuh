// Goldilocks field GF(2^64 - 2^32 + 1) and its quadratic extension F[X]/(X^2-7) for sm_100a.
// Montgomery-free: a 128-bit product is folded with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p).
// Replaces plonky2's field/src/goldilocks_field.rs + goldilocks_extensions.rs (un-vendored
// dependency pinned at /root/reference/Cargo.toml:12) on the device side.
// Convention: every value stored to HBM is canonical (< p); in-register values may be any u64
// where a function says "lazy".
#pragma once
#include <stdint.h>

#define GL_P   0xFFFFFFFF00000001ULL
#define GL_EPS 0x00000000FFFFFFFFULL

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

typedef uint64_t gl_t;

// Device forms are carry-chain PTX: the compiler's compare-and-select versions cost 8-11 SASS
// instructions, most of them on the 16-lane ALU pipe, which is the pipe the NTT saturates
// (profiles/r1_kernels_final_ncu_summary.txt: alu 71 %).
GL_HD gl_t gl_canon(gl_t a) {
#if defined(__CUDA_ARCH__)
    // a >= p  <=>  a + (2^32 - 1) carries out of 64 bits (c = 1); then a - p = (low - 1, 0), and
    // the high word is all ones, so both words are fixed by one independent add of -+c each
    gl_t r;
    asm("{\n\t.reg .u32 a0, a1, t, c;\n\t"
        "mov.b64 {a0,a1}, %1;\n\t"
        "add.cc.u32 t, a0, 0xffffffff;\n\taddc.cc.u32 t, a1, 0;\n\taddc.u32 c, 0, 0;\n\t"
        "sub.u32 a0, a0, c;\n\tadd.u32 a1, a1, c;\n\t"
        "mov.b64 %0, {a0,a1};\n\t}" : "=l"(r) : "l"(a));
    return r;
#else
    return a >= GL_P ? a - GL_P : a;
#endif
}

// canonical - canonical -> canonical  (device: also  any u64 - canonical -> same residue, any u64)
GL_HD gl_t gl_sub(gl_t a, gl_t b) {
#if defined(__CUDA_ARCH__)
    // d = a - b; on borrow add p = 2^64 - 2^32 + 1, i.e. subtract the all-ones borrow mask from the
    // low word and propagate: 5 carry-chain instructions, no compare, no select
    gl_t d;
    asm("{\n\t.reg .u32 a0, a1, b0, b1, m;\n\t"
        "mov.b64 {a0,a1}, %1;\n\tmov.b64 {b0,b1}, %2;\n\t"
        "sub.cc.u32 a0, a0, b0;\n\tsubc.cc.u32 a1, a1, b1;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 a0, a0, m;\n\tsubc.u32 a1, a1, 0;\n\t"
        "mov.b64 %0, {a0,a1};\n\t}" : "=l"(d) : "l"(a), "l"(b));
    return d;
#else
    gl_t d = a - b;
    return (a < b) ? d + GL_P : d;
#endif
}
// canonical + canonical -> canonical
GL_HD gl_t gl_add(gl_t a, gl_t b) {
#if defined(__CUDA_ARCH__)
    // a + b = a - (p - b): two more carry-chain instructions in front of gl_sub (p - 0 = p is fine)
    gl_t d;
    asm("{\n\t.reg .u32 a0, a1, b0, b1, m;\n\t"
        "mov.b64 {a0,a1}, %1;\n\tmov.b64 {b0,b1}, %2;\n\t"
        "sub.cc.u32 b0, 1, b0;\n\tsubc.u32 b1, 0xffffffff, b1;\n\t"
        "sub.cc.u32 a0, a0, b0;\n\tsubc.cc.u32 a1, a1, b1;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 a0, a0, m;\n\tsubc.u32 a1, a1, 0;\n\t"
        "mov.b64 %0, {a0,a1};\n\t}" : "=l"(d) : "l"(a), "l"(b));
    return d;
#else
    gl_t s = a + b;
    return (s < a || s >= GL_P) ? s - GL_P : s;
#endif
}
GL_HD gl_t gl_neg(gl_t a) { return a ? GL_P - a : 0; }

// lazy add: a any u64, b any u64 with at least one of them canonical -> any u64 (same residue)
GL_HD gl_t gl_add_lazy(gl_t a, gl_t b) {
    gl_t s = a + b;
    return (s < a) ? s + GL_EPS : s;
}

// (hi:lo) 128-bit -> u64 residue (not necessarily canonical)
GL_HD gl_t gl_reduce128_lazy(gl_t lo, gl_t hi) {
    gl_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    gl_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;          // borrow: subtract 2^64 = EPS (mod p)
    gl_t t1 = (hi_lo << 32) - hi_lo;        // hi_lo * (2^32 - 1)
    gl_t t2 = t0 + t1;
    if (t2 < t1) t2 += GL_EPS;
    return t2;
}
GL_HD void gl_mul_wide(gl_t a, gl_t b, gl_t& lo, gl_t& hi) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    lo = (gl_t)x; hi = (gl_t)(x >> 64);
#endif
}
#if defined(__CUDACC__)
__constant__ uint32_t GL_EPS_DEV = 0xffffffffu;  // kept in constant memory so ptxas keeps h*EPS as one IMAD.WIDE
#endif
#if defined(__CUDA_ARCH__)
// ---- device multiply on 32-bit halves (the 64-bit mul.hi of the compiler costs 13 issue cycles) ----
__device__ __forceinline__ void gl_unpack(gl_t x, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
__device__ __forceinline__ gl_t gl_pack(uint32_t lo, uint32_t hi) { gl_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ gl_t gl_mulw(uint32_t a, uint32_t b) { gl_t r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }

// (l1:l0) + h0*2^64 -> lazy residue.  h0*2^64 = h0*2^32 - h0: subtract h0 from the low word and add
// it (minus the borrow) to the high word -- two carry-chain instructions instead of an IMAD.WIDE
// whose 64-bit addend is never register-pair aligned here (ptxas split it into IMAD.WIDE + IADD3 +
// IMAD.X).  The 64-bit sum wrapped  <=>  the high word went down: then add 2^64 = EPS once more
// (the sum is below 2^64 + 2^64 - 2^33, so this cannot wrap again).
__device__ __forceinline__ void gl_fold3w(uint32_t l0, uint32_t l1, uint32_t h0, uint32_t& w0, uint32_t& w1) {
    // the wrap test "high word went down" is itself a borrow (w1 - l1), turned into the all-ones mask
    // by subc: carry-chain / IMAD.X instructions instead of ISETP + SEL, which only the ALU pipe runs
    asm("{\n\t.reg .u32 nh, m, t;\n\t"
        "neg.s32 nh, %4;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, nh;\n\t"
        "sub.cc.u32 t, %1, %3;\n\tsubc.u32 m, 0, 0;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "=&r"(w0), "=&r"(w1) : "r"(l0), "r"(l1), "r"(h0));
}
__device__ __forceinline__ gl_t gl_fold3(uint32_t l0, uint32_t l1, uint32_t h0) {
    uint32_t w0, w1; gl_fold3w(l0, l1, h0, w0, w1);
    return gl_pack(w0, w1);
}
// (l1:l0) + h0*2^64 + (h2:h1)*2^96 -> lazy residue  (2^96 = -1: subtract (h2:h1), -EPS on borrow;
// h2 is the small overflow word of multi-term accumulations, 2^128 = -2^32)
__device__ __forceinline__ gl_t gl_fold5(uint32_t l0, uint32_t l1, uint32_t h0, uint32_t h1, uint32_t h2) {
    uint32_t w0, w1, m; gl_fold3w(l0, l1, h0, w0, w1);
    asm("sub.cc.u32 %0, %0, %3;\n\tsubc.cc.u32 %1, %1, %4;\n\tsubc.u32 %2, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\tsubc.u32 %1, %1, 0;"
        : "+r"(w0), "+r"(w1), "=&r"(m) : "r"(h1), "r"(h2));
    return gl_pack(w0, w1);
}
__device__ __forceinline__ gl_t gl_fold4(uint32_t l0, uint32_t l1, uint32_t h0, uint32_t h1) { return gl_fold5(l0, l1, h0, h1, 0); }
// 128-bit product of two u64 as four 32-bit words
// P2G_PMUL_V 1: the product is left to the compiler -- ptxas fuses mul.lo.u64 + mul.hi.u64 into four IMAD.WIDE whose
// 64-bit addend and carry (IMAD.WIDE.U32.X) absorb the partial-product additions: 7 instructions instead of 4 + 6.
#ifndef P2G_PMUL_V
#define P2G_PMUL_V 1
#endif
__device__ __forceinline__ void pmul128(gl_t a, gl_t b, uint32_t& l0, uint32_t& l1, uint32_t& h0, uint32_t& h1) {
#if P2G_PMUL_V == 1
    const unsigned __int128 p = (unsigned __int128)a * b;
    gl_unpack((uint64_t)p, l0, l1); gl_unpack((uint64_t)(p >> 64), h0, h1);
    return;
#endif
    uint32_t a0, a1, b0, b1; gl_unpack(a, a0, a1); gl_unpack(b, b0, b1);
    uint32_t c0, m1l, m1h, m2l, m2h, p11l, p11h;
    gl_unpack(gl_mulw(a0, b0), l0, c0); gl_unpack(gl_mulw(a0, b1), m1l, m1h);
    gl_unpack(gl_mulw(a1, b0), m2l, m2h); gl_unpack(gl_mulw(a1, b1), p11l, p11h);
    asm("add.cc.u32 %0, %3, %4;\n\taddc.cc.u32 %1, %5, %6;\n\taddc.u32 %2, %7, 0;\n\t"
        "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.u32 %2, %2, 0;"
        : "=&r"(l1), "=&r"(h0), "=&r"(h1)
        : "r"(c0), "r"(m1l), "r"(p11l), "r"(m1h), "r"(p11h), "r"(m2l), "r"(m2h));
}
__device__ __forceinline__ gl_t gl_mul_lazy(gl_t a, gl_t b) {
    uint32_t l0, l1, h0, h1; pmul128(a, b, l0, l1, h0, h1);
    return gl_fold4(l0, l1, h0, h1);
}
// a*b + c*d (all any u64) -> lazy residue: the two 128-bit products are added before the single fold
__device__ __forceinline__ gl_t gl_mul2_lazy(gl_t a, gl_t b, gl_t c, gl_t d) {
    uint32_t l0, l1, h0, h1, m0, m1, n0, n1, h2;
    pmul128(a, b, l0, l1, h0, h1); pmul128(c, d, m0, m1, n0, n1);
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, 0, 0;"
        : "+r"(l0), "+r"(l1), "+r"(h0), "+r"(h1), "=r"(h2) : "r"(m0), "r"(m1), "r"(n0), "r"(n1));
    return gl_fold5(l0, l1, h0, h1, h2);
}
// a*b + c (all any u64) -> lazy residue
__device__ __forceinline__ gl_t gl_mad_lazy(gl_t a, gl_t b, gl_t c) {
    uint32_t l0, l1, h0, h1, c0, c1, h2;
    pmul128(a, b, l0, l1, h0, h1); gl_unpack(c, c0, c1);
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\taddc.u32 %4, 0, 0;"
        : "+r"(l0), "+r"(l1), "+r"(h0), "+r"(h1), "=r"(h2) : "r"(c0), "r"(c1));
    return gl_fold5(l0, l1, h0, h1, h2);
}
// a*b + c*d + e*f + g (all any u64) -> lazy residue: three 128-bit products and a word summed in 160 bits, one fold
__device__ __forceinline__ gl_t gl_mad3_lazy(gl_t a, gl_t b, gl_t c, gl_t d, gl_t e, gl_t f, gl_t g) {
    uint32_t l0, l1, h0, h1, m0, m1, n0, n1, p0, p1, q0, q1, g0, g1, h2;
    pmul128(a, b, l0, l1, h0, h1); pmul128(c, d, m0, m1, n0, n1); pmul128(e, f, p0, p1, q0, q1); gl_unpack(g, g0, g1);
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, 0, 0;"
        : "+r"(l0), "+r"(l1), "+r"(h0), "+r"(h1), "=r"(h2) : "r"(m0), "r"(m1), "r"(n0), "r"(n1));
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, %4, 0;"
        : "+r"(l0), "+r"(l1), "+r"(h0), "+r"(h1), "+r"(h2) : "r"(p0), "r"(p1), "r"(q0), "r"(q1));
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\taddc.u32 %4, %4, 0;"
        : "+r"(l0), "+r"(l1), "+r"(h0), "+r"(h1), "+r"(h2) : "r"(g0), "r"(g1));
    return gl_fold5(l0, l1, h0, h1, h2);
}
#else
// any u64 * any u64 -> lazy residue
GL_HD gl_t gl_mul_lazy(gl_t a, gl_t b) {
    gl_t lo, hi;
    gl_mul_wide(a, b, lo, hi);
    return gl_reduce128_lazy(lo, hi);
}
GL_HD gl_t gl_mul2_lazy(gl_t a, gl_t b, gl_t c, gl_t d) {
    return gl_add_lazy(gl_mul_lazy(a, b), gl_canon(gl_mul_lazy(c, d)));
}
GL_HD gl_t gl_mad_lazy(gl_t a, gl_t b, gl_t c) { return gl_add_lazy(gl_mul_lazy(a, b), gl_canon(c)); }
GL_HD gl_t gl_mad3_lazy(gl_t a, gl_t b, gl_t c, gl_t d, gl_t e, gl_t f, gl_t g) {
    return gl_add_lazy(gl_add_lazy(gl_mul_lazy(a, b), gl_canon(gl_mul_lazy(c, d))), gl_canon(gl_add_lazy(gl_mul_lazy(e, f), gl_canon(g))));
}
#endif
// any * any -> canonical
GL_HD gl_t gl_mul(gl_t a, gl_t b) { return gl_canon(gl_mul_lazy(a, b)); }
GL_HD gl_t gl_sqr(gl_t a) { return gl_mul(a, a); }

// canonical x -> canonical x * 2^K, 0 < K < 96, by shifts and the 2^64 = 2^32 - 1, 2^96 = -1 folds.
// In Goldilocks 2 has order 192 and plonky2's roots of unity of order <= 64 are powers of two
// (w_64 = 2^39, w_16 = 2^156 = -2^60, w_4 = 2^48), so the twiddles of the last NTT stages need no
// general multiplication.
template <int K>
GL_HD gl_t gl_mul_pow2(gl_t x) {
    static_assert(K > 0 && K < 96, "shift out of range");
    if (K < 64) {
        const int k = K < 64 ? K : 1;
        return gl_canon(gl_reduce128_lazy(x << k, x >> (64 - k)));
    } else {
        // x 2^K = (A + B 2^64) 2^64 with A = low 64 bits of x << (K-64), B the bits above:
        // A 2^64 folds as usual and B 2^128 = -B 2^32
        const int j = K >= 64 ? K - 64 : 0;
        const gl_t A = x << j, B = j ? x >> ((64 - j) & 63) : 0;
        return gl_sub(gl_canon(gl_reduce128_lazy(0, A)), B << 32);
    }
}
GL_HD gl_t gl_pow(gl_t b, uint64_t e) {
    gl_t r = 1;
    while (e) { if (e & 1) r = gl_mul(r, b); b = gl_sqr(b); e >>= 1; }
    return r;
}
// Fermat inverse, a^(p-2); p-2 = 0xFFFFFFFEFFFFFFFF
GL_HD gl_t gl_inv(gl_t a) {
    // addition chain: a^(2^32-1) then assemble exponent 2^64 - 2^32 - 1
    gl_t t = a;                       // a^(2^1-1)
    gl_t x = a;
    // e31 = a^(2^31-1)
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 1; i < 31; i++) { x = gl_sqr(x); x = gl_mul(x, a); }
    gl_t e31 = x;
    // a^(2^32-2) = e31^2 ; a^(2^32 - 1) not needed.  exponent p-2 = (2^31-1)*2^33 + (2^32-1)
    // = e31 << 33 | (2^32 - 1):  p-2 = 0xFFFFFFFE_FFFFFFFF = (2^31-1)<<33 + 2^32-1
    gl_t y = e31;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < 33; i++) y = gl_sqr(y);
    gl_t e32 = gl_mul(gl_sqr(e31), a);  // a^(2^32-1)
    (void)t;
    return gl_mul(y, e32);
}
GL_HD gl_t gl_root_of_unity(int k) {
    gl_t g = 1753635133440165772ULL;  // POWER_OF_TWO_GENERATOR, order 2^32
    for (int i = 0; i < 32 - k; i++) g = gl_sqr(g);
    return g;
}

struct ext_t { gl_t c0, c1; };
GL_HD ext_t ext_make(gl_t a, gl_t b) { ext_t r; r.c0 = a; r.c1 = b; return r; }
GL_HD ext_t ext_add(ext_t a, ext_t b) { return ext_make(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
GL_HD ext_t ext_sub(ext_t a, ext_t b) { return ext_make(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
GL_HD ext_t ext_mul(ext_t a, ext_t b) {
    gl_t a1b1 = gl_mul(a.c1, b.c1);
    gl_t c0 = gl_add(gl_mul(a.c0, b.c0), gl_mul(7, a1b1));
    gl_t c1 = gl_add(gl_mul(a.c0, b.c1), gl_mul(a.c1, b.c0));
    return ext_make(c0, c1);
}
GL_HD ext_t ext_mul_base(ext_t a, gl_t b) { return ext_make(gl_mul(a.c0, b), gl_mul(a.c1, b)); }
GL_HD ext_t ext_add_base(ext_t a, gl_t b) { return ext_make(gl_add(a.c0, b), a.c1); }
GL_HD ext_t ext_inv(ext_t a) {
    gl_t d = gl_sub(gl_sqr(a.c0), gl_mul(7, gl_sqr(a.c1)));
    gl_t di = gl_inv(d);
    return ext_make(gl_mul(a.c0, di), gl_mul(gl_neg(a.c1), di));
}
GL_HD ext_t ext_pow(ext_t b, uint64_t e) {
    ext_t r = ext_make(1, 0);
    while (e) { if (e & 1) r = ext_mul(r, b); b = ext_mul(b, b); e >>= 1; }
    return r;
}
GL_HD uint32_t gl_bitrev(uint32_t x, int bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
#endif
}
