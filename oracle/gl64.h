/* TEST INFRASTRUCTURE — CPU oracle for the plonky2 prove() hot path.  Not part of the product;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  PARITY UNPINNED: the algorithm lives in the un-vendored dependency
 * plonky2 = { git 0xPARC/plonky2, rev 109d517d09c210ae4c2cee381d3e3fbc04aa3812 }
 * (/root/reference/Cargo.toml:12); this file restates its published algorithm from upstream
 * 0xPolygonZero/plonky2 (field/src/goldilocks_field.rs, field/src/extension/quadratic.rs).
 *
 * Goldilocks field F = GF(p), p = 2^64 - 2^32 + 1, and its quadratic extension F[X]/(X^2 - 7).
 * All values are kept canonical (< p) at every function boundary.
 */
#ifndef ORACLE_GL64_H
#define ORACLE_GL64_H
#include <stdint.h>
#include <stddef.h>

typedef uint64_t gl_t;
#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */
#define GL_GENERATOR 7ULL           /* MULTIPLICATIVE_GROUP_GENERATOR == coset_shift() */
#define GL_POW2_GENERATOR 1753635133440165772ULL /* POWER_OF_TWO_GENERATOR, order 2^32 */
#define GL_EXT_W 7ULL

/* all reductions are written branch-free (data-dependent branches mispredict heavily) */
static inline gl_t gl_add(gl_t a, gl_t b) {
    uint64_t s = a + b;
    uint64_t over = (uint64_t)(s < a) | (uint64_t)(s >= GL_P);
    return s - (GL_P & (0 - over));
}
static inline gl_t gl_sub(gl_t a, gl_t b) {
    uint64_t d = a - b;
    return d + (GL_P & (0 - (uint64_t)(a < b)));
}
static inline gl_t gl_neg(gl_t a) { return a ? GL_P - a : 0; }
static inline gl_t gl_reduce128(unsigned __int128 x) {
    /* upstream reduce128: x_lo - x_hi_hi + x_hi_lo * EPSILON, folded with 2^64 = 2^32 - 1, 2^96 = -1 */
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    t0 -= GL_EPS & (0 - (uint64_t)(lo < hi_hi));
    uint64_t t1 = hi_lo * GL_EPS;
    uint64_t t2 = t0 + t1;
    t2 += GL_EPS & (0 - (uint64_t)(t2 < t1));
    t2 -= GL_P & (0 - (uint64_t)(t2 >= GL_P));
    return t2;
}
static inline gl_t gl_mul(gl_t a, gl_t b) { return gl_reduce128((unsigned __int128)a * b); }
static inline gl_t gl_sqr(gl_t a) { return gl_mul(a, a); }
static inline gl_t gl_pow(gl_t b, uint64_t e) {
    gl_t r = 1;
    while (e) { if (e & 1) r = gl_mul(r, b); b = gl_sqr(b); e >>= 1; }
    return r;
}
static inline gl_t gl_inv(gl_t a) { return gl_pow(a, GL_P - 2); }
static inline gl_t gl_from_u64(uint64_t x) { return x >= GL_P ? x - GL_P : x; }
/* primitive_root_of_unity(k): POWER_OF_TWO_GENERATOR^(2^(32-k)) */
static inline gl_t gl_root_of_unity(int k) {
    gl_t g = GL_POW2_GENERATOR;
    for (int i = 0; i < 32 - k; i++) g = gl_sqr(g);
    return g;
}

typedef struct { gl_t c0, c1; } ext_t;
static inline ext_t ext_make(gl_t a, gl_t b) { ext_t r = {a, b}; return r; }
static inline ext_t ext_from(gl_t a) { ext_t r = {a, 0}; return r; }
static inline ext_t ext_add(ext_t a, ext_t b) { return ext_make(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
static inline ext_t ext_sub(ext_t a, ext_t b) { return ext_make(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
static inline ext_t ext_neg(ext_t a) { return ext_make(gl_neg(a.c0), gl_neg(a.c1)); }
static inline ext_t ext_mul(ext_t a, ext_t b) {
    gl_t c0 = gl_add(gl_mul(a.c0, b.c0), gl_mul(GL_EXT_W, gl_mul(a.c1, b.c1)));
    gl_t c1 = gl_add(gl_mul(a.c0, b.c1), gl_mul(a.c1, b.c0));
    return ext_make(c0, c1);
}
static inline ext_t ext_mul_base(ext_t a, gl_t b) { return ext_make(gl_mul(a.c0, b), gl_mul(a.c1, b)); }
static inline ext_t ext_inv(ext_t a) {
    /* 1/(a0 + a1 X) = (a0 - a1 X) / (a0^2 - 7 a1^2) */
    gl_t d = gl_sub(gl_sqr(a.c0), gl_mul(GL_EXT_W, gl_sqr(a.c1)));
    gl_t di = gl_inv(d);
    return ext_make(gl_mul(a.c0, di), gl_mul(gl_neg(a.c1), di));
}
static inline int ext_eq(ext_t a, ext_t b) { return a.c0 == b.c0 && a.c1 == b.c1; }
static inline ext_t ext_pow(ext_t b, uint64_t e) {
    ext_t r = ext_from(1);
    while (e) { if (e & 1) r = ext_mul(r, b); b = ext_mul(b, b); e >>= 1; }
    return r;
}
static inline ext_t ext_exp_pow2(ext_t b, int k) { for (int i = 0; i < k; i++) b = ext_mul(b, b); return b; }
#endif
