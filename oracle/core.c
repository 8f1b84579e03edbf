/* TEST INFRASTRUCTURE — CPU oracle, part 1: Poseidon, hashing, FFT, Merkle tree,
 * PolynomialBatch, Challenger.  See oracle.h for provenance (PARITY UNPINNED; restated from
 * upstream plonky2 paths, no line numbers available because the dependency is not vendored). */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Poseidon-Goldilocks, width 12 (upstream plonky2/src/hash/poseidon.rs, poseidon_goldilocks.rs)
 * naive round form: add constants, S-box x^7, MDS; 4 full + 22 partial + 4 full rounds.
 * ------------------------------------------------------------------------------------------ */
static const gl_t RC[360] = {
#include "poseidon_rc.inc"
};
static const uint64_t MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const uint64_t MDS_DIAG[12] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

static inline gl_t sbox7(gl_t x) {
    gl_t x2 = gl_sqr(x), x4 = gl_sqr(x2), x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}
static inline void mds_layer(gl_t s[12]) {
    /* out[r] = sum_i s[(i+r)%12]*CIRC[i] + s[r]*DIAG[r].  The coefficients are < 2^6, so the
     * 32-bit halves of the state can be accumulated separately in u64 without overflow and
     * recombined with one reduction (same value as the u128 accumulation upstream uses). */
    static const uint32_t C32[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[24], hi[24];
    for (int i = 0; i < 12; i++) { lo[i] = lo[i + 12] = (uint32_t)s[i]; hi[i] = hi[i + 12] = (uint32_t)(s[i] >> 32); }
    uint64_t al[12], ah[12];
    for (int r = 0; r < 12; r++) { al[r] = 0; ah[r] = 0; }
    for (int i = 0; i < 12; i++)
        for (int r = 0; r < 12; r++) { al[r] += (uint64_t)lo[i + r] * C32[i]; ah[r] += (uint64_t)hi[i + r] * C32[i]; }
    al[0] += (uint64_t)lo[0] * 8; ah[0] += (uint64_t)hi[0] * 8;
    for (int r = 0; r < 12; r++) s[r] = gl_reduce128((unsigned __int128)al[r] + ((unsigned __int128)ah[r] << 32));
}
/* sum_j a[j]*b[j] mod p for 11 terms: the 32-bit halves of b are accumulated separately so no
 * 128-bit carry tracking is needed (12 * 2^96 < 2^128) */
static inline gl_t dot11(const gl_t* a, const gl_t* b, unsigned __int128 init) {
    unsigned __int128 accl = init, acch = 0;
    for (int j = 0; j < 11; j++) {
        accl += (unsigned __int128)a[j] * (uint32_t)b[j];
        acch += (unsigned __int128)a[j] * (uint32_t)(b[j] >> 32);
    }
    return gl_add(gl_reduce128(accl), gl_mul(gl_reduce128(acch), (gl_t)1 << 32));
}
void orc_poseidon_naive(gl_t s[12]) {
    int r = 0;
    for (int phase = 0; phase < 3; phase++) {
        int nr = phase == 1 ? 22 : 4;
        for (int k = 0; k < nr; k++, r++) {
            for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], RC[12 * r + i]);
            if (phase == 1) s[0] = sbox7(s[0]);
            else for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
            mds_layer(s);
        }
    }
}

/* The 22 partial rounds in their sparse form (upstream poseidon.rs: partial_first_constant_layer,
 * mds_partial_layer_init, mds_partial_layer_fast); constants derived by tools/gen_poseidon_fast.py
 * and checked there against the naive rounds (the leading FIRST_C / K values also match the
 * recalled upstream FAST_PARTIAL_* tables). */
#include "poseidon_fast.inc"
static inline void full_round(gl_t s[12], int r) {
    for (int i = 0; i < 12; i++) s[i] = sbox7(gl_add(s[i], RC[12 * r + i]));
    mds_layer(s);
}
void orc_poseidon(gl_t s[12]) {
    for (int r = 0; r < 4; r++) full_round(s, r);
    for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], PFAST_FIRST_C[i]);
    {
        gl_t t[11];
        for (int r = 0; r < 11; r++) t[r] = dot11(PFAST_INIT + r * 11, s + 1, 0);
        for (int r = 0; r < 11; r++) s[r + 1] = t[r];
    }
    for (int r = 0; r < 22; r++) {
        gl_t t = gl_add(sbox7(s[0]), PFAST_K[r]);
        gl_t s0 = dot11(PFAST_VROW + r * 11, s + 1, (unsigned __int128)t * 25);
        for (int j = 0; j < 11; j++) s[j + 1] = gl_add(s[j + 1], gl_mul(PFAST_WCOL[r * 11 + j], t));
        s[0] = s0;
    }
    for (int r = 26; r < 30; r++) full_round(s, r);
}

/* hashing.rs: hash_n_to_m_no_pad — overwrite-mode sponge, rate 8 */
void orc_hash_n_to_m_no_pad(const gl_t* in, size_t n, gl_t* out, size_t m) {
    gl_t s[12] = {0};
    for (size_t off = 0; off < n; off += 8) {
        size_t len = n - off < 8 ? n - off : 8;
        memcpy(s, in + off, len * sizeof(gl_t));
        orc_poseidon(s);
    }
    size_t got = 0;
    for (;;) {
        for (int i = 0; i < 8; i++) { out[got++] = s[i]; if (got == m) return; }
        orc_poseidon(s);
    }
}
void orc_hash_no_pad(const gl_t* in, size_t n, gl_t out[4]) { orc_hash_n_to_m_no_pad(in, n, out, 4); }
/* hash_or_noop: inputs of <= 4 elements are the digest itself (zero padded) */
void orc_hash_or_noop(const gl_t* in, size_t n, gl_t out[4]) {
    if (n <= 4) { for (size_t i = 0; i < 4; i++) out[i] = i < n ? in[i] : 0; }
    else orc_hash_no_pad(in, n, out);
}
void orc_two_to_one(const gl_t l[4], const gl_t r[4], gl_t out[4]) {
    gl_t s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    orc_poseidon(s);
    memcpy(out, s, 4 * sizeof(gl_t));
}

/* ------------------------------------------------------------------------------------------
 * FFT (upstream field/src/fft.rs): natural order in/out; bit-reverse then DIT butterflies.
 * ------------------------------------------------------------------------------------------ */
size_t orc_reverse_bits(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
#define MAX_LOG 24
static gl_t* g_roots[MAX_LOG + 1][2]; /* [log][inverse]: n/2 powers of the root */
static const gl_t* root_table(int log_n, int inverse) {
    gl_t* t;
#pragma omp critical(orc_roots)
    {
        t = g_roots[log_n][inverse];
        if (!t) {
            size_t half = log_n ? ((size_t)1 << (log_n - 1)) : 1;
            t = (gl_t*)malloc(half * sizeof(gl_t));
            gl_t w = gl_root_of_unity(log_n);
            if (inverse) w = gl_inv(w);
            gl_t x = 1;
            for (size_t i = 0; i < half; i++) { t[i] = x; x = gl_mul(x, w); }
            g_roots[log_n][inverse] = t;
        }
    }
    return t;
}
static void fft_core(gl_t* a, int log_n, int inverse) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) {
        size_t j = orc_reverse_bits(i, log_n);
        if (i < j) { gl_t t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    const gl_t* roots = root_table(log_n, inverse);
    for (int s = 1; s <= log_n; s++) {
        size_t m = (size_t)1 << s, half = m >> 1, stride = n >> s;
        for (size_t k = 0; k < n; k += m)
            for (size_t j = 0; j < half; j++) {
                gl_t w = roots[j * stride];
                gl_t u = a[k + j], t = gl_mul(a[k + j + half], w);
                a[k + j] = gl_add(u, t);
                a[k + j + half] = gl_sub(u, t);
            }
    }
}
void orc_fft(gl_t* a, int log_n) { fft_core(a, log_n, 0); }
void orc_ifft(gl_t* a, int log_n) {
    fft_core(a, log_n, 1);
    gl_t ninv = gl_inv((gl_t)1 << log_n);
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) a[i] = gl_mul(a[i], ninv);
}
void orc_coset_fft(gl_t* a, int log_n, gl_t shift) {
    size_t n = (size_t)1 << log_n;
    gl_t s = 1;
    for (size_t i = 0; i < n; i++) { a[i] = gl_mul(a[i], s); s = gl_mul(s, shift); }
    orc_fft(a, log_n);
}
void orc_coset_ifft(gl_t* a, int log_n, gl_t shift) {
    size_t n = (size_t)1 << log_n;
    orc_ifft(a, log_n);
    gl_t si = gl_inv(shift), s = 1;
    for (size_t i = 0; i < n; i++) { a[i] = gl_mul(a[i], s); s = gl_mul(s, si); }
}

/* ------------------------------------------------------------------------------------------
 * MerkleTree::new(leaves, cap_height)   (upstream hash/merkle_tree.rs)
 * ------------------------------------------------------------------------------------------ */
static int log2_exact(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }

orc_merkle* orc_merkle_new(const gl_t* leaves, size_t num_leaves, size_t leaf_len, int cap_height) {
    orc_merkle* t = (orc_merkle*)calloc(1, sizeof(*t));
    int log_l = log2_exact(num_leaves);
    if (cap_height > log_l) cap_height = log_l; /* upstream asserts; callers never hit this */
    t->num_leaves = num_leaves; t->leaf_len = leaf_len; t->cap_height = cap_height; t->leaves = leaves;
    int L = log_l - cap_height;
    size_t total = 0;
    for (int k = 0; k < L; k++) total += (num_leaves >> k) * 4;
    t->digests = (gl_t*)malloc((total ? total : 4) * sizeof(gl_t));
    t->cap = (gl_t*)malloc(((size_t)4 << cap_height) * sizeof(gl_t));
    gl_t* prev = NULL; gl_t* cur = t->digests;
    for (int k = 0; k <= L; k++) {
        size_t cnt = num_leaves >> k;
        gl_t* dst = (k == L) ? t->cap : cur;
        if (k == 0) {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < cnt; i++) orc_hash_or_noop(leaves + i * leaf_len, leaf_len, dst + 4 * i);
        } else {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < cnt; i++) orc_two_to_one(prev + 8 * i, prev + 8 * i + 4, dst + 4 * i);
        }
        prev = dst;
        if (k < L) cur += cnt * 4;
    }
    return t;
}
void orc_merkle_free(orc_merkle* t) { if (t) { free(t->digests); free(t->cap); free(t); } }
size_t orc_merkle_path_len(const orc_merkle* t) { return (size_t)(log2_exact(t->num_leaves) - t->cap_height); }
const gl_t* orc_merkle_level(const orc_merkle* t, int level) {
    int L = (int)orc_merkle_path_len(t);
    if (level >= L) return t->cap;
    const gl_t* p = t->digests;
    for (int k = 0; k < level; k++) p += (t->num_leaves >> k) * 4;
    return p;
}
void orc_merkle_prove(const orc_merkle* t, size_t leaf_index, gl_t* siblings) {
    int L = (int)orc_merkle_path_len(t);
    size_t idx = leaf_index;
    for (int k = 0; k < L; k++) {
        const gl_t* lvl = orc_merkle_level(t, k);
        memcpy(siblings + 4 * k, lvl + 4 * (idx ^ 1), 4 * sizeof(gl_t));
        idx >>= 1;
    }
}
/* merkle_proofs.rs: verify_merkle_proof_to_cap */
int orc_merkle_verify(const gl_t* leaf, size_t leaf_len, size_t leaf_index, const gl_t* cap, int cap_height,
                      const gl_t* siblings, size_t path_len) {
    gl_t cur[4], nxt[4];
    orc_hash_or_noop(leaf, leaf_len, cur);
    size_t idx = leaf_index;
    for (size_t k = 0; k < path_len; k++) {
        if (idx & 1) orc_two_to_one(siblings + 4 * k, cur, nxt); else orc_two_to_one(cur, siblings + 4 * k, nxt);
        memcpy(cur, nxt, sizeof(cur));
        idx >>= 1;
    }
    if (idx >= ((size_t)1 << cap_height)) return 0;
    return memcmp(cur, cap + 4 * idx, sizeof(cur)) == 0;
}

/* ------------------------------------------------------------------------------------------
 * PolynomialBatch::from_values / from_coeffs (upstream fri/oracle.rs), blinding = false
 * ------------------------------------------------------------------------------------------ */
orc_batch* orc_batch_from_coeffs(const gl_t* cols, int ncols, int log_n, int rate_bits, int cap_height) {
    orc_batch* b = (orc_batch*)calloc(1, sizeof(*b));
    size_t n = (size_t)1 << log_n, N = n << rate_bits;
    int log_N = log_n + rate_bits;
    b->ncols = ncols; b->log_n = log_n; b->rate_bits = rate_bits; b->cap_height = cap_height;
    b->coeffs = (gl_t*)malloc((size_t)ncols * n * sizeof(gl_t));
    memcpy(b->coeffs, cols, (size_t)ncols * n * sizeof(gl_t));
    b->leaves = (gl_t*)malloc((size_t)ncols * N * sizeof(gl_t));
    root_table(log_N, 0);
#pragma omp parallel
    {
        gl_t* tmp = (gl_t*)malloc(N * sizeof(gl_t));
#pragma omp for schedule(dynamic, 1)
        for (int c = 0; c < ncols; c++) {
            /* lde(rate_bits): zero pad; coset_fft(shift = 7) */
            memcpy(tmp, b->coeffs + (size_t)c * n, n * sizeof(gl_t));
            memset(tmp + n, 0, (N - n) * sizeof(gl_t));
            orc_coset_fft(tmp, log_N, GL_GENERATOR);
            /* transpose + reverse_index_bits_in_place */
            for (size_t i = 0; i < N; i++) b->leaves[orc_reverse_bits(i, log_N) * ncols + c] = tmp[i];
        }
        free(tmp);
    }
    b->tree = orc_merkle_new(b->leaves, N, (size_t)ncols, cap_height);
    return b;
}
orc_batch* orc_batch_from_values(const gl_t* cols, int ncols, int log_n, int rate_bits, int cap_height) {
    size_t n = (size_t)1 << log_n;
    gl_t* co = (gl_t*)malloc((size_t)ncols * n * sizeof(gl_t));
    memcpy(co, cols, (size_t)ncols * n * sizeof(gl_t));
    root_table(log_n, 1);
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < ncols; c++) orc_ifft(co + (size_t)c * n, log_n);
    orc_batch* b = orc_batch_from_coeffs(co, ncols, log_n, rate_bits, cap_height);
    free(co);
    return b;
}
void orc_batch_free(orc_batch* b) {
    if (!b) return;
    orc_merkle_free(b->tree); free(b->coeffs); free(b->leaves); free(b);
}

/* ------------------------------------------------------------------------------------------
 * Challenger (upstream iop/challenger.rs): duplex sponge in overwrite mode
 * ------------------------------------------------------------------------------------------ */
void orc_challenger_init(orc_challenger* c) { memset(c, 0, sizeof(*c)); }
static void duplexing(orc_challenger* c) {
    for (int i = 0; i < c->in_len; i++) c->state[i] = c->in_buf[i];
    c->in_len = 0;
    orc_poseidon(c->state);
    memcpy(c->out_buf, c->state, 8 * sizeof(gl_t));
    c->out_len = 8;
}
void orc_challenger_observe(orc_challenger* c, gl_t x) {
    c->out_len = 0;
    c->in_buf[c->in_len++] = x;
    if (c->in_len == 8) duplexing(c);
}
void orc_challenger_observe_many(orc_challenger* c, const gl_t* x, size_t n) {
    for (size_t i = 0; i < n; i++) orc_challenger_observe(c, x[i]);
}
gl_t orc_challenger_get(orc_challenger* c) {
    if (c->in_len != 0 || c->out_len == 0) duplexing(c);
    return c->out_buf[--c->out_len];
}
ext_t orc_challenger_get_ext(orc_challenger* c) {
    ext_t r; r.c0 = orc_challenger_get(c); r.c1 = orc_challenger_get(c); return r;
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
