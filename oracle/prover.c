/* TEST INFRASTRUCTURE — CPU oracle, part 2: the whole prove() path and the verifier.
 * PARITY UNPINNED (see oracle.h): restated from upstream plonky2 by path —
 *   plonk/prover.rs        prove_with_partition_witness, all_wires_permutation_partial_products,
 *                          compute_lookup_polys, compute_quotient_polys
 *   plonk/proof.rs         OpeningSet::new, to_fri_openings
 *   fri/oracle.rs          PolynomialBatch::prove_openings
 *   fri/prover.rs          fri_committed_trees, fri_proof_of_work, fri_prover_query_rounds
 *   plonk/verifier.rs, fri/verifier.rs, plonk/get_challenges.rs   orc_verify
 * The reference enters all of it through `data.prove(pw)` / `data.verify(proof)`
 * (/root/reference/aes-gcm/src/circuit_gcm.rs:781-782).
 * Deviation kept on purpose: the PoW witness is the LOWEST valid nonce (north_star), where
 * upstream takes any (`find_any`).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
#define STAGE(name) do { if (timing) { double t_ = now_s(); fprintf(stderr, "[oracle] %-18s %.3f s\n", name, t_ - t_last); t_last = t_; } } while (0)

/* constants the PoseidonGate constraints need (same tables as core.c) */
static const gl_t VAN_RC[360] = {
#include "poseidon_rc.inc"
};
static const gl_t VAN_MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
#include "poseidon_fast.inc"
/* ---- two instantiations of the vanishing-polynomial evaluator ---- */
#define VNAME(x) vb_##x
#define FT gl_t
#define F_ZERO 0
#define F_ONE 1
#define F_ADD gl_add
#define F_SUB gl_sub
#define F_MUL gl_mul
#define F_MULB gl_mul
#define F_FROMB(x) (x)
#define F_ADDB gl_add
#define F_SUBB gl_sub
#define F_BSUB(b, x) gl_sub((b), (x))
#include "vanishing_impl.h"
#undef VNAME
#undef FT
#undef F_ZERO
#undef F_ONE
#undef F_ADD
#undef F_SUB
#undef F_MUL
#undef F_MULB
#undef F_FROMB
#undef F_ADDB
#undef F_SUBB
#undef F_BSUB

static inline ext_t ext_addb(ext_t a, gl_t b) { a.c0 = gl_add(a.c0, b); return a; }
static inline ext_t ext_subb(ext_t a, gl_t b) { a.c0 = gl_sub(a.c0, b); return a; }
static inline ext_t ext_bsub(gl_t b, ext_t a) { return ext_make(gl_sub(b, a.c0), gl_neg(a.c1)); }
#define VNAME(x) ve_##x
#define FT ext_t
#define F_ZERO ext_from(0)
#define F_ONE ext_from(1)
#define F_ADD ext_add
#define F_SUB ext_sub
#define F_MUL ext_mul
#define F_MULB ext_mul_base
#define F_FROMB ext_from
#define F_ADDB ext_addb
#define F_SUBB ext_subb
#define F_BSUB ext_bsub
#include "vanishing_impl.h"

/* ------------------------------------------------------------------------------------------ */
struct orc_prover_data {
    orc_circuit c;
    orc_gate* gates; int32_t* lut_lens; uint16_t* lut_data; int32_t* lookup_rows; gl_t* k_is; gl_t* cs_values;
    orc_batch* cs;      /* constants_sigmas commitment */
    gl_t* subgroup;     /* g^i, i < n */
};

static int cfg_num_lookup_polys(const orc_circuit* c) {
    if (c->num_luts == 0) return 0;
    int lu_slots = c->num_routed_wires / 2, lu_degree = c->quotient_degree_factor - 1;
    return (lu_slots + lu_degree - 1) / lu_degree + 1;
}
static int cfg_nc(const orc_circuit* c) { return c->num_selectors + c->num_lookup_selectors + c->num_constants; }
static int cfg_zs_cols(const orc_circuit* c) {
    return c->num_challenges * (1 + c->num_partial_products) + c->num_challenges * cfg_num_lookup_polys(c);
}
static int cfg_lut_total(const orc_circuit* c) { int t = 0; for (int i = 0; i < c->num_luts; i++) t += c->lut_lens[i]; return t; }

orc_prover_data* orc_circuit_load(const orc_circuit* c) {
    orc_prover_data* pd = (orc_prover_data*)calloc(1, sizeof(*pd));
    pd->c = *c;
    size_t n = (size_t)1 << c->degree_bits;
    int ncs = cfg_nc(c) + c->num_routed_wires;
#define DUP(dst, src, cnt, T) do { pd->dst = (T*)malloc(((cnt) ? (cnt) : 1) * sizeof(T)); memcpy(pd->dst, src, (cnt) * sizeof(T)); } while (0)
    DUP(gates, c->gates, (size_t)c->num_gates, orc_gate);
    DUP(lut_lens, c->lut_lens, (size_t)c->num_luts, int32_t);
    DUP(lut_data, c->lut_data, (size_t)2 * cfg_lut_total(c), uint16_t);
    DUP(lookup_rows, c->lookup_rows, (size_t)3 * c->num_luts, int32_t);
    DUP(k_is, c->k_is, (size_t)c->num_routed_wires, gl_t);
    DUP(cs_values, c->constants_sigmas, (size_t)ncs * n, gl_t);
    pd->c.gates = pd->gates; pd->c.lut_lens = pd->lut_lens; pd->c.lut_data = pd->lut_data;
    pd->c.lookup_rows = pd->lookup_rows; pd->c.k_is = pd->k_is; pd->c.constants_sigmas = pd->cs_values;
    pd->cs = orc_batch_from_values(pd->cs_values, ncs, c->degree_bits, c->rate_bits, c->cap_height);
    pd->subgroup = (gl_t*)malloc(n * sizeof(gl_t));
    gl_t g = gl_root_of_unity(c->degree_bits), x = 1;
    for (size_t i = 0; i < n; i++) { pd->subgroup[i] = x; x = gl_mul(x, g); }
    return pd;
}
void orc_circuit_free(orc_prover_data* pd) {
    if (!pd) return;
    orc_batch_free(pd->cs);
    free(pd->gates); free(pd->lut_lens); free(pd->lut_data); free(pd->lookup_rows); free(pd->k_is); free(pd->cs_values);
    free(pd->subgroup); free(pd);
}
const gl_t* orc_circuit_cap(const orc_prover_data* pd) { return pd->cs->tree->cap; }

/* ---- proof layout ---- */
static size_t fri_final_len(const orc_circuit* c) {
    size_t n = (size_t)1 << c->degree_bits;
    for (int i = 0; i < c->num_reduction_arity_bits; i++) n >>= c->reduction_arity_bits[i];
    return n;
}
size_t orc_proof_len(const orc_circuit* c) {
    size_t cap = (size_t)4 << c->cap_height;
    int nlp = cfg_num_lookup_polys(c), nch = c->num_challenges;
    int logN = c->degree_bits + c->rate_bits;
    size_t open = 2 * ((size_t)cfg_nc(c) + c->num_routed_wires + c->num_wires + 2 * nch + (size_t)nch * c->num_partial_products +
                       (size_t)nch * c->quotient_degree_factor + 2 * (size_t)nch * nlp);
    size_t len = 3 * cap + open + (size_t)c->num_reduction_arity_bits * cap;
    size_t per_q = 0;
    int cols[4] = {cfg_nc(c) + c->num_routed_wires, c->num_wires, cfg_zs_cols(c), nch * c->quotient_degree_factor};
    for (int o = 0; o < 4; o++) per_q += cols[o] + 1 + 4 * (size_t)(logN - c->cap_height);
    int lg = logN;
    for (int l = 0; l < c->num_reduction_arity_bits; l++) {
        int ab = c->reduction_arity_bits[l];
        lg -= ab;
        per_q += (2u << ab) + 1 + 4 * (size_t)(lg - c->cap_height);
    }
    len += per_q * c->num_query_rounds + 2 * fri_final_len(c) + 1 + c->num_public_inputs;
    return len;
}

static ext_t eval_poly_base_at_ext(const gl_t* coeffs, size_t n, ext_t z) {
    ext_t acc = ext_from(0);
    for (size_t i = n; i-- > 0;) acc = ext_addb(ext_mul(acc, z), coeffs[i]);
    return acc;
}
static ext_t eval_poly_ext(const ext_t* coeffs, size_t n, ext_t z) {
    ext_t acc = ext_from(0);
    for (size_t i = n; i-- > 0;) acc = ext_add(ext_mul(acc, z), coeffs[i]);
    return acc;
}
static void batch_inverse(gl_t* x, size_t n, gl_t* tmp) {
    if (!n) return;
    tmp[0] = x[0];
    for (size_t i = 1; i < n; i++) tmp[i] = gl_mul(tmp[i - 1], x[i]);
    gl_t inv = gl_inv(tmp[n - 1]);
    for (size_t i = n - 1; i > 0; i--) { gl_t xi = x[i]; x[i] = gl_mul(inv, tmp[i - 1]); inv = gl_mul(inv, xi); }
    x[0] = inv;
}

/* get_lut_poly(...).eval(delta): coefficients are the (padded) combos reversed */
static void compute_lut_evals(const orc_circuit* c, const gl_t* deltas, gl_t* out /*[nch][num_luts]*/) {
    int lut_slots = c->num_routed_wires / 3;
    for (int i = 0; i < c->num_challenges; i++) {
        gl_t b = deltas[4 * i + 1], delta = deltas[4 * i + 3];
        const uint16_t* data = c->lut_data;
        for (int r = 0; r < c->num_luts; r++) {
            int len = c->lut_lens[r];
            int rows = (len + lut_slots - 1) / lut_slots;
            int degree = rows * lut_slots;
            /* sum_e combo_e * delta^(degree-1-e): Horner over e ascending */
            gl_t acc = 0;
            for (int e = 0; e < degree; e++) {
                gl_t combo = e < len ? gl_add(data[2 * e], gl_mul(b, data[2 * e + 1])) : 0;
                acc = gl_add(gl_mul(acc, delta), combo);
            }
            out[i * c->num_luts + r] = acc;
            data += 2 * len;
        }
    }
}

/* ---- writer ---- */
typedef struct { gl_t* p; size_t pos, cap; } wr_t;
static void wr(wr_t* w, const gl_t* src, size_t n) { if (w->pos + n <= w->cap) memcpy(w->p + w->pos, src, n * sizeof(gl_t)); w->pos += n; }
static void wr1(wr_t* w, gl_t v) { wr(w, &v, 1); }
static void wr_ext(wr_t* w, ext_t e) { wr1(w, e.c0); wr1(w, e.c1); }

long orc_prove(const orc_prover_data* pd, const gl_t* wires, const gl_t* public_inputs, gl_t* proof_out, size_t proof_cap) {
    return orc_prove_debug(pd, wires, public_inputs, proof_out, proof_cap, NULL, NULL, NULL);
}

long orc_prove_debug(const orc_prover_data* pd, const gl_t* wires, const gl_t* public_inputs, gl_t* proof_out,
                     size_t proof_cap, orc_transcript* tr, gl_t* zs_out, gl_t* qchunks_out) {
    const orc_circuit* c = &pd->c;
    const int nch = c->num_challenges, R = c->num_routed_wires, W = c->num_wires, qdf = c->quotient_degree_factor;
    const int num_prods = c->num_partial_products, nlp = cfg_num_lookup_polys(c), NC = cfg_nc(c);
    const int has_lookup = c->num_luts > 0;
    const int logn = c->degree_bits, logN = logn + c->rate_bits;
    const size_t n = (size_t)1 << logn, N = (size_t)1 << logN;
    const int zs_cols = cfg_zs_cols(c);
    if (qdf != (1 << c->rate_bits) || nch > 4) return -2;

    const int timing = getenv("ORC_TIMING") != NULL; double t_last = now_s();
    gl_t pi_hash[4];
    orc_hash_no_pad(public_inputs, (size_t)c->num_public_inputs, pi_hash);

    /* ---- stage A: wires commitment ---- */
    orc_batch* wb = orc_batch_from_values(wires, W, logn, c->rate_bits, c->cap_height);
    STAGE("wires commit");
    orc_challenger ch; orc_challenger_init(&ch);
    orc_challenger_observe_many(&ch, c->circuit_digest, 4);
    orc_challenger_observe_many(&ch, pi_hash, 4);
    orc_challenger_observe_many(&ch, wb->tree->cap, (size_t)4 << c->cap_height);
    gl_t betas[4], gammas[4], deltas[16], alphas[4];
    for (int i = 0; i < nch; i++) betas[i] = orc_challenger_get(&ch);
    for (int i = 0; i < nch; i++) gammas[i] = orc_challenger_get(&ch);
    if (has_lookup) {
        /* deltas = betas || gammas || 2*nch more, then cut into chunks of NUM_COINS_LOOKUP = 4 */
        int k = 0;
        for (int i = 0; i < nch; i++) deltas[k++] = betas[i];
        for (int i = 0; i < nch; i++) deltas[k++] = gammas[i];
        for (int i = 0; i < 2 * nch; i++) deltas[k++] = orc_challenger_get(&ch);
    } else memset(deltas, 0, sizeof(deltas));

    /* ---- Z and partial products (wires_permutation_partial_products_and_zs) ---- */
    gl_t* zs = (gl_t*)calloc((size_t)zs_cols * n, sizeof(gl_t));
    const gl_t* sig_vals = c->constants_sigmas + (size_t)NC * n;   /* [R][n] */
    for (int i = 0; i < nch; i++) {
        gl_t* chunk = (gl_t*)malloc((size_t)(num_prods + 1) * n * sizeof(gl_t));   /* [row][num_prods+1] */
#pragma omp parallel
        {
            gl_t num[256], den[256], tmp[256];
#pragma omp for schedule(static)
            for (size_t r = 0; r < n; r++) {
                gl_t x = pd->subgroup[r];
                for (int j = 0; j < R; j++) {
                    gl_t w = wires[(size_t)j * n + r];
                    num[j] = gl_add(gl_add(w, gl_mul(betas[i], gl_mul(c->k_is[j], x))), gammas[i]);
                    den[j] = gl_add(gl_add(w, gl_mul(betas[i], sig_vals[(size_t)j * n + r])), gammas[i]);
                }
                batch_inverse(den, (size_t)R, tmp);
                for (int ck = 0; ck <= num_prods; ck++) {
                    gl_t p = 1;
                    int lo = ck * qdf, hi = lo + qdf < R ? lo + qdf : R;
                    for (int j = lo; j < hi; j++) p = gl_mul(p, gl_mul(num[j], den[j]));
                    chunk[r * (num_prods + 1) + ck] = p;
                }
            }
        }
        gl_t z = 1;
        gl_t* Z = zs + (size_t)i * n;
        gl_t* PP = zs + ((size_t)nch + (size_t)i * num_prods) * n;
        for (size_t r = 0; r < n; r++) {
            Z[r] = z;
            gl_t acc = z;
            for (int ck = 0; ck <= num_prods; ck++) {
                acc = gl_mul(acc, chunk[r * (num_prods + 1) + ck]);
                if (ck < num_prods) PP[(size_t)ck * n + r] = acc;
            }
            z = acc;   /* Z(g x) */
        }
        free(chunk);
    }
    /* ---- lookup polys (compute_lookup_polys) ---- */
    if (has_lookup) {
        const int lu_slots = R / 2, lut_slots = R / 3, lu_degree = qdf - 1;
        const int num_partial = nlp - 1;
        const int lut_degree = (lut_slots + num_partial - 1) / num_partial;
        for (int i = 0; i < nch; i++) {
            gl_t da = deltas[4 * i], db = deltas[4 * i + 1], dalpha = deltas[4 * i + 2], ddelta = deltas[4 * i + 3];
            gl_t* base = zs + ((size_t)nch * (1 + num_prods) + (size_t)i * nlp) * n;   /* [nlp][n] */
            for (int l = 0; l < c->num_luts; l++) {
                int last_lu = c->lookup_rows[3 * l], last_lut = c->lookup_rows[3 * l + 1], first_lut = c->lookup_rows[3 * l + 2];
                gl_t inv[64], tmp[64];
                for (int row = first_lut; row >= last_lut; row--) {
                    gl_t re = base[(size_t)row + 1];
                    for (int s = 0; s < lut_slots; s++) {
                        gl_t in = wires[(size_t)(3 * s) * n + row], o = wires[(size_t)(3 * s + 1) * n + row];
                        inv[s] = gl_sub(dalpha, gl_add(in, gl_mul(da, o)));
                        re = gl_add(gl_mul(re, ddelta), gl_add(in, gl_mul(db, o)));
                    }
                    batch_inverse(inv, (size_t)lut_slots, tmp);
                    base[row] = re;
                    for (int slot = 0; slot < num_partial; slot++) {
                        gl_t prev = slot ? base[(size_t)slot * n + row] : base[(size_t)num_partial * n + row + 1];
                        int s0 = slot * lut_degree, s1 = s0 + lut_degree < lut_slots ? s0 + lut_degree : lut_slots;
                        for (int s = s0; s < s1; s++) prev = gl_add(prev, gl_mul(wires[(size_t)(3 * s + 2) * n + row], inv[s]));
                        base[(size_t)(slot + 1) * n + row] = prev;
                    }
                }
                for (int row = last_lut - 1; row >= last_lu; row--) {
                    for (int s = 0; s < lu_slots; s++) {
                        gl_t in = wires[(size_t)(2 * s) * n + row], o = wires[(size_t)(2 * s + 1) * n + row];
                        inv[s] = gl_sub(dalpha, gl_add(in, gl_mul(da, o)));
                    }
                    batch_inverse(inv, (size_t)lu_slots, tmp);
                    for (int slot = 0; slot < num_partial; slot++) {
                        gl_t prev = slot ? base[(size_t)slot * n + row] : base[(size_t)num_partial * n + row + 1];
                        int s0 = slot * lu_degree, s1 = s0 + lu_degree < lu_slots ? s0 + lu_degree : lu_slots;
                        gl_t sum = 0;
                        for (int s = s0; s < s1; s++) sum = gl_add(sum, inv[s]);
                        base[(size_t)(slot + 1) * n + row] = gl_sub(prev, sum);
                    }
                }
            }
        }
    }
    STAGE("zs/pp/lookup build");
    if (zs_out) memcpy(zs_out, zs, (size_t)zs_cols * n * sizeof(gl_t));

    /* ---- stage B: commit Z / partial products / lookup polys ---- */
    orc_batch* zb = orc_batch_from_values(zs, zs_cols, logn, c->rate_bits, c->cap_height);
    free(zs);
    STAGE("zs commit");
    orc_challenger_observe_many(&ch, zb->tree->cap, (size_t)4 << c->cap_height);
    for (int i = 0; i < nch; i++) alphas[i] = orc_challenger_get(&ch);

    /* ---- stage C: compute_quotient_polys ---- */
    gl_t lut_evals[64];
    if (has_lookup) compute_lut_evals(c, deltas, lut_evals);
    vb_ctx vc = {c, betas, gammas, alphas, deltas, lut_evals, pi_hash};
    gl_t* qvals = (gl_t*)malloc((size_t)nch * N * sizeof(gl_t));   /* [nch][N] natural order */
    {
        /* ZeroPolyOnCoset: Z_H(x) = 7^n * w_rate^(i mod rate) - 1 */
        const int rate = 1 << c->rate_bits;
        gl_t zh[64], zh_inv[64];
        gl_t sn = gl_pow(GL_GENERATOR, n), wr_ = gl_root_of_unity(c->rate_bits), t = 1;
        for (int k = 0; k < rate; k++) { zh[k] = gl_sub(gl_mul(sn, t), 1); zh_inv[k] = gl_inv(zh[k]); t = gl_mul(t, wr_); }
        const gl_t wN = gl_root_of_unity(logN);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < N; i++) {
            gl_t x = gl_mul(GL_GENERATOR, gl_pow(wN, i));
            size_t j = orc_reverse_bits(i, logN), jn = orc_reverse_bits((i + (size_t)qdf) % N, logN);
            const gl_t* cs_row = pd->cs->leaves + j * (size_t)(NC + R);
            const gl_t* w_row = wb->leaves + j * (size_t)W;
            const gl_t* z_row = zb->leaves + j * (size_t)zs_cols;
            const gl_t* zn_row = zb->leaves + jn * (size_t)zs_cols;
            /* eval_l_0(i, x) = Z_H(x) / (n (x - 1)) */
            gl_t l0 = gl_mul(zh[i % rate], gl_inv(gl_mul((gl_t)n, gl_sub(x, 1))));
            gl_t out[4];
            vb_eval(&vc, x, l0, cs_row, w_row, z_row, zn_row, z_row + nch, z_row + nch * (1 + num_prods),
                    zn_row + nch * (1 + num_prods), cs_row + NC, out);
            for (int k = 0; k < nch; k++) qvals[(size_t)k * N + i] = gl_mul(out[k], zh_inv[i % rate]);
        }
    }
    STAGE("quotient eval");
    gl_t* qchunks = (gl_t*)malloc((size_t)nch * N * sizeof(gl_t));   /* nch*qdf polys of n coeffs == [nch][N] */
#pragma omp parallel for
    for (int k = 0; k < nch; k++) {
        orc_coset_ifft(qvals + (size_t)k * N, logN, GL_GENERATOR);
        memcpy(qchunks + (size_t)k * N, qvals + (size_t)k * N, N * sizeof(gl_t));
    }
    free(qvals);
    if (qchunks_out) memcpy(qchunks_out, qchunks, (size_t)nch * N * sizeof(gl_t));
    /* ---- stage D ---- */
    orc_batch* qb = orc_batch_from_coeffs(qchunks, nch * qdf, logn, c->rate_bits, c->cap_height);
    free(qchunks);
    STAGE("quotient commit");
    orc_challenger_observe_many(&ch, qb->tree->cap, (size_t)4 << c->cap_height);
    ext_t zeta = orc_challenger_get_ext(&ch);
    gl_t g = gl_root_of_unity(logn);
    ext_t zeta_next = ext_mul_base(zeta, g);

    /* ---- openings ---- */
    const orc_batch* oracles[4] = {pd->cs, wb, zb, qb};
    const int ocols[4] = {NC + R, W, zs_cols, nch * qdf};
    ext_t* open_zeta[4];
    for (int o = 0; o < 4; o++) {
        open_zeta[o] = (ext_t*)malloc((size_t)ocols[o] * sizeof(ext_t));
#pragma omp parallel for schedule(dynamic, 4)
        for (int k = 0; k < ocols[o]; k++) open_zeta[o][k] = eval_poly_base_at_ext(oracles[o]->coeffs + (size_t)k * n, n, zeta);
    }
    ext_t* open_next = (ext_t*)malloc((size_t)zs_cols * sizeof(ext_t));
#pragma omp parallel for schedule(dynamic, 4)
    for (int k = 0; k < zs_cols; k++) open_next[k] = eval_poly_base_at_ext(zb->coeffs + (size_t)k * n, n, zeta_next);

    STAGE("openings");
    wr_t w = {proof_out, 0, proof_cap};
    const size_t capw = (size_t)4 << c->cap_height;
    wr(&w, wb->tree->cap, capw); wr(&w, zb->tree->cap, capw); wr(&w, qb->tree->cap, capw);
    size_t open_start = w.pos;
    const int zpp = nch * (1 + num_prods);
    for (int k = 0; k < NC; k++) wr_ext(&w, open_zeta[0][k]);                 /* constants */
    for (int k = 0; k < R; k++) wr_ext(&w, open_zeta[0][NC + k]);             /* plonk_sigmas */
    for (int k = 0; k < W; k++) wr_ext(&w, open_zeta[1][k]);                  /* wires */
    for (int k = 0; k < nch; k++) wr_ext(&w, open_zeta[2][k]);                /* plonk_zs */
    for (int k = 0; k < nch; k++) wr_ext(&w, open_next[k]);                   /* plonk_zs_next */
    for (int k = nch; k < zpp; k++) wr_ext(&w, open_zeta[2][k]);              /* partial_products */
    for (int k = 0; k < nch * qdf; k++) wr_ext(&w, open_zeta[3][k]);          /* quotient_polys */
    for (int k = zpp; k < zs_cols; k++) wr_ext(&w, open_zeta[2][k]);          /* lookup_zs */
    for (int k = zpp; k < zs_cols; k++) wr_ext(&w, open_next[k]);             /* lookup_zs_next */
    (void)open_start;
    /* observe_openings(to_fri_openings): batch zeta then batch zeta_next */
    for (int k = 0; k < NC + R; k++) { orc_challenger_observe(&ch, open_zeta[0][k].c0); orc_challenger_observe(&ch, open_zeta[0][k].c1); }
    for (int k = 0; k < W; k++) { orc_challenger_observe(&ch, open_zeta[1][k].c0); orc_challenger_observe(&ch, open_zeta[1][k].c1); }
    for (int k = 0; k < zpp; k++) { orc_challenger_observe(&ch, open_zeta[2][k].c0); orc_challenger_observe(&ch, open_zeta[2][k].c1); }
    for (int k = 0; k < nch * qdf; k++) { orc_challenger_observe(&ch, open_zeta[3][k].c0); orc_challenger_observe(&ch, open_zeta[3][k].c1); }
    for (int k = zpp; k < zs_cols; k++) { orc_challenger_observe(&ch, open_zeta[2][k].c0); orc_challenger_observe(&ch, open_zeta[2][k].c1); }
    for (int k = 0; k < nch; k++) { orc_challenger_observe(&ch, open_next[k].c0); orc_challenger_observe(&ch, open_next[k].c1); }
    for (int k = zpp; k < zs_cols; k++) { orc_challenger_observe(&ch, open_next[k].c0); orc_challenger_observe(&ch, open_next[k].c1); }

    /* ---- prove_openings ---- */
    ext_t fri_alpha = orc_challenger_get_ext(&ch);
    /* polynomial lists of the two batches as (oracle, index) */
    int tot0 = NC + R + W + zpp + nch * qdf + (zs_cols - zpp), tot1 = nch + (zs_cols - zpp);
    int (*b0)[2] = malloc(sizeof(int[2]) * (size_t)tot0);
    int (*b1)[2] = malloc(sizeof(int[2]) * (size_t)(tot1 ? tot1 : 1));
    {
        int k = 0;
        for (int i = 0; i < NC + R; i++, k++) { b0[k][0] = 0; b0[k][1] = i; }
        for (int i = 0; i < W; i++, k++) { b0[k][0] = 1; b0[k][1] = i; }
        for (int i = 0; i < zpp; i++, k++) { b0[k][0] = 2; b0[k][1] = i; }
        for (int i = 0; i < nch * qdf; i++, k++) { b0[k][0] = 3; b0[k][1] = i; }
        for (int i = zpp; i < zs_cols; i++, k++) { b0[k][0] = 2; b0[k][1] = i; }
        k = 0;
        for (int i = 0; i < nch; i++, k++) { b1[k][0] = 2; b1[k][1] = i; }
        for (int i = zpp; i < zs_cols; i++, k++) { b1[k][0] = 2; b1[k][1] = i; }
    }
    ext_t* final_poly = (ext_t*)calloc(N, sizeof(ext_t));   /* n coeffs, zero padded to N (lde) */
    for (int b = 0; b < 2; b++) {
        int tot = b ? tot1 : tot0; int (*lst)[2] = b ? b1 : b0;
        ext_t point = b ? zeta_next : zeta;
        ext_t* comp = (ext_t*)calloc(n, sizeof(ext_t));
        /* reduce_polys_base: sum_j alpha^j f_j, Horner from the last polynomial */
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n; k++) {
            ext_t acc = ext_from(0);
            for (int j = tot - 1; j >= 0; j--)
                acc = ext_addb(ext_mul(acc, fri_alpha), oracles[lst[j][0]]->coeffs[(size_t)lst[j][1] * n + k]);
            comp[k] = acc;
        }
        /* divide_by_linear(point), pad back with one zero coefficient */
        ext_t* quot = (ext_t*)calloc(n, sizeof(ext_t));
        ext_t acc = ext_from(0);
        for (size_t k = n; k-- > 0;) { acc = ext_add(ext_mul(acc, point), comp[k]); if (k > 0) quot[k - 1] = acc; }
        /* alpha.shift_poly(final_poly): final *= alpha^count ; final += quotient */
        ext_t sh = ext_pow(fri_alpha, (uint64_t)tot);
        for (size_t k = 0; k < n; k++) final_poly[k] = ext_add(ext_mul(final_poly[k], sh), quot[k]);
        free(comp); free(quot);
    }
    free(b0); free(b1);
    /* lde(rate_bits).coset_fft(7): component-wise */
    gl_t* fv0 = (gl_t*)malloc(N * sizeof(gl_t)); gl_t* fv1 = (gl_t*)malloc(N * sizeof(gl_t));
    for (size_t k = 0; k < N; k++) { fv0[k] = final_poly[k].c0; fv1[k] = final_poly[k].c1; }
    orc_coset_fft(fv0, logN, GL_GENERATOR); orc_coset_fft(fv1, logN, GL_GENERATOR);

    STAGE("fri combine+lde");
    /* ---- fri_committed_trees ---- */
    const int nl = c->num_reduction_arity_bits;
    orc_merkle* trees[16]; gl_t* tree_leaves[16]; size_t tree_nleaves[16];
    ext_t fri_betas[16];
    size_t cur = N; int cur_log = logN;
    ext_t* coeffs = final_poly;     /* length cur */
    gl_t shift = GL_GENERATOR;
    for (int l = 0; l < nl; l++) {
        int ab = c->reduction_arity_bits[l]; size_t arity = (size_t)1 << ab;
        /* reverse_index_bits_in_place(values); leaves = chunks of `arity` ext values flattened */
        gl_t* leaves = (gl_t*)malloc(2 * cur * sizeof(gl_t));
        for (size_t i = 0; i < cur; i++) { size_t j = orc_reverse_bits(i, cur_log); leaves[2 * j] = fv0[i]; leaves[2 * j + 1] = fv1[i]; }
        tree_leaves[l] = leaves; tree_nleaves[l] = cur >> ab;
        trees[l] = orc_merkle_new(leaves, cur >> ab, 2 * arity, c->cap_height);
        orc_challenger_observe_many(&ch, trees[l]->cap, capw);
        wr(&w, trees[l]->cap, capw);
        ext_t beta = orc_challenger_get_ext(&ch);
        fri_betas[l] = beta;
        /* coeffs'[k] = sum_{i<arity} coeffs[arity*k + i] * beta^i */
        size_t nxt = cur >> ab;
        for (size_t k = 0; k < nxt; k++) {
            ext_t acc = ext_from(0);
            for (size_t i = arity; i-- > 0;) acc = ext_add(ext_mul(acc, beta), coeffs[arity * k + i]);
            coeffs[k] = acc;
        }
        cur = nxt; cur_log -= ab;
        shift = gl_pow(shift, arity);
        for (size_t k = 0; k < cur; k++) { fv0[k] = coeffs[k].c0; fv1[k] = coeffs[k].c1; }
        orc_coset_fft(fv0, cur_log, shift); orc_coset_fft(fv1, cur_log, shift);
    }
    size_t final_len = cur >> c->rate_bits;
    for (size_t k = final_len; k < cur; k++) if (coeffs[k].c0 || coeffs[k].c1) { /* should always be zero */ }
    for (size_t k = 0; k < final_len; k++) { orc_challenger_observe(&ch, coeffs[k].c0); orc_challenger_observe(&ch, coeffs[k].c1); }

    STAGE("fri commit phase");
    /* ---- fri_proof_of_work: lowest nonce ---- */
    gl_t pow_witness = 0;
    {
        gl_t st[12]; memcpy(st, ch.state, sizeof(st));
        int pos = ch.in_len;
        for (int i = 0; i < pos; i++) st[i] = ch.in_buf[i];
        /* ascending windows searched in parallel; the minimum hit of the first window with a hit wins */
        const gl_t WIN = 1 << 14;
        int found = 0;
        for (gl_t base = 0; !found; base += WIN) {
            gl_t best = ~(gl_t)0;
#pragma omp parallel for schedule(static) reduction(min : best)
            for (gl_t cand = base; cand < base + WIN; cand++) {
                gl_t s2[12]; memcpy(s2, st, sizeof(st));
                s2[pos] = cand;
                orc_poseidon(s2);
                if ((s2[7] >> (64 - c->pow_bits)) == 0 && cand < best) best = cand;
            }
            if (best != ~(gl_t)0) { pow_witness = best; found = 1; }
        }
        orc_challenger_observe(&ch, pow_witness);
        gl_t resp = orc_challenger_get(&ch);
        if ((resp >> (64 - c->pow_bits)) != 0) return -4;
    }

    STAGE("pow");
    /* ---- query rounds (written after caps; final poly and pow go last) ---- */
    uint64_t qidx[64];
    for (int q = 0; q < c->num_query_rounds; q++) {
        gl_t xq = orc_challenger_get(&ch);
        size_t x_index = (size_t)(xq % N);
        if (q < 64) qidx[q] = x_index;
        gl_t sib[4 * 32];
        for (int o = 0; o < 4; o++) {
            wr(&w, oracles[o]->leaves + x_index * (size_t)ocols[o], (size_t)ocols[o]);
            size_t pl = orc_merkle_path_len(oracles[o]->tree);
            orc_merkle_prove(oracles[o]->tree, x_index, sib);
            wr1(&w, (gl_t)pl); wr(&w, sib, 4 * pl);
        }
        for (int l = 0; l < nl; l++) {
            int ab = c->reduction_arity_bits[l]; size_t arity = (size_t)1 << ab;
            size_t leaf = x_index >> ab;
            wr(&w, tree_leaves[l] + leaf * 2 * arity, 2 * arity);
            size_t pl = orc_merkle_path_len(trees[l]);
            orc_merkle_prove(trees[l], leaf, sib);
            wr1(&w, (gl_t)pl); wr(&w, sib, 4 * pl);
            x_index = leaf;
        }
    }
    for (size_t k = 0; k < final_len; k++) wr_ext(&w, coeffs[k]);
    wr1(&w, pow_witness);
    wr(&w, public_inputs, (size_t)c->num_public_inputs);

    STAGE("queries");
    if (tr) {
        memset(tr, 0, sizeof(*tr));
        memcpy(tr->betas, betas, sizeof(gl_t) * nch); memcpy(tr->gammas, gammas, sizeof(gl_t) * nch);
        memcpy(tr->deltas, deltas, sizeof(gl_t) * 4 * nch); memcpy(tr->alphas, alphas, sizeof(gl_t) * nch);
        tr->zeta = zeta; tr->fri_alpha = fri_alpha;
        for (int l = 0; l < nl; l++) tr->fri_betas[l] = fri_betas[l];
        tr->pow_witness = pow_witness;
        for (int q = 0; q < c->num_query_rounds && q < 64; q++) tr->query_indices[q] = qidx[q];
    }
    for (int l = 0; l < nl; l++) { orc_merkle_free(trees[l]); free(tree_leaves[l]); }
    (void)tree_nleaves;
    for (int o = 0; o < 4; o++) free(open_zeta[o]);
    free(open_next); free(final_poly); free(fv0); free(fv1);
    orc_batch_free(wb); orc_batch_free(zb); orc_batch_free(qb);
    if (w.pos > proof_cap) return -5;
    return (long)w.pos;
}

/* ================================== verifier ================================================ */
typedef struct { const gl_t* p; size_t pos, len; int bad; } rd_t;
static const gl_t* rd(rd_t* r, size_t n) {
    if (r->pos + n > r->len) { r->bad = 1; return r->p; }
    const gl_t* q = r->p + r->pos; r->pos += n; return q;
}
static ext_t rd_ext_at(const gl_t* p, size_t k) { return ext_make(p[2 * k], p[2 * k + 1]); }

/* fri/verifier.rs compute_evaluation: interpolate the arity points of the coset and evaluate at beta */
static ext_t compute_evaluation(gl_t x, size_t x_index_within_coset, int arity_bits, const ext_t* evals_in, ext_t beta) {
    size_t arity = (size_t)1 << arity_bits;
    gl_t g = gl_root_of_unity(arity_bits);
    ext_t evals[64]; gl_t pts[64];
    for (size_t i = 0; i < arity; i++) evals[orc_reverse_bits(i, arity_bits)] = evals_in[i];
    size_t rev = orc_reverse_bits(x_index_within_coset, arity_bits);
    gl_t start = gl_mul(x, gl_pow(g, arity - rev));
    gl_t y = 1;
    for (size_t i = 0; i < arity; i++) { pts[i] = gl_mul(start, y); y = gl_mul(y, g); }
    /* Lagrange interpolation at beta (points are base field, values ext) */
    ext_t res = ext_from(0);
    for (size_t i = 0; i < arity; i++) {
        ext_t num = ext_from(1); gl_t den = 1;
        for (size_t j = 0; j < arity; j++) if (j != i) {
            num = ext_mul(num, ext_subb(beta, pts[j]));
            den = gl_mul(den, gl_sub(pts[i], pts[j]));
        }
        res = ext_add(res, ext_mul(evals[i], ext_mul_base(num, gl_inv(den))));
    }
    return res;
}

int orc_verify(const orc_prover_data* pd, const gl_t* proof, size_t proof_len) {
    const orc_circuit* c = &pd->c;
    const int nch = c->num_challenges, R = c->num_routed_wires, W = c->num_wires, qdf = c->quotient_degree_factor;
    const int num_prods = c->num_partial_products, nlp = cfg_num_lookup_polys(c), NC = cfg_nc(c);
    const int has_lookup = c->num_luts > 0;
    const int logn = c->degree_bits, logN = logn + c->rate_bits;
    const size_t n = (size_t)1 << logn, N = (size_t)1 << logN;
    const int zs_cols = cfg_zs_cols(c), zpp = nch * (1 + num_prods), nlz = zs_cols - zpp;
    const size_t capw = (size_t)4 << c->cap_height;
    const int nl = c->num_reduction_arity_bits;
    if (proof_len != orc_proof_len(c)) return -1;
    rd_t r = {proof, 0, proof_len, 0};
    const gl_t* wires_cap = rd(&r, capw); const gl_t* zs_cap = rd(&r, capw); const gl_t* q_cap = rd(&r, capw);
    const gl_t* o_const = rd(&r, 2 * (size_t)NC); const gl_t* o_sig = rd(&r, 2 * (size_t)R); const gl_t* o_wires = rd(&r, 2 * (size_t)W);
    const gl_t* o_zs = rd(&r, 2 * (size_t)nch); const gl_t* o_zs_next = rd(&r, 2 * (size_t)nch);
    const gl_t* o_pp = rd(&r, 2 * (size_t)nch * num_prods); const gl_t* o_q = rd(&r, 2 * (size_t)nch * qdf);
    const gl_t* o_lz = rd(&r, 2 * (size_t)nlz); const gl_t* o_lz_next = rd(&r, 2 * (size_t)nlz);
    const gl_t* fri_caps = rd(&r, (size_t)nl * capw);
    size_t queries_pos = r.pos;
    /* skip queries to reach final poly / pow / public inputs */
    size_t per_q = 0; { int cols[4] = {NC + R, W, zs_cols, nch * qdf};
        for (int o = 0; o < 4; o++) per_q += cols[o] + 1 + 4 * (size_t)(logN - c->cap_height);
        int lg = logN; for (int l = 0; l < nl; l++) { int ab = c->reduction_arity_bits[l]; lg -= ab; per_q += (2u << ab) + 1 + 4 * (size_t)(lg - c->cap_height); } }
    rd(&r, per_q * c->num_query_rounds);
    size_t final_len = fri_final_len(c);
    const gl_t* final_poly = rd(&r, 2 * final_len);
    gl_t pow_witness = *rd(&r, 1);
    const gl_t* public_inputs = rd(&r, (size_t)c->num_public_inputs);
    if (r.bad) return -1;

    /* ---- challenges (get_challenges.rs) ---- */
    gl_t pi_hash[4]; orc_hash_no_pad(public_inputs, (size_t)c->num_public_inputs, pi_hash);
    orc_challenger ch; orc_challenger_init(&ch);
    orc_challenger_observe_many(&ch, c->circuit_digest, 4);
    orc_challenger_observe_many(&ch, pi_hash, 4);
    orc_challenger_observe_many(&ch, wires_cap, capw);
    gl_t betas[4], gammas[4], deltas[16], alphas[4];
    for (int i = 0; i < nch; i++) betas[i] = orc_challenger_get(&ch);
    for (int i = 0; i < nch; i++) gammas[i] = orc_challenger_get(&ch);
    if (has_lookup) { int k = 0; for (int i = 0; i < nch; i++) deltas[k++] = betas[i]; for (int i = 0; i < nch; i++) deltas[k++] = gammas[i];
        for (int i = 0; i < 2 * nch; i++) deltas[k++] = orc_challenger_get(&ch); }
    orc_challenger_observe_many(&ch, zs_cap, capw);
    for (int i = 0; i < nch; i++) alphas[i] = orc_challenger_get(&ch);
    orc_challenger_observe_many(&ch, q_cap, capw);
    ext_t zeta = orc_challenger_get_ext(&ch);
    /* observe openings: batch zeta = constants, sigmas, wires, zs, pps, quotient, lookup_zs ; batch next = zs_next, lookup_zs_next */
    orc_challenger_observe_many(&ch, o_const, 2 * (size_t)NC); orc_challenger_observe_many(&ch, o_sig, 2 * (size_t)R);
    orc_challenger_observe_many(&ch, o_wires, 2 * (size_t)W); orc_challenger_observe_many(&ch, o_zs, 2 * (size_t)nch);
    orc_challenger_observe_many(&ch, o_pp, 2 * (size_t)nch * num_prods); orc_challenger_observe_many(&ch, o_q, 2 * (size_t)nch * qdf);
    orc_challenger_observe_many(&ch, o_lz, 2 * (size_t)nlz);
    orc_challenger_observe_many(&ch, o_zs_next, 2 * (size_t)nch); orc_challenger_observe_many(&ch, o_lz_next, 2 * (size_t)nlz);
    ext_t fri_alpha = orc_challenger_get_ext(&ch);
    ext_t fri_betas[16];
    for (int l = 0; l < nl; l++) { orc_challenger_observe_many(&ch, fri_caps + (size_t)l * capw, capw); fri_betas[l] = orc_challenger_get_ext(&ch); }
    orc_challenger_observe_many(&ch, final_poly, 2 * final_len);
    orc_challenger_observe(&ch, pow_witness);
    gl_t pow_resp = orc_challenger_get(&ch);
    if ((pow_resp >> (64 - c->pow_bits)) != 0) return -10;

    /* ---- vanishing(zeta) == Z_H(zeta) * t(zeta)  (plonk/verifier.rs) ---- */
    {
        gl_t lut_evals[64]; if (has_lookup) compute_lut_evals(c, deltas, lut_evals);
        ve_ctx vc = {c, betas, gammas, alphas, deltas, lut_evals, pi_hash};
        ext_t* consts = malloc(sizeof(ext_t) * (size_t)NC); ext_t* wv = malloc(sizeof(ext_t) * (size_t)W); ext_t* sg = malloc(sizeof(ext_t) * (size_t)R);
        ext_t zsv[4], znv[4]; ext_t* pp = malloc(sizeof(ext_t) * (size_t)(nch * num_prods + 1));
        ext_t* lz = malloc(sizeof(ext_t) * (size_t)(nlz + 1)); ext_t* lzn = malloc(sizeof(ext_t) * (size_t)(nlz + 1));
        for (int k = 0; k < NC; k++) consts[k] = rd_ext_at(o_const, k);
        for (int k = 0; k < W; k++) wv[k] = rd_ext_at(o_wires, k);
        for (int k = 0; k < R; k++) sg[k] = rd_ext_at(o_sig, k);
        for (int k = 0; k < nch; k++) { zsv[k] = rd_ext_at(o_zs, k); znv[k] = rd_ext_at(o_zs_next, k); }
        for (int k = 0; k < nch * num_prods; k++) pp[k] = rd_ext_at(o_pp, k);
        for (int k = 0; k < nlz; k++) { lz[k] = rd_ext_at(o_lz, k); lzn[k] = rd_ext_at(o_lz_next, k); }
        ext_t zeta_pow = ext_exp_pow2(zeta, logn);
        ext_t z_h = ext_subb(zeta_pow, 1);
        /* eval_l_0(n, x) = (x^n - 1) / (n (x - 1)) */
        ext_t l0 = ext_mul(z_h, ext_inv(ext_mul_base(ext_subb(zeta, 1), (gl_t)n)));
        ext_t van[4];
        ve_eval(&vc, zeta, l0, consts, wv, zsv, znv, pp, lz, lzn, sg, van);
        int ok = 1;
        for (int i = 0; i < nch; i++) {
            ext_t acc = ext_from(0);
            for (int k = qdf - 1; k >= 0; k--) acc = ext_add(ext_mul(acc, zeta_pow), rd_ext_at(o_q, (size_t)i * qdf + k));
            if (!ext_eq(van[i], ext_mul(z_h, acc))) ok = 0;
        }
        free(consts); free(wv); free(sg); free(pp); free(lz); free(lzn);
        if (!ok) return -20;
    }

    /* ---- verify_fri_proof ---- */
    gl_t g = gl_root_of_unity(logn);
    ext_t zeta_next = ext_mul_base(zeta, g);
    int tot0 = NC + R + W + zpp + nch * qdf + nlz, tot1 = nch + nlz;
    /* PrecomputedReducedOpenings: reduce(batch values) with alpha, Horner from the end */
    ext_t red0 = ext_from(0), red1 = ext_from(0);
    {
        ext_t* v0 = malloc(sizeof(ext_t) * (size_t)tot0); int k = 0;
        for (int i = 0; i < NC; i++) v0[k++] = rd_ext_at(o_const, i);
        for (int i = 0; i < R; i++) v0[k++] = rd_ext_at(o_sig, i);
        for (int i = 0; i < W; i++) v0[k++] = rd_ext_at(o_wires, i);
        for (int i = 0; i < nch; i++) v0[k++] = rd_ext_at(o_zs, i);
        for (int i = 0; i < nch * num_prods; i++) v0[k++] = rd_ext_at(o_pp, i);
        for (int i = 0; i < nch * qdf; i++) v0[k++] = rd_ext_at(o_q, i);
        for (int i = 0; i < nlz; i++) v0[k++] = rd_ext_at(o_lz, i);
        for (int i = tot0 - 1; i >= 0; i--) red0 = ext_add(ext_mul(red0, fri_alpha), v0[i]);
        free(v0);
        ext_t* v1 = malloc(sizeof(ext_t) * (size_t)(tot1 + 1)); k = 0;
        for (int i = 0; i < nch; i++) v1[k++] = rd_ext_at(o_zs_next, i);
        for (int i = 0; i < nlz; i++) v1[k++] = rd_ext_at(o_lz_next, i);
        for (int i = tot1 - 1; i >= 0; i--) red1 = ext_add(ext_mul(red1, fri_alpha), v1[i]);
        free(v1);
    }
    const gl_t* caps[4] = {pd->cs->tree->cap, wires_cap, zs_cap, q_cap};
    const int ocols[4] = {NC + R, W, zs_cols, nch * qdf};
    rd_t q = {proof, queries_pos, proof_len, 0};
    ext_t final_c[256];
    if (final_len > 256) return -1;
    for (size_t k = 0; k < final_len; k++) final_c[k] = rd_ext_at(final_poly, k);
    for (int round = 0; round < c->num_query_rounds; round++) {
        size_t x_index = (size_t)(orc_challenger_get(&ch) % N);
        const gl_t* evals[4];
        for (int o = 0; o < 4; o++) {
            evals[o] = rd(&q, (size_t)ocols[o]);
            size_t pl = (size_t)*rd(&q, 1);
            if (pl != (size_t)(logN - c->cap_height)) return -30;
            const gl_t* sib = rd(&q, 4 * pl);
            if (!orc_merkle_verify(evals[o], (size_t)ocols[o], x_index, caps[o], c->cap_height, sib, pl)) return -31;
        }
        gl_t subgroup_x = gl_mul(GL_GENERATOR, gl_pow(gl_root_of_unity(logN), orc_reverse_bits(x_index, logN)));
        /* fri_combine_initial */
        ext_t sum = ext_from(0);
        for (int b = 0; b < 2; b++) {
            ext_t acc = ext_from(0);
            if (b == 0) {
                /* reverse order of batch0 = [oracle0.., oracle1.., oracle2[0..zpp], oracle3.., oracle2[zpp..]] */
                for (int i = zs_cols - 1; i >= zpp; i--) acc = ext_addb(ext_mul(acc, fri_alpha), evals[2][i]);
                for (int i = nch * qdf - 1; i >= 0; i--) acc = ext_addb(ext_mul(acc, fri_alpha), evals[3][i]);
                for (int i = zpp - 1; i >= 0; i--) acc = ext_addb(ext_mul(acc, fri_alpha), evals[2][i]);
                for (int i = W - 1; i >= 0; i--) acc = ext_addb(ext_mul(acc, fri_alpha), evals[1][i]);
                for (int i = NC + R - 1; i >= 0; i--) acc = ext_addb(ext_mul(acc, fri_alpha), evals[0][i]);
            } else {
                for (int i = zs_cols - 1; i >= zpp; i--) acc = ext_addb(ext_mul(acc, fri_alpha), evals[2][i]);
                for (int i = nch - 1; i >= 0; i--) acc = ext_addb(ext_mul(acc, fri_alpha), evals[2][i]);
            }
            ext_t numerator = ext_sub(acc, b ? red1 : red0);
            ext_t denominator = ext_bsub(subgroup_x, b ? zeta_next : zeta);
            sum = ext_mul(sum, ext_pow(fri_alpha, (uint64_t)(b ? tot1 : tot0)));
            sum = ext_add(sum, ext_mul(numerator, ext_inv(denominator)));
        }
        ext_t old_eval = sum;
        for (int l = 0; l < nl; l++) {
            int ab = c->reduction_arity_bits[l]; size_t arity = (size_t)1 << ab;
            const gl_t* ev = rd(&q, 2 * arity);
            size_t pl = (size_t)*rd(&q, 1);
            const gl_t* sib = rd(&q, 4 * pl);
            if (q.bad) return -1;
            size_t coset_index = x_index >> ab, within = x_index & (arity - 1);
            ext_t evs[64];
            for (size_t i = 0; i < arity; i++) evs[i] = rd_ext_at(ev, i);
            if (!ext_eq(evs[within], old_eval)) return -40 - l;
            old_eval = compute_evaluation(subgroup_x, within, ab, evs, fri_betas[l]);
            if (!orc_merkle_verify(ev, 2 * arity, coset_index, fri_caps + (size_t)l * capw, c->cap_height, sib, pl)) return -50 - l;
            for (int k = 0; k < ab; k++) subgroup_x = gl_sqr(subgroup_x);
            x_index = coset_index;
        }
        if (!ext_eq(eval_poly_ext(final_c, final_len, ext_from(subgroup_x)), old_eval)) return -60;
    }
    if (q.bad) return -1;
    return 0;
}
