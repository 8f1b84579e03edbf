/* TEST INFRASTRUCTURE — CPU oracle.  Restates plonky2/src/plonk/vanishing_poly.rs
 * (eval_vanishing_poly / eval_vanishing_poly_base_batch, check_lookup_constraints,
 * evaluate_gate_constraints) and plonk/plonk_common.rs (check_partial_products,
 * reduce_with_powers_multi) of the dependency pinned at /root/reference/Cargo.toml:12.
 * PARITY UNPINNED (see oracle.h).  Included twice by prover.c: once over the base field
 * (prover, LDE points) and once over the quadratic extension (verifier, point zeta).
 *
 * Required macros: VNAME(x), FT, F_ZERO, F_ONE, F_ADD, F_SUB, F_MUL, F_MULB (FT * gl_t),
 * F_FROMB (gl_t -> FT), F_ADDB (FT + gl_t), F_SUBB (FT - gl_t), F_BSUB (gl_t - FT)
 */

typedef struct {
    const orc_circuit* c;
    const gl_t* betas; const gl_t* gammas; const gl_t* alphas; const gl_t* deltas; /* deltas: [nch][4] */
    const gl_t* lut_evals;   /* [nch][num_luts]: get_lut_poly(...).eval(delta) */
    const gl_t* pi_hash;     /* [4] */
} VNAME(ctx);

/* gate.rs compute_filter: prod_{i in group, i != row} (i - s) * (UNUSED - s if many selectors) */
static FT VNAME(filter)(int row, int gs, int ge, FT s, int many) {
    FT f = F_ONE;
    for (int i = gs; i < ge; i++) if (i != row) f = F_MUL(f, F_BSUB((gl_t)i, s));
    if (many) f = F_MUL(f, F_BSUB((gl_t)0xFFFFFFFFULL, s));
    return f;
}

/* out[nch].  consts: all constant columns (selectors, lookup selectors, gate constants);
 * zs/next_zs [nch]; pps [nch*num_prods]; lzs/next_lzs [nch*num_lookup_polys]; sig [routed]. */
static void VNAME(eval)(const VNAME(ctx)* v, FT x, FT l0x, const FT* consts, const FT* wires, const FT* zs,
                        const FT* next_zs, const FT* pps, const FT* lzs, const FT* next_lzs, const FT* sig, FT* out) {
    const orc_circuit* c = v->c;
    const int nch = c->num_challenges, R = c->num_routed_wires, qdf = c->quotient_degree_factor;
    const int num_prods = c->num_partial_products;
    const int has_lookup = c->num_luts > 0;
    const int lu_slots = R / 2, lut_slots = R / 3;
    const int lu_degree = qdf - 1;
    const int num_sldc = has_lookup ? (lu_slots + lu_degree - 1) / lu_degree : 0;
    const int num_lookup_polys = has_lookup ? num_sldc + 1 : 0;
    const int lut_degree = has_lookup ? (lut_slots + num_sldc - 1) / num_sldc : 0;
    const int n_lookup_terms = has_lookup ? 4 + c->num_luts + 2 * num_sldc : 0;
    const int nterms = nch + nch * (num_prods + 1) + nch * n_lookup_terms + c->num_gate_constraints;
    FT terms[640];
    int t = 0;
    /* Z(x) - 1 terms */
    for (int i = 0; i < nch; i++) terms[t++] = F_MUL(l0x, F_SUBB(zs[i], 1));
    /* partial product checks */
    for (int i = 0; i < nch; i++) {
        FT prev = zs[i];
        for (int ck = 0; ck <= num_prods; ck++) {
            FT np = F_ONE, dp = F_ONE;
            int lo = ck * qdf, hi = lo + qdf < R ? lo + qdf : R;
            for (int j = lo; j < hi; j++) {
                FT sid = F_MULB(x, c->k_is[j]);
                FT num = F_ADDB(F_ADD(wires[j], F_MULB(sid, v->betas[i])), v->gammas[i]);
                FT den = F_ADDB(F_ADD(wires[j], F_MULB(sig[j], v->betas[i])), v->gammas[i]);
                np = F_MUL(np, num); dp = F_MUL(dp, den);
            }
            FT next = ck == num_prods ? next_zs[i] : pps[i * num_prods + ck];
            terms[t++] = F_SUB(F_MUL(prev, np), F_MUL(next, dp));
            prev = next;
        }
    }
    /* lookup terms */
    if (has_lookup) {
        const FT* lsel = consts + c->num_selectors;
        for (int i = 0; i < nch; i++) {
            const gl_t da = v->deltas[4 * i + 0], db = v->deltas[4 * i + 1], dalpha = v->deltas[4 * i + 2], ddelta = v->deltas[4 * i + 3];
            const FT* lz = lzs + i * num_lookup_polys; const FT* nlz = next_lzs + i * num_lookup_polys;
            FT z_re = lz[0], next_z_re = nlz[0];
            const FT* sl = lz + 1; const FT* nsl = nlz + 1;
            FT looked[32], looking[64], lookupc[32];
            for (int s = 0; s < lut_slots; s++) {
                FT in = wires[3 * s], o = wires[3 * s + 1];
                looked[s] = F_ADD(in, F_MULB(o, da));
                lookupc[s] = F_ADD(in, F_MULB(o, db));
            }
            for (int s = 0; s < lu_slots; s++) looking[s] = F_ADD(wires[2 * s], F_MULB(wires[2 * s + 1], da));
            terms[t++] = F_MUL(lsel[3], sl[num_sldc - 1]);           /* LastLdc */
            terms[t++] = F_MUL(lsel[2], sl[0]);                      /* InitSre * first SLDC */
            terms[t++] = F_MUL(lsel[2], z_re);                       /* InitSre * RE */
            for (int r = 0; r < c->num_luts; r++)
                terms[t++] = F_MUL(lsel[4 + r], F_SUBB(z_re, v->lut_evals[i * c->num_luts + r]));
            FT cur = next_z_re;
            for (int s = 0; s < lut_slots; s++) cur = F_ADD(F_MULB(cur, ddelta), lookupc[s]);
            terms[t++] = F_MUL(lsel[0], F_SUB(z_re, cur));           /* TransSre * RE transition */
            for (int poly = 0; poly < num_sldc; poly++) {
                int a0 = poly * lut_degree, a1 = a0 + lut_degree < lut_slots ? a0 + lut_degree : lut_slots;
                int b0 = poly * lu_degree, b1 = b0 + lu_degree < lu_slots ? b0 + lu_degree : lu_slots;
                FT lut_prod = F_ONE, lu_prod = F_ONE;
                for (int k = a0; k < a1; k++) lut_prod = F_MUL(lut_prod, F_BSUB(dalpha, looked[k]));
                for (int k = b0; k < b1; k++) lu_prod = F_MUL(lu_prod, F_BSUB(dalpha, looking[k]));
                FT lu_sum = F_ZERO, lut_sum_mul = F_ZERO;
                for (int k = b0; k < b1; k++) {
                    FT pr = F_ONE;
                    for (int j = b0; j < b1; j++) if (j != k) pr = F_MUL(pr, F_BSUB(dalpha, looking[j]));
                    lu_sum = F_ADD(lu_sum, pr);
                }
                for (int k = a0; k < a1; k++) {
                    FT pr = F_ONE;
                    for (int j = a0; j < a1; j++) if (j != k) pr = F_MUL(pr, F_BSUB(dalpha, looked[j]));
                    lut_sum_mul = F_ADD(lut_sum_mul, F_MUL(wires[3 * k + 2], pr));
                }
                FT prev = poly == 0 ? nsl[num_sldc - 1] : sl[poly - 1];
                FT diff = F_SUB(sl[poly], prev);
                terms[t++] = F_MUL(lsel[0], F_SUB(F_MUL(lut_prod, diff), lut_sum_mul));   /* Sum transition */
                terms[t++] = F_MUL(lsel[1], F_ADD(F_MUL(lu_prod, diff), lu_sum));         /* LDC transition */
            }
        }
    }
    /* gate constraints */
    {
        FT* gc = terms + t;
        for (int k = 0; k < c->num_gate_constraints; k++) gc[k] = F_ZERO;
        const FT* gconst = consts + c->num_selectors + c->num_lookup_selectors;
        for (int g = 0; g < c->num_gates; g++) {
            const orc_gate* G = &c->gates[g];
            if (G->num_constraints == 0) continue;
            FT f = VNAME(filter)(g, G->group_start, G->group_end, consts[G->selector_index], c->num_selectors > 1);
            switch (G->kind) {
            case ORC_GATE_ARITHMETIC:
                for (int k = 0; k < G->param0; k++) {
                    FT m0 = wires[4 * k], m1 = wires[4 * k + 1], ad = wires[4 * k + 2], o = wires[4 * k + 3];
                    FT comp = F_ADD(F_MUL(F_MUL(m0, m1), gconst[0]), F_MUL(ad, gconst[1]));
                    gc[k] = F_ADD(gc[k], F_MUL(f, F_SUB(o, comp)));
                }
                break;
            case ORC_GATE_CONSTANT:
                for (int k = 0; k < G->param0; k++) gc[k] = F_ADD(gc[k], F_MUL(f, F_SUB(gconst[k], wires[k])));
                break;
            case ORC_GATE_PUBLIC_INPUT:
                for (int k = 0; k < 4; k++) gc[k] = F_ADD(gc[k], F_MUL(f, F_SUBB(wires[k], v->pi_hash[k])));
                break;
            case ORC_GATE_POSEIDON: {
                /* gates/poseidon.rs eval_unfiltered: wires 0..11 input, 12..23 output, 24 swap, 25..28 delta,
                 * 29..64 full-round S-box inputs (rounds 1..3), 65..86 partial S-box inputs, 87..134 second
                 * full rounds; the partial rounds are constrained in their sparse form */
                FT st[12]; int k = 0;
                FT swap = wires[24];
                gc[k] = F_ADD(gc[k], F_MUL(f, F_MUL(swap, F_SUBB(swap, 1)))); k++;
                for (int i = 0; i < 4; i++) {
                    FT d = wires[25 + i];
                    gc[k] = F_ADD(gc[k], F_MUL(f, F_SUB(F_MUL(swap, F_SUB(wires[i + 4], wires[i])), d))); k++;
                    st[i] = F_ADD(wires[i], d); st[i + 4] = F_SUB(wires[i + 4], d);
                }
                for (int i = 8; i < 12; i++) st[i] = wires[i];
#define PG_SBOX(x) do { FT x2_ = F_MUL(x, x), x4_ = F_MUL(x2_, x2_), x3_ = F_MUL(x, x2_); x = F_MUL(x3_, x4_); } while (0)
#define PG_MDS() do { FT o_[12]; for (int r_ = 0; r_ < 12; r_++) { FT a_ = F_ZERO; for (int i_ = 0; i_ < 12; i_++) a_ = F_ADD(a_, F_MULB(st[(i_ + r_) % 12], VAN_MDS_CIRC[i_])); \
                      if (r_ == 0) a_ = F_ADD(a_, F_MULB(st[0], 8)); o_[r_] = a_; } for (int r_ = 0; r_ < 12; r_++) st[r_] = o_[r_]; } while (0)
                for (int r = 0; r < 4; r++) {
                    for (int i = 0; i < 12; i++) st[i] = F_ADDB(st[i], VAN_RC[12 * r + i]);
                    if (r != 0) for (int i = 0; i < 12; i++) {
                        FT in = wires[29 + 12 * (r - 1) + i];
                        gc[k] = F_ADD(gc[k], F_MUL(f, F_SUB(st[i], in))); k++; st[i] = in;
                    }
                    for (int i = 0; i < 12; i++) PG_SBOX(st[i]);
                    PG_MDS();
                }
                for (int i = 0; i < 12; i++) st[i] = F_ADDB(st[i], PFAST_FIRST_C[i]);
                { FT t_[11]; for (int r = 0; r < 11; r++) { FT a = F_ZERO; for (int cc = 0; cc < 11; cc++) a = F_ADD(a, F_MULB(st[cc + 1], PFAST_INIT[r * 11 + cc])); t_[r] = a; }
                  for (int r = 0; r < 11; r++) st[r + 1] = t_[r]; }
                for (int r = 0; r < 22; r++) {
                    FT in = wires[65 + r];
                    gc[k] = F_ADD(gc[k], F_MUL(f, F_SUB(st[0], in))); k++;
                    st[0] = in; PG_SBOX(st[0]);
                    st[0] = F_ADDB(st[0], PFAST_K[r]);
                    FT s0 = F_MULB(st[0], 25);
                    for (int j = 0; j < 11; j++) s0 = F_ADD(s0, F_MULB(st[j + 1], PFAST_VROW[r * 11 + j]));
                    for (int j = 0; j < 11; j++) st[j + 1] = F_ADD(st[j + 1], F_MULB(st[0], PFAST_WCOL[r * 11 + j]));
                    st[0] = s0;
                }
                for (int r = 0; r < 4; r++) {
                    for (int i = 0; i < 12; i++) st[i] = F_ADDB(st[i], VAN_RC[12 * (26 + r) + i]);
                    for (int i = 0; i < 12; i++) {
                        FT in = wires[87 + 12 * r + i];
                        gc[k] = F_ADD(gc[k], F_MUL(f, F_SUB(st[i], in))); k++; st[i] = in;
                    }
                    for (int i = 0; i < 12; i++) PG_SBOX(st[i]);
                    PG_MDS();
                }
                for (int i = 0; i < 12; i++) { gc[k] = F_ADD(gc[k], F_MUL(f, F_SUB(st[i], wires[12 + i]))); k++; }
#undef PG_SBOX
#undef PG_MDS
                break; }
            default: break;
            }
        }
        t += c->num_gate_constraints;
    }
    (void)nterms;
    /* reduce_with_powers_multi: sum_k terms[k] * alpha_i^k */
    for (int i = 0; i < nch; i++) {
        FT acc = F_ZERO;
        for (int k = t - 1; k >= 0; k--) acc = F_ADD(F_MULB(acc, v->alphas[i]), terms[k]);
        out[i] = acc;
    }
}
