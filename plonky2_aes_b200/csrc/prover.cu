// C ABI of libp2gpu.so, part 2: circuit upload and the whole prove() path
// (plonk/prover.rs::prove_with_partition_witness of the dependency pinned at
// /root/reference/Cargo.toml:12; reference call sites: every `data.prove(pw)`, e.g.
// /root/reference/aes-gcm/src/circuit_gcm.rs:781).  Everything O(n) or larger runs on the
// device; the host only drives the Fiat-Shamir transcript (a few hundred Poseidon permutations
// per proof), inverts the 64..256-point final FRI layer and assembles the proof words.
#include "ctx.h"
#include "prover_kernels.cuh"
#include <string.h>
#include <stdio.h>
#include <algorithm>
#include <stdlib.h>
#include <thread>
#include <functional>

// ---------------------------------------------------------------------------------------------
// host transcript: Challenger<F, PoseidonHash> (iop/challenger.rs)
// ---------------------------------------------------------------------------------------------
struct Challenger {
    gl_t state[12]; gl_t in_buf[8]; int in_len; gl_t out_buf[8]; int out_len;
    Challenger() { memset(this, 0, sizeof(*this)); }
    void duplexing() {
        for (int i = 0; i < in_len; i++) state[i] = in_buf[i];
        in_len = 0;
        poseidon_permute(state);
        memcpy(out_buf, state, sizeof(out_buf)); out_len = 8;
    }
    void observe(gl_t x) { out_len = 0; in_buf[in_len++] = x; if (in_len == 8) duplexing(); }
    void observe_many(const gl_t* x, size_t n) { for (size_t i = 0; i < n; i++) observe(x[i]); }
    gl_t get() { if (in_len != 0 || out_len == 0) duplexing(); return out_buf[--out_len]; }
    ext_t get_ext() { ext_t r; r.c0 = get(); r.c1 = get(); return r; }
};
static void host_hash_no_pad(const gl_t* in, size_t n, gl_t out[4]) {
    gl_t s[12] = {0};
    for (size_t off = 0; off < n; off += 8) {
        size_t len = std::min<size_t>(8, n - off);
        memcpy(s, in + off, len * sizeof(gl_t));
        poseidon_permute(s);
    }
    memcpy(out, s, 4 * sizeof(gl_t));
}

// ---------------------------------------------------------------------------------------------
struct p2g_circuit {
    p2g_circuit_desc d;
    std::vector<p2g_gate> gates; std::vector<int32_t> lut_lens, lookup_rows; std::vector<uint16_t> lut_data;
    std::vector<gl_t> k_is;
    CircuitDev cd;
    p2g_batch* cs = nullptr;            // constants_sigmas commitment
    gl_t* d_sigmas = nullptr;           // [R][n] values on H
    gl_t* d_subgroup = nullptr;         // g^i
    gl_t* d_domain = nullptr;           // x_j = 7 w_N^bitrev(j)
    gl_t* d_l0inv = nullptr;            // 1 / (n (x_j - 1))
    gl_t* d_qtable = nullptr;           // [8][n]: h_s^(-j)/8
    gl_t* d_small = nullptr;            // w8inv_pows[8], shift_n_inv_pows[8]
    p2g_gate* d_gates = nullptr;
    uint8_t* d_row_kind = nullptr;
    uint16_t* d_lut_data = nullptr; int* d_lut_off = nullptr; int* d_lut_len = nullptr;
    size_t proof_words = 0;
    int final_len = 0;
};
// releases a circuit whose construction stopped half way (every error return of p2g_circuit_load)
struct CircuitGuard {
    p2g_ctx* ctx; p2g_circuit* c; gl_t* tmp;
    ~CircuitGuard() { if (tmp) ctx_free(ctx, tmp); if (c) p2g_circuit_free(ctx, c); }
};

static int num_lookup_polys(const p2g_circuit_desc& d) {
    if (d.num_luts == 0) return 0;
    int lu_slots = d.num_routed_wires / 2, lu_degree = d.quotient_degree_factor - 1;
    return (lu_slots + lu_degree - 1) / lu_degree + 1;
}
static size_t proof_len(const p2g_circuit_desc& d) {
    size_t cap = (size_t)4 << d.cap_height;
    int nlp = num_lookup_polys(d), nch = d.num_challenges;
    int NC = d.num_selectors + d.num_lookup_selectors + d.num_constants;
    int logN = d.degree_bits + d.rate_bits;
    int zs_cols = nch * (1 + d.num_partial_products + nlp);
    size_t open = 2 * ((size_t)NC + d.num_routed_wires + d.num_wires + 2 * nch + (size_t)nch * d.num_partial_products +
                       (size_t)nch * d.quotient_degree_factor + 2 * (size_t)nch * nlp);
    size_t len = 3 * cap + open + (size_t)d.num_reduction_arity_bits * cap;
    size_t per_q = 0;
    int cols[4] = {NC + d.num_routed_wires, d.num_wires, zs_cols, nch * d.quotient_degree_factor};
    for (int o = 0; o < 4; o++) per_q += cols[o] + 1 + 4 * (size_t)(logN - d.cap_height);
    int lg = logN;
    size_t fin = (size_t)1 << d.degree_bits;
    for (int l = 0; l < d.num_reduction_arity_bits; l++) {
        int ab = d.reduction_arity_bits[l];
        lg -= ab; fin >>= ab;
        per_q += (2u << ab) + 1 + 4 * (size_t)(lg - d.cap_height);
    }
    return len + per_q * d.num_query_rounds + 2 * fin + 1 + d.num_public_inputs;
}

// ---------------------------------------------------------------------------------------------
// ProofWithPublicInputs::to_bytes / from_bytes (util/serialization/mod.rs): the flat proof words
//  <-> upstream's wire format.  Pure host code, needs no device.
// Walks the proof in WORD order and reports for every segment where it goes in BYTE order.
// ---------------------------------------------------------------------------------------------
struct ProofSeg { size_t word_pos, words; int kind; };        // kind 0: field words, 1: Merkle path length (u8)
static void proof_segments(const p2g_circuit_desc& d, std::vector<ProofSeg>& word_order, std::vector<int>& byte_order) {
    const size_t cap = (size_t)4 << d.cap_height;
    const int nlp = num_lookup_polys(d), nch = d.num_challenges;
    const int NC = d.num_selectors + d.num_lookup_selectors + d.num_constants;
    const int logN = d.degree_bits + d.rate_bits;
    const int zs_cols = nch * (1 + d.num_partial_products + nlp);
    size_t pos = 0;
    auto seg = [&](size_t words, int kind) { word_order.push_back({pos, words, kind}); pos += words; return (int)word_order.size() - 1; };
    for (int i = 0; i < 3; i++) byte_order.push_back(seg(cap, 0));
    // openings, word order: constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products,
    // quotient_polys, lookup_zs, lookup_zs_next; upstream writes the two lookup vectors right after plonk_zs_next
    const size_t ow[9] = {(size_t)NC, (size_t)d.num_routed_wires, (size_t)d.num_wires, (size_t)nch, (size_t)nch,
                          (size_t)nch * d.num_partial_products, (size_t)nch * d.quotient_degree_factor, (size_t)nch * nlp, (size_t)nch * nlp};
    int os[9];
    for (int i = 0; i < 9; i++) os[i] = seg(2 * ow[i], 0);
    const int oorder[9] = {0, 1, 2, 3, 4, 7, 8, 5, 6};
    for (int i = 0; i < 9; i++) byte_order.push_back(os[oorder[i]]);
    for (int l = 0; l < d.num_reduction_arity_bits; l++) byte_order.push_back(seg(cap, 0));
    const int cols[4] = {NC + d.num_routed_wires, d.num_wires, zs_cols, nch * d.quotient_degree_factor};
    for (int q = 0; q < d.num_query_rounds; q++) {
        for (int o = 0; o < 4; o++) {
            byte_order.push_back(seg((size_t)cols[o], 0));
            byte_order.push_back(seg(1, 1));
            byte_order.push_back(seg(4 * (size_t)(logN - d.cap_height), 0));
        }
        int lg = logN;
        for (int l = 0; l < d.num_reduction_arity_bits; l++) {
            const int ab = d.reduction_arity_bits[l];
            lg -= ab;
            byte_order.push_back(seg((size_t)2 << ab, 0));
            byte_order.push_back(seg(1, 1));
            byte_order.push_back(seg(4 * (size_t)(lg - d.cap_height), 0));
        }
    }
    size_t fin = (size_t)1 << d.degree_bits;
    for (int l = 0; l < d.num_reduction_arity_bits; l++) fin >>= d.reduction_arity_bits[l];
    byte_order.push_back(seg(2 * fin, 0));
    byte_order.push_back(seg(1, 0));                                   // pow_witness
    byte_order.push_back(seg((size_t)d.num_public_inputs, 0));        // preceded by its length (usize = u64 LE)
}
static bool proof_desc_ok(const p2g_circuit_desc* d) {
    return d && d->degree_bits >= 0 && d->degree_bits <= 30 && d->rate_bits >= 0 && d->rate_bits <= 8 && d->cap_height >= 0 &&
           d->cap_height <= d->degree_bits + d->rate_bits && d->num_reduction_arity_bits >= 0 && d->num_reduction_arity_bits <= 16 &&
           d->num_query_rounds >= 0 && d->num_public_inputs >= 0 && d->num_challenges >= 1 && d->quotient_degree_factor >= 2;
}
extern "C" size_t p2g_proof_bytes_len(const p2g_circuit_desc* d) {
    if (!proof_desc_ok(d)) return 0;
    std::vector<ProofSeg> segs; std::vector<int> order;
    proof_segments(*d, segs, order);
    size_t len = 8;                                                    // public-input count
    for (const auto& s : segs) len += s.kind == 1 ? 1 : 8 * s.words;
    return len;
}
extern "C" int32_t p2g_proof_to_bytes(const p2g_circuit_desc* d, const uint64_t* words, size_t nwords, uint8_t* out, size_t cap,
                                      size_t* len_out) {
    if (!proof_desc_ok(d) || !words || !out) return P2G_E_BADARG;
    std::vector<ProofSeg> segs; std::vector<int> order;
    proof_segments(*d, segs, order);
    if (nwords != segs.back().word_pos + segs.back().words || cap < p2g_proof_bytes_len(d)) return P2G_E_BADARG;
    uint8_t* o = out;
    auto put64 = [&](uint64_t v) { for (int b = 0; b < 8; b++) *o++ = (uint8_t)(v >> (8 * b)); };
    for (size_t k = 0; k < order.size(); k++) {
        const ProofSeg& s = segs[order[k]];
        if (k + 1 == order.size()) put64((uint64_t)d->num_public_inputs);
        if (s.kind == 1) {
            if (words[s.word_pos] > 255) return P2G_E_BADARG;
            *o++ = (uint8_t)words[s.word_pos];
        } else {
            for (size_t i = 0; i < s.words; i++) put64(words[s.word_pos + i]);
        }
    }
    if (len_out) *len_out = (size_t)(o - out);
    return P2G_OK;
}
extern "C" int32_t p2g_proof_from_bytes(const p2g_circuit_desc* d, const uint8_t* bytes, size_t len, uint64_t* words_out,
                                        size_t cap_words, size_t* words_len) {
    if (!proof_desc_ok(d) || !bytes || !words_out) return P2G_E_BADARG;
    std::vector<ProofSeg> segs; std::vector<int> order;
    proof_segments(*d, segs, order);
    const size_t nwords = segs.back().word_pos + segs.back().words;
    if (len != p2g_proof_bytes_len(d) || cap_words < nwords) return P2G_E_BADARG;
    const uint8_t* p = bytes;
    auto get64 = [&]() { uint64_t v = 0; for (int b = 0; b < 8; b++) v |= (uint64_t)*p++ << (8 * b); return v; };
    for (size_t k = 0; k < order.size(); k++) {
        const ProofSeg& s = segs[order[k]];
        if (k + 1 == order.size() && get64() != (uint64_t)d->num_public_inputs) return P2G_E_BADARG;
        if (s.kind == 1) {
            words_out[s.word_pos] = *p++;
            // the path length is implied by the circuit: anything else is a malformed proof
            const size_t expect = segs[order[k + 1]].words / 4;
            if (words_out[s.word_pos] != expect) return P2G_E_BADARG;
        } else {
            for (size_t i = 0; i < s.words; i++) {
                const uint64_t v = get64();
                if (v >= GL_P) return P2G_E_BADARG;                    // upstream: non-canonical field element
                words_out[s.word_pos + i] = v;
            }
        }
    }
    if (words_len) *words_len = nwords;
    return P2G_OK;
}

extern "C" int32_t p2g_circuit_load(p2g_ctx* ctx, const p2g_circuit_desc* desc, p2g_circuit** out, uint64_t* cap_out) {
    if (!ctx || !desc || !out) return P2G_E_BADARG;
    const p2g_circuit_desc& d = *desc;
    if (d.num_challenges < 1 || d.num_challenges > MAX_CH || d.num_routed_wires > MAX_ROUTED || d.num_luts > 8 ||
        d.quotient_degree_factor != (1 << d.rate_bits) || d.rate_bits != 3 || d.num_public_inputs < 0 ||
        d.pow_bits < 1 || d.pow_bits > 32 || d.cap_height < 0 || d.num_query_rounds < 1 || d.num_query_rounds > 64 ||
        d.degree_bits < 2 || d.degree_bits > P2G_MAX_LOG_N || d.num_partial_products + 1 > 16 ||
        d.num_routed_wires / 2 > 8 * (d.quotient_degree_factor - 1) || d.quotient_degree_factor - 1 > 8) {
        ctx->err = "unsupported circuit configuration"; return P2G_E_BADARG;
    }
    for (int g = 0; g < d.num_gates; g++)
        if (d.gates[g].kind < P2G_GATE_NOOP || d.gates[g].kind > P2G_GATE_POSEIDON) { ctx->err = "unknown gate kind"; return P2G_E_BADARG; }
    CU(cudaSetDevice(ctx->device));
    p2g_circuit* C = new p2g_circuit();
    CircuitGuard guard{ctx, C, nullptr};
    C->d = d;
    C->gates.assign(d.gates, d.gates + d.num_gates);
    C->lut_lens.assign(d.lut_lens, d.lut_lens + d.num_luts);
    size_t lut_total = 0; for (int i = 0; i < d.num_luts; i++) lut_total += d.lut_lens[i];
    C->lut_data.assign(d.lut_data, d.lut_data + 2 * lut_total);
    C->lookup_rows.assign(d.lookup_rows, d.lookup_rows + 3 * d.num_luts);
    C->k_is.assign(d.k_is, d.k_is + d.num_routed_wires);
    C->d.gates = C->gates.data(); C->d.lut_lens = C->lut_lens.data(); C->d.lut_data = C->lut_data.data();
    C->d.lookup_rows = C->lookup_rows.data(); C->d.k_is = C->k_is.data(); C->d.constants_sigmas = nullptr;
    CircuitDev& cd = C->cd;
    cd.logn = d.degree_bits; cd.rate_bits = d.rate_bits; cd.nch = d.num_challenges; cd.R = d.num_routed_wires; cd.W = d.num_wires;
    cd.num_sel = d.num_selectors; cd.num_lsel = d.num_lookup_selectors; cd.num_consts = d.num_constants;
    cd.NC = cd.num_sel + cd.num_lsel + cd.num_consts;
    cd.num_prods = d.num_partial_products; cd.qdf = d.quotient_degree_factor;
    cd.nlp = num_lookup_polys(d); cd.num_sldc = cd.nlp ? cd.nlp - 1 : 0;
    cd.lu_slots = cd.R / 2; cd.lut_slots = cd.R / 3; cd.lu_degree = cd.qdf - 1;
    cd.lut_degree = cd.num_sldc ? (cd.lut_slots + cd.num_sldc - 1) / cd.num_sldc : 0;
    cd.num_luts = d.num_luts; cd.num_gates = d.num_gates; cd.num_gate_constraints = d.num_gate_constraints;
    cd.zs_cols = cd.nch * (1 + cd.num_prods + cd.nlp);
    if (cd.lut_degree > 8) { ctx->err = "lut_degree > 8"; return P2G_E_BADARG; }
    int nterms = cd.nch + cd.nch * (cd.num_prods + 1) + (cd.num_luts ? cd.nch * (4 + cd.num_luts + 2 * cd.num_sldc) : 0) + cd.num_gate_constraints;
    if (nterms > 256) { ctx->err = "too many vanishing terms"; return P2G_E_BADARG; }
    C->proof_words = proof_len(d);
    const size_t n = (size_t)1 << cd.logn, N = n << cd.rate_bits;
    const int logN = cd.logn + cd.rate_bits;
    int rc;
    // preprocessed commitment
    const int ncs = cd.NC + cd.R;
    gl_t* d_vals;
    if ((rc = ctx_alloc(ctx, &d_vals, (size_t)ncs * n))) return rc;
    guard.tmp = d_vals;
    CU(cudaMemcpyAsync(d_vals, d.constants_sigmas, (size_t)ncs * n * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st));
    if ((rc = commit_dev(ctx, d_vals, ncs, cd.logn, cd.rate_bits, d.cap_height, true, &C->cs, true))) return rc;
    if (cap_out) memcpy(cap_out, C->cs->cap_host.data(), C->cs->cap_host.size() * sizeof(gl_t));
    if ((rc = ctx_alloc(ctx, &C->d_sigmas, (size_t)cd.R * n))) return rc;
    CU(cudaMemcpyAsync(C->d_sigmas, d_vals + (size_t)cd.NC * n, (size_t)cd.R * n * sizeof(gl_t), cudaMemcpyDeviceToDevice, ctx->st));
    ctx_free(ctx, d_vals); guard.tmp = nullptr;
    // domain tables
    if ((rc = ctx_alloc(ctx, &C->d_subgroup, n))) return rc;
    if ((rc = ctx_alloc(ctx, &C->d_domain, N))) return rc;
    P2G_COUNT_LAUNCH(1); domain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->st>>>(cd.logn, gl_root_of_unity(cd.logn), 1, C->d_subgroup, 0);
    P2G_COUNT_LAUNCH(1); domain_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->st>>>(logN, gl_root_of_unity(logN), 7, C->d_domain, 1);
    if ((rc = ctx_alloc(ctx, &C->d_l0inv, N))) return rc;
    P2G_COUNT_LAUNCH(1); l0inv_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->st>>>(N, (gl_t)n, C->d_domain, C->d_l0inv);
    CU(cudaGetLastError());
    // quotient coefficient recovery tables
    {
        std::vector<gl_t> tab(8 * n), small(16);
        gl_t wN = gl_root_of_unity(logN), inv8 = gl_inv(8);
        for (int s = 0; s < 8; s++) {
            gl_t hinv = gl_inv(gl_mul(7, gl_pow(wN, s)));
            gl_t x = inv8;
            for (size_t j = 0; j < n; j++) { tab[(size_t)s * n + j] = x; x = gl_mul(x, hinv); }
        }
        gl_t w8inv = gl_inv(gl_root_of_unity(3)), sninv = gl_inv(gl_pow(7, n));
        small[0] = 1; small[8] = 1;
        for (int k = 1; k < 8; k++) { small[k] = gl_mul(small[k - 1], w8inv); small[8 + k] = gl_mul(small[8 + k - 1], sninv); }
        if ((rc = ctx_alloc(ctx, &C->d_qtable, 8 * n))) return rc;
        if ((rc = ctx_alloc(ctx, &C->d_small, 16))) return rc;
        CU(cudaMemcpyAsync(C->d_qtable, tab.data(), tab.size() * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(C->d_small, small.data(), small.size() * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st));
        CU(ctx_wait(ctx));
    }
    // gates and row kinds
    {
        std::vector<uint8_t> kind(n, 0);
        for (int l = 0; l < d.num_luts; l++) {
            int last_lu = d.lookup_rows[3 * l], last_lut = d.lookup_rows[3 * l + 1], first_lut = d.lookup_rows[3 * l + 2];
            if (last_lu < 0 || first_lut + 1 >= (int)n || last_lu > last_lut || last_lut > first_lut) { ctx->err = "bad lookup rows"; return P2G_E_BADARG; }
            for (int r = last_lu; r < last_lut; r++) kind[r] = 1;
            for (int r = last_lut; r <= first_lut; r++) kind[r] = 2;
        }
        CU(cudaMallocFromPoolAsync((void**)&C->d_row_kind, n, ctx->pool, ctx->st));
        CU(cudaMallocFromPoolAsync((void**)&C->d_gates, sizeof(p2g_gate) * std::max(1, d.num_gates), ctx->pool, ctx->st));
        CU(cudaMemcpyAsync(C->d_row_kind, kind.data(), n, cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(C->d_gates, C->gates.data(), sizeof(p2g_gate) * d.num_gates, cudaMemcpyHostToDevice, ctx->st));
        std::vector<int> off(8, 0), len(8, 0);
        { int t = 0; for (int l = 0; l < d.num_luts; l++) { off[l] = t; len[l] = d.lut_lens[l]; t += d.lut_lens[l]; } }
        CU(cudaMallocFromPoolAsync((void**)&C->d_lut_data, std::max<size_t>(4, 4 * lut_total), ctx->pool, ctx->st));
        CU(cudaMallocFromPoolAsync((void**)&C->d_lut_off, 8 * sizeof(int), ctx->pool, ctx->st));
        CU(cudaMallocFromPoolAsync((void**)&C->d_lut_len, 8 * sizeof(int), ctx->pool, ctx->st));
        CU(cudaMemcpyAsync(C->d_lut_data, C->lut_data.data(), 4 * lut_total, cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(C->d_lut_off, off.data(), 8 * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(C->d_lut_len, len.data(), 8 * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
        CU(ctx_wait(ctx));
    }
    size_t fin = n; for (int l = 0; l < d.num_reduction_arity_bits; l++) fin >>= d.reduction_arity_bits[l];
    C->final_len = (int)fin;
    guard.c = nullptr;
    *out = C;
    return P2G_OK;
}
extern "C" int32_t p2g_circuit_free(p2g_ctx* ctx, p2g_circuit* C) {
    if (!ctx || !C) return P2G_E_BADARG;
    cudaSetDevice(ctx->device);
    if (C->cs) p2g_batch_free(ctx, C->cs);
    ctx_free(ctx, C->d_sigmas); ctx_free(ctx, C->d_subgroup); ctx_free(ctx, C->d_domain); ctx_free(ctx, C->d_l0inv); ctx_free(ctx, C->d_qtable);
    ctx_free(ctx, C->d_small); ctx_free(ctx, C->d_row_kind); ctx_free(ctx, C->d_gates);
    ctx_free(ctx, C->d_lut_data); ctx_free(ctx, C->d_lut_off); ctx_free(ctx, C->d_lut_len);
    delete C;
    return P2G_OK;
}
extern "C" size_t p2g_proof_words(const p2g_circuit* C) { return C ? C->proof_words : 0; }

// ---------------------------------------------------------------------------------------------
struct StageTimer {
    p2g_ctx* ctx; cudaEvent_t ev[16]; int n; bool on;
    StageTimer(p2g_ctx* c) : ctx(c), n(0), on(c->timing) { if (on) for (auto& e : ev) cudaEventCreate(&e); mark(); }
    void mark() { if (on && n < 16) cudaEventRecord(ev[n++], ctx->st); }
    void finish(p2g_timings* t) {
        if (!on) return;
        cudaEventSynchronize(ev[n - 1]);
        float* f = &t->h2d;
        for (int i = 0; i + 1 < n && i < 11; i++) cudaEventElapsedTime(&f[i], ev[i], ev[i + 1]);
        cudaEventElapsedTime(&t->total, ev[0], ev[n - 1]);
        for (auto& e : ev) cudaEventDestroy(e);
    }
};

struct Pow2Table { ext_t p[32]; };
__global__ void __launch_bounds__(256)
ext_powers_kernel2(Pow2Table tb, size_t count, gl_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    ext_t r = ext_make(1, 0);
    size_t e = j;
    for (int b = 0; e; b++, e >>= 1) if (e & 1) r = ext_mul(r, tb.p[b]);
    out[2 * j] = r.c0; out[2 * j + 1] = r.c1;
}

static ext_t ext_reduce_with_powers(const ext_t* v, size_t n, ext_t alpha) {
    ext_t acc = ext_make(0, 0);
    for (size_t i = n; i-- > 0;) acc = ext_add(ext_mul(acc, alpha), v[i]);
    return acc;
}

// Owner of the device temporaries of one prove() call.  Buffers are released as soon as the stage that
// needs them is over (S.free), and the destructor releases the rest, so the early returns of
// prove_impl (CUDA errors, P2G_E_UNSAT for a witness that does not satisfy the circuit, ...) do not
// leak pool memory.
struct Scratch {
    p2g_ctx* ctx;
    std::vector<void*> ptrs;
    std::vector<p2g_batch*> batches;
    explicit Scratch(p2g_ctx* c) : ctx(c) {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    int alloc(gl_t** p, size_t words) {
        int rc = ctx_alloc(ctx, p, words);
        if (rc == P2G_OK) ptrs.push_back(*p);
        return rc;
    }
    cudaError_t alloc_bytes(void** p, size_t bytes) {
        cudaError_t e = cudaMallocFromPoolAsync(p, bytes ? bytes : 1, ctx->pool, ctx->st);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    void free(const void* p) {
        if (!p) return;
        for (size_t i = 0; i < ptrs.size(); i++)
            if (ptrs[i] == p) { ptrs.erase(ptrs.begin() + (long)i); cudaFreeAsync((void*)p, ctx->st); return; }
    }
    void own(p2g_batch* b) { if (b) batches.push_back(b); }
    void free_batch(p2g_batch* b) {
        for (size_t i = 0; i < batches.size(); i++)
            if (batches[i] == b) { batches.erase(batches.begin() + (long)i); p2g_batch_free(ctx, b); return; }
    }
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, ctx->st);
        for (p2g_batch* b : batches) p2g_batch_free(ctx, b);
    }
};

// Challenge-dependent constants of one proof.  deltas_flat: [nch][4] = (a, b, alpha, delta) per challenge (only with
// lookups), alphas may be NULL (filled later by set_alphas).
static void fill_consts(ProofConsts* pc, const p2g_circuit* C, const gl_t* betas, const gl_t* gammas, const gl_t* deltas_flat,
                        const gl_t pi_hash[4]) {
    const CircuitDev& cd = C->cd;
    const size_t n = (size_t)1 << cd.logn;
    memset(pc, 0, sizeof(ProofConsts));
    for (int i = 0; i < cd.nch; i++) { pc->betas[i] = betas[i]; pc->gammas[i] = gammas[i]; }
    if (cd.num_luts > 0)
        for (int i = 0; i < cd.nch; i++) for (int j = 0; j < 4; j++) pc->deltas[i][j] = deltas_flat[4 * i + j];
    for (int j = 0; j < cd.R; j++) {
        pc->k_is[j] = C->k_is[j];
        for (int i = 0; i < cd.nch; i++) pc->beta_kis[i][j] = gl_mul(pc->betas[i], C->k_is[j]);
    }
    memcpy(pc->pi_hash, pi_hash, 4 * sizeof(gl_t));
    gl_t sn = gl_pow(7, n), w8 = gl_root_of_unity(cd.rate_bits), t = 1;
    for (int s = 0; s < (1 << cd.rate_bits); s++) { pc->zh[s] = gl_sub(gl_mul(sn, t), 1); pc->zh_inv[s] = gl_inv(pc->zh[s]); t = gl_mul(t, w8); }
    if (cd.num_luts > 0) for (int i = 0; i < cd.nch; i++) pc->delta_pow_slots[i] = gl_pow(pc->deltas[i][3], cd.lut_slots);
}
static void set_alphas(ProofConsts* pc, int nch, const gl_t* alphas) {
    for (int i = 0; i < nch; i++) {
        pc->alphas[i] = alphas[i];
        gl_t p = 1;
        for (int k = 0; k < 256; k++) { pc->alpha_pows[i][k] = p; p = gl_mul(p, alphas[i]); }
    }
}
// compute_quotient_polys on (a shard of) the LDE domain: quotient kernel, per-coset inverse NTTs, optional
// all-gather of the interpolants, cross-coset combination -> the nch * 8 chunk coefficient columns d_qc [nch*8][n]
struct ShardHost;
static int quotient_chunks(p2g_ctx* ctx, const p2g_circuit* C, ShardDev shd, const ProofConsts* d_pc, const gl_t* d_lut_evals,
                           const p2g_batch* wb, const p2g_batch* zb, gl_t* d_qv, gl_t* d_qa, gl_t* d_qc,
                           const std::function<int(const gl_t*, const gl_t**)>& gather_interpolants) {
    const CircuitDev& cd = C->cd;
    const int logn = cd.logn, nch = cd.nch;
    const size_t n = (size_t)1 << logn, N_loc = (size_t)shd.blk_count << logn;
    cudaStream_t st = ctx->st;
    int rc;
    bool has_pos = false;
    for (const auto& g : C->gates) has_pos |= g.kind == P2G_GATE_POSEIDON;
    P2G_COUNT_LAUNCH(1);
    if (has_pos) quotient_kernel<true><<<(unsigned)((N_loc + 127) / 128), 128, 0, st>>>(cd, shd, d_pc, d_lut_evals, C->d_gates, C->cs->lde, wb->lde, zb->lde, C->d_domain, C->d_l0inv, d_qv);
    else quotient_kernel<false><<<(unsigned)((N_loc + 127) / 128), 128, 0, st>>>(cd, shd, d_pc, d_lut_evals, C->d_gates, C->cs->lde, wb->lde, zb->lde, C->d_domain, C->d_l0inv, d_qv);
    CU(cudaGetLastError());
    const NttPlan* inv;
    if ((rc = ctx_get_plan(ctx, NTT_KIND_INV, logn, 0, &inv))) return rc;
    // (large n: the outer stages run in place over d_qv, which is not read again)
    if (ntt_launch(inv, d_qv, n, d_qa, n, nch * (int)shd.blk_count, 1, st, 0, 0, d_qv)) { ctx->err = "quotient intt"; return P2G_E_CUDA; }
    // the 8-point cross-coset combination needs every coset's interpolant: all-gather of 16 N / world bytes when sharded
    const gl_t* d_qall = d_qa;
    if (gather_interpolants && (rc = gather_interpolants(d_qa, &d_qall))) return rc;
    dim3 grid((unsigned)((n + 255) / 256), nch);
    P2G_COUNT_LAUNCH(1); quotient_combine_kernel<<<grid, 256, 0, st>>>(logn, nch, shd.blk_count, d_qall, C->d_qtable, C->d_small, C->d_small + 8, d_qc);
    CU(cudaGetLastError());
    return P2G_OK;
}

// Coset shard of ONE proof over `world` GPUs (one process per GPU): this rank extends, hashes and evaluates only the
// leaf blocks [blk_first, blk_first + blk_count); what the ranks need from each other -- Merkle cap entries, the
// per-coset quotient interpolants, the last FRI layer, the query records -- goes through `exchange`, an all-gather
// of `bytes` bytes per rank from `send` into `recv` ([world][bytes], rank order) that the caller implements (NCCL).
struct ShardHost {
    uint32_t blk_first, blk_count, world, rank;
    gl_t* send; gl_t* recv; size_t buf_bytes;
    p2g_exchange_fn exchange; void* user;
};

// p2g_fri_prove: prove_impl entered at prove_openings with the caller's batches, evaluation point and transcript
struct FriResume {
    const p2g_batch *wires, *zs, *quotient;
    ext_t zeta;
    uint64_t* challenger_io;      // [12 state][8 input buffer][8 output buffer][input length][output length]
};
static size_t fri_part_words(const p2g_circuit* C) {
    // the flat proof minus the three caps, the openings and the public inputs
    const p2g_circuit_desc& d = C->d;
    const int nlp = num_lookup_polys(d), nch = d.num_challenges;
    const int NC = d.num_selectors + d.num_lookup_selectors + d.num_constants;
    const size_t open = 2 * ((size_t)NC + d.num_routed_wires + d.num_wires + 2 * nch + (size_t)nch * d.num_partial_products +
                             (size_t)nch * d.quotient_degree_factor + 2 * (size_t)nch * nlp);
    return C->proof_words - 3 * ((size_t)4 << d.cap_height) - open - (size_t)d.num_public_inputs;
}
static int32_t prove_impl(p2g_ctx* ctx, const p2g_circuit* C, const gl_t* d_wires_in, bool wires_on_host,
                          const uint64_t* public_inputs, uint64_t* proof_out, size_t proof_cap, size_t* proof_words_out,
                          const ShardHost* sh = nullptr, const FriResume* fr = nullptr) {
    if (!ctx || !C || (!fr && !d_wires_in) || !proof_out) return P2G_E_BADARG;
    if (proof_cap < (fr ? fri_part_words(C) : C->proof_words)) return P2G_E_BADARG;
    if (!fr && C->d.num_public_inputs > 0 && !public_inputs) { ctx->err = "public_inputs is NULL"; return P2G_E_BADARG; }
    CU(cudaSetDevice(ctx->device));
    const p2g_circuit_desc& d = C->d;
    const CircuitDev& cd = C->cd;
    const int nch = cd.nch, R = cd.R, W = cd.W, NC = cd.NC, qdf = cd.qdf, num_prods = cd.num_prods, nlp = cd.nlp;
    const int logn = cd.logn, logN = logn + cd.rate_bits;
    const size_t n = (size_t)1 << logn, N = (size_t)1 << logN;
    const int zs_cols = cd.zs_cols, zpp = nch * (1 + num_prods), nlz = zs_cols - zpp;
    const size_t capw = (size_t)4 << d.cap_height;
    const bool has_lookup = cd.num_luts > 0;
    cudaStream_t st = ctx->st;
    int rc;
    Scratch S(ctx);     // every temporary of this proof; whatever is still held when the function returns
                        // (an error path, e.g. an unsatisfied witness) is released by its destructor
    auto S_own = [&](p2g_batch* b) { S.own(b); };
    // ---- coset shard geometry (one shard = everything when sh is null) ----
    const uint32_t nblk = 1u << cd.rate_bits;
    const uint32_t b0 = sh ? sh->blk_first : 0, bc = sh ? sh->blk_count : nblk;
    uint32_t blk_log = 0; while ((1u << blk_log) < bc) blk_log++;
    const size_t N_loc = (size_t)bc << logn, j_off = (size_t)b0 << logn;
    const ShardDev shd = {b0, bc};
    const int cap_h_loc = sh ? d.cap_height + (int)blk_log - cd.rate_bits : d.cap_height;   // cap entries below a shard's subtree roots
    if (sh) {
        if ((1u << blk_log) != bc || b0 % bc || b0 + bc > nblk || sh->world * bc != nblk || sh->rank * bc != b0 || !sh->exchange ||
            !sh->send || !sh->recv || cap_h_loc < 0) { ctx->err = "bad coset shard"; return P2G_E_BADARG; }
    }
    // all-gather of `bytes` device bytes per rank; afterwards sh->recv holds [world][bytes]
    auto gather = [&](int stage, const void* d_src, size_t bytes) -> int {
        if (bytes > sh->buf_bytes) { ctx->err = "exchange buffer too small"; return P2G_E_BADARG; }
        CU(cudaMemcpyAsync(sh->send, d_src, bytes, cudaMemcpyDeviceToDevice, st));
        CU(ctx_wait(ctx));
        if (sh->exchange(sh->user, stage, (uint64_t)bytes) != 0) { ctx->err = "exchange callback failed"; return P2G_E_CUDA; }
        return P2G_OK;
    };
    // PolynomialBatch commitment of this shard's blocks; cap_full receives the whole cap (gathered when sharded)
    auto commit = [&](int stage, const gl_t* cols, uint32_t ncols, bool from_values, p2g_batch** out, std::vector<gl_t>& cap_full) -> int {
        int r = commit_dev(ctx, cols, ncols, logn, cd.rate_bits, (uint32_t)cap_h_loc, from_values, out, !sh, sh ? b0 : 0, sh ? bc : 0);
        if (r) return r;
        S_own(*out);
        cap_full.resize(capw);
        if (!sh) { cap_full = (*out)->cap_host; return P2G_OK; }
        const size_t part = (size_t)4 << cap_h_loc;
        if ((r = gather(stage, (*out)->cap, part * sizeof(gl_t)))) return r;
        CU(cudaMemcpyAsync(ctx->pinned, sh->recv, capw * sizeof(gl_t), cudaMemcpyDeviceToHost, st));
        CU(ctx_wait(ctx));
        memcpy(cap_full.data(), ctx->pinned, capw * sizeof(gl_t));
        return P2G_OK;
    };
    StageTimer tm(ctx);
    const bool dbg = getenv("P2G_DEBUG") != nullptr;
#define DBG(msg) do { if (dbg) { cudaError_t e_ = cudaStreamSynchronize(st); fprintf(stderr, "[p2g] %s (%s)\n", msg, cudaGetErrorString(e_)); } } while (0)

    gl_t* d_wires = nullptr;
    p2g_batch *wb = nullptr, *zb = nullptr, *qb = nullptr;
    std::vector<gl_t> wcap, zcap, qcap;
    Challenger ch;
    ProofConsts* pc_host = (ProofConsts*)(ctx->pinned + ctx->pinned_words / 2);   // upper half of the pinned staging buffer (the lower half receives D2H results)
    static_assert(sizeof(ProofConsts) < 16384, "ProofConsts too large");
    gl_t deltas_flat[16] = {0};
    ProofConsts* d_pc = nullptr;
    gl_t* d_lut_evals = nullptr;
    ext_t zeta;
    if (fr) {
        // p2g_fri_prove: the commitments, the transcript up to the openings and zeta are the caller's
        wb = const_cast<p2g_batch*>(fr->wires); zb = const_cast<p2g_batch*>(fr->zs); qb = const_cast<p2g_batch*>(fr->quotient);
        memcpy(ch.state, fr->challenger_io, 12 * sizeof(gl_t));
        memcpy(ch.in_buf, fr->challenger_io + 12, 8 * sizeof(gl_t));
        memcpy(ch.out_buf, fr->challenger_io + 20, 8 * sizeof(gl_t));
        ch.in_len = (int)fr->challenger_io[28]; ch.out_len = (int)fr->challenger_io[29];
        zeta = fr->zeta;
    } else {
    gl_t pi_hash[4];
    host_hash_no_pad(public_inputs, (size_t)d.num_public_inputs, pi_hash);

    // ---- stage A: wires ----
    const gl_t* wires = d_wires_in;
    if (wires_on_host) {
        if ((rc = S.alloc(&d_wires, (size_t)W * n))) return rc;
        CU(cudaMemcpyAsync(d_wires, d_wires_in, (size_t)W * n * sizeof(gl_t), cudaMemcpyHostToDevice, st));
        wires = d_wires;
    }
    tm.mark();
    if ((rc = commit(1, wires, W, true, &wb, wcap))) return rc;
    tm.mark();

    DBG("wires committed");
    ch.observe_many(d.circuit_digest, 4);
    ch.observe_many(pi_hash, 4);
    ch.observe_many(wcap.data(), capw);
    {
        gl_t betas[MAX_CH], gammas[MAX_CH];
        for (int i = 0; i < nch; i++) betas[i] = ch.get();
        for (int i = 0; i < nch; i++) gammas[i] = ch.get();
        if (has_lookup) {      // get_n_challenges(4 * nch) with the first 2 * nch taken from (betas, gammas)
            int k = 0;
            for (int i = 0; i < nch; i++) deltas_flat[k++] = betas[i];
            for (int i = 0; i < nch; i++) deltas_flat[k++] = gammas[i];
            for (int i = 0; i < 2 * nch; i++) deltas_flat[k++] = ch.get();
        }
        fill_consts(pc_host, C, betas, gammas, deltas_flat, pi_hash);
    }
    CU(S.alloc_bytes((void**)&d_pc, sizeof(ProofConsts)));
    CU(cudaMemcpyAsync(d_pc, pc_host, sizeof(ProofConsts), cudaMemcpyHostToDevice, st));
    if ((rc = S.alloc(&d_lut_evals, MAX_CH * 8))) return rc;
    if (has_lookup) {
        P2G_COUNT_LAUNCH(1); lut_eval_kernel<<<dim3(cd.num_luts, nch), 1024, 0, st>>>(cd, d_pc, C->d_lut_data, C->d_lut_off, C->d_lut_len, d_lut_evals);
    }

    // ---- Z, partial products, lookup polys ----
    gl_t *d_zs, *d_rowprod;
    if ((rc = S.alloc(&d_zs, (size_t)zs_cols * n))) return rc;
    if ((rc = S.alloc(&d_rowprod, (size_t)nch * n))) return rc;
    {
        dim3 grid((unsigned)((n + 127) / 128), nch);
        P2G_COUNT_LAUNCH(1); zs_chunk_kernel<<<grid, 128, 0, st>>>(cd, d_pc, wires, C->d_sigmas, C->d_subgroup, d_zs, d_rowprod);
        P2G_COUNT_LAUNCH(1); zs_scan_kernel<<<nch, 1024, 0, st>>>(cd, d_rowprod);
        P2G_COUNT_LAUNCH(1); zs_apply_kernel<<<grid, 128, 0, st>>>(cd, d_rowprod, d_zs);
        if (has_lookup) {
            P2G_COUNT_LAUNCH(1); lookup_rows_kernel<<<grid, 128, 0, st>>>(cd, d_pc, wires, C->d_row_kind, d_zs);
            P2G_COUNT_LAUNCH(1); lookup_scan_kernel<<<nch, 1024, 0, st>>>(cd, d_pc, C->d_row_kind, d_zs);
        }
        CU(cudaGetLastError());
    }
    DBG("zs built");
    if (ctx->keep_debug) {
        ctx->last_zs.resize((size_t)zs_cols * n);
        CU(cudaMemcpyAsync(ctx->last_zs.data(), d_zs, ctx->last_zs.size() * sizeof(gl_t), cudaMemcpyDeviceToHost, st));
    }
    tm.mark();
    if ((rc = commit(2, d_zs, zs_cols, true, &zb, zcap))) return rc;
    S.free(d_zs); S.free(d_rowprod);
    tm.mark();
    ch.observe_many(zcap.data(), capw);
    // pc_host (pinned) was consumed by the H2D copy above once commit_dev synchronised
    {
        gl_t alphas[MAX_CH];
        for (int i = 0; i < nch; i++) alphas[i] = ch.get();
        set_alphas(pc_host, nch, alphas);
    }
    CU(cudaMemcpyAsync(d_pc, pc_host, sizeof(ProofConsts), cudaMemcpyHostToDevice, st));

    DBG("zs committed");
    // ---- quotient ----
    gl_t *d_qv, *d_qa, *d_qc;
    if ((rc = S.alloc(&d_qv, (size_t)nch * N_loc))) return rc;
    if ((rc = S.alloc(&d_qa, (size_t)nch * N_loc))) return rc;
    if ((rc = S.alloc(&d_qc, (size_t)nch * N))) return rc;
    {
        std::function<int(const gl_t*, const gl_t**)> gq;
        if (sh) gq = [&](const gl_t* own, const gl_t** all) -> int {
            int r = gather(3, own, (size_t)nch * N_loc * sizeof(gl_t));
            *all = sh->recv;
            return r;
        };
        if ((rc = quotient_chunks(ctx, C, shd, d_pc, d_lut_evals, wb, zb, d_qv, d_qa, d_qc, gq))) return rc;
    }
    DBG("quotient evaluated");
    if (ctx->keep_debug) {
        ctx->last_quotient_chunks.resize((size_t)nch * N);
        CU(cudaMemcpyAsync(ctx->last_quotient_chunks.data(), d_qc, (size_t)nch * N * sizeof(gl_t), cudaMemcpyDeviceToHost, st));
    }
    tm.mark();
    if ((rc = commit(4, d_qc, nch * qdf, false, &qb, qcap))) return rc;
    S.free(d_qv); S.free(d_qa); S.free(d_qc);
    tm.mark();
    ch.observe_many(qcap.data(), capw);
    zeta = ch.get_ext();
    DBG("quotient committed");
    }   // !fr
    const gl_t g = gl_root_of_unity(logn);
    const ext_t zeta_next = ext_mul_base(zeta, g);

    // ---- openings ----
    const p2g_batch* oracles[4] = {C->cs, wb, zb, qb};
    const int tot0 = NC + R + W + zpp + nch * qdf + nlz, tot1 = nch + nlz;
    std::vector<const gl_t*> plist(tot0 + tot1);
    {
        int k = 0;
        for (int i = 0; i < NC + R; i++) plist[k++] = oracles[0]->coeffs + (size_t)i * n;
        for (int i = 0; i < W; i++) plist[k++] = oracles[1]->coeffs + (size_t)i * n;
        for (int i = 0; i < zpp; i++) plist[k++] = oracles[2]->coeffs + (size_t)i * n;
        for (int i = 0; i < nch * qdf; i++) plist[k++] = oracles[3]->coeffs + (size_t)i * n;
        for (int i = zpp; i < zs_cols; i++) plist[k++] = oracles[2]->coeffs + (size_t)i * n;
        for (int i = 0; i < nch; i++) plist[k++] = oracles[2]->coeffs + (size_t)i * n;
        for (int i = zpp; i < zs_cols; i++) plist[k++] = oracles[2]->coeffs + (size_t)i * n;
    }
    const gl_t** d_plist; gl_t *d_zp, *d_open;
    CU(S.alloc_bytes((void**)&d_plist, plist.size() * sizeof(gl_t*)));
    CU(cudaMemcpyAsync(d_plist, plist.data(), plist.size() * sizeof(gl_t*), cudaMemcpyHostToDevice, st));
    if ((rc = S.alloc(&d_zp, 4 * n))) return rc;
    if ((rc = S.alloc(&d_open, 2 * (size_t)(tot0 + tot1)))) return rc;
    {
        Pow2Table t0, t1;
        t0.p[0] = zeta; t1.p[0] = zeta_next;
        for (int b = 1; b < 32; b++) { t0.p[b] = ext_mul(t0.p[b - 1], t0.p[b - 1]); t1.p[b] = ext_mul(t1.p[b - 1], t1.p[b - 1]); }
        P2G_COUNT_LAUNCH(1); ext_powers_kernel2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t0, n, d_zp);
        P2G_COUNT_LAUNCH(1); ext_powers_kernel2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t1, n, d_zp + 2 * n);
        P2G_COUNT_LAUNCH(1); eval_polys_kernel<<<tot0, 256, 0, st>>>(d_plist, d_zp, n, d_open);
        P2G_COUNT_LAUNCH(1); eval_polys_kernel<<<tot1, 256, 0, st>>>(d_plist + tot0, d_zp + 2 * n, n, d_open + 2 * (size_t)tot0);
        CU(cudaGetLastError());
    }
    std::vector<ext_t> open((size_t)tot0 + tot1);
    CU(cudaMemcpyAsync(ctx->pinned, d_open, open.size() * sizeof(ext_t), cudaMemcpyDeviceToHost, st));
    CU(ctx_wait(ctx));
    memcpy(open.data(), ctx->pinned, open.size() * sizeof(ext_t));
    tm.mark();
    // ---- proof assembly starts: caps + openings (p2g_fri_prove: the caller has them and has observed them) ----
    gl_t* w = proof_out;
    if (!fr) {
    ch.observe_many((const gl_t*)open.data(), 2 * open.size());
    memcpy(w, wcap.data(), capw * 8); w += capw;
    memcpy(w, zcap.data(), capw * 8); w += capw;
    memcpy(w, qcap.data(), capw * 8); w += capw;
    {
        const ext_t* o0 = open.data(); const ext_t* o1 = open.data() + tot0;
        auto put = [&](const ext_t* p, int cnt) { memcpy(w, p, (size_t)cnt * sizeof(ext_t)); w += 2 * (size_t)cnt; };
        int off = 0;
        put(o0 + off, NC); off += NC;            // constants
        put(o0 + off, R); off += R;              // plonk_sigmas
        put(o0 + off, W); off += W;              // wires
        put(o0 + off, nch);                      // plonk_zs
        put(o1, nch);                            // plonk_zs_next
        put(o0 + off + nch, nch * num_prods); off += zpp;   // partial_products
        put(o0 + off, nch * qdf); off += nch * qdf;         // quotient_polys
        put(o0 + off, nlz);                      // lookup_zs
        put(o1 + nch, nlz);                      // lookup_zs_next
    }
    }   // !fr

    DBG("openings done");
    // ---- prove_openings: batch combination ----
    const ext_t fri_alpha = ch.get_ext();
    gl_t *d_comp, *d_comp_lde, *d_vals;
    if ((rc = S.alloc(&d_comp, 4 * n))) return rc;
    if ((rc = S.alloc(&d_comp_lde, 4 * N_loc))) return rc;
    if ((rc = S.alloc(&d_vals, 2 * N_loc))) return rc;
    gl_t* d_apow;
    {
        // alpha^j for j < max(|batch 0|, |batch 1|)
        const size_t na_ = (size_t)(tot0 > tot1 ? tot0 : tot1);
        if ((rc = S.alloc(&d_apow, 2 * na_))) return rc;
        Pow2Table ta;
        ta.p[0] = fri_alpha;
        for (int b = 1; b < 32; b++) ta.p[b] = ext_mul(ta.p[b - 1], ta.p[b - 1]);
        const size_t na = (size_t)(tot0 > tot1 ? tot0 : tot1);
        P2G_COUNT_LAUNCH(1); ext_powers_kernel2<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(ta, na, d_apow);
    }
    P2G_COUNT_LAUNCH(1); fri_compose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_plist, tot0, d_apow, n, d_comp, d_comp + n);
    P2G_COUNT_LAUNCH(1); fri_compose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_plist + tot0, tot1, d_apow, n, d_comp + 2 * n, d_comp + 3 * n);
    CU(cudaGetLastError());
    {
        const NttPlan* lde;
        if ((rc = ctx_get_plan(ctx, NTT_KIND_LDE, logn, cd.rate_bits, &lde))) return rc;
        if (ntt_launch(lde, d_comp, n, d_comp_lde, N_loc, 4, 0, st, sh ? b0 : 0, sh ? bc : 0)) { ctx->err = "fri lde"; return P2G_E_CUDA; }
        ext_t comp0_at = ext_reduce_with_powers(open.data(), tot0, fri_alpha);
        ext_t comp1_at = ext_reduce_with_powers(open.data() + tot0, tot1, fri_alpha);
        ext_t shift0 = ext_pow(fri_alpha, (uint64_t)tot1);
        P2G_COUNT_LAUNCH(1); fri_final_values_kernel<<<(unsigned)((N_loc + 255) / 256), 256, 0, st>>>(d_comp_lde, N_loc, C->d_domain + j_off, zeta, zeta_next, comp0_at, comp1_at, shift0, d_vals);
        CU(cudaGetLastError());
    }
    tm.mark();

    DBG("fri combined");
    // ---- fri_committed_trees ----
    const int nl = d.num_reduction_arity_bits;
    struct Layer { gl_t* vals; gl_t* digests; gl_t* cap; int log_len, arity_bits; };
    std::vector<Layer> layers(nl);
    gl_t* cur_vals = d_vals; int cur_log = logN;
    gl_t shift = 7;
    ext_t fri_betas[16];
    for (int l = 0; l < nl; l++) {
        const int ab = d.reduction_arity_bits[l];
        if (ab > 4) { ctx->err = "arity > 16"; return P2G_E_BADARG; }
        Layer& L = layers[l];
        L.vals = cur_vals; L.log_len = cur_log; L.arity_bits = ab;
        const uint32_t log_leaves = cur_log - ab;                     // of the whole layer
        // a coset shard holds the leaves of its blocks: leaf index (bit-reversed order) = block | position, so a leaf of
        // 2^ab consecutive values, its subtree and its cap entries never straddle two shards
        if (sh && (log_leaves < (uint32_t)cd.rate_bits || (int)(log_leaves - cd.rate_bits + blk_log) < cap_h_loc)) {
            ctx->err = "FRI layer too small for this cap height / shard count"; return P2G_E_BADARG;
        }
        const uint32_t log_leaves_loc = sh ? log_leaves - cd.rate_bits + blk_log : log_leaves;
        const size_t part = (size_t)4 << cap_h_loc;
        if ((rc = S.alloc(&L.digests, merkle_digest_words(log_leaves_loc, (uint32_t)cap_h_loc)))) return rc;
        if ((rc = S.alloc(&L.cap, part))) return rc;
        if (merkle_build(cur_vals, 0, 0, 2u << ab, log_leaves_loc, (uint32_t)cap_h_loc, L.digests, L.cap, st)) { ctx->err = "fri merkle"; return P2G_E_CUDA; }
        const gl_t* d_cap_full = L.cap;
        if (sh) { if ((rc = gather(5 + l, L.cap, part * sizeof(gl_t)))) return rc; d_cap_full = sh->recv; }
        CU(cudaMemcpyAsync(ctx->pinned, d_cap_full, capw * sizeof(gl_t), cudaMemcpyDeviceToHost, st));
        CU(ctx_wait(ctx));
        memcpy(w, ctx->pinned, capw * sizeof(gl_t));
        ch.observe_many(w, capw);
        w += capw;
        const ext_t beta = ch.get_ext();
        fri_betas[l] = beta;
        gl_t* nxt;
        const size_t chunks = (size_t)1 << log_leaves_loc, chunk_first = sh ? ((size_t)b0 << log_leaves) >> cd.rate_bits : 0;
        if ((rc = S.alloc(&nxt, 2 * chunks))) return rc;
        P2G_COUNT_LAUNCH(1); fri_fold_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, st>>>(cur_vals, cur_log, ab, gl_inv(shift), gl_inv(gl_root_of_unity(cur_log)),
                                                                           gl_inv(gl_root_of_unity(ab)), gl_inv((gl_t)1 << ab), beta, nxt, chunk_first, chunks);
        CU(cudaGetLastError());
        cur_vals = nxt; cur_log = (int)log_leaves;
        shift = gl_pow(shift, (uint64_t)1 << ab);
    }
    // final polynomial: interpolate the last layer on the host (<= a few hundred points)
    const size_t fl = (size_t)1 << cur_log;
    std::vector<ext_t> fvals(fl), fcoef(fl);
    if (sh && fl < nblk) { ctx->err = "final FRI layer smaller than the number of cosets"; return P2G_E_BADARG; }
    const gl_t* d_final = cur_vals;
    if (sh) { if ((rc = gather(21, cur_vals, (fl >> cd.rate_bits << blk_log) * sizeof(ext_t)))) return rc; d_final = sh->recv; }
    CU(cudaMemcpyAsync(ctx->pinned, d_final, fl * sizeof(ext_t), cudaMemcpyDeviceToHost, st));
    CU(ctx_wait(ctx));
    memcpy(fvals.data(), ctx->pinned, fl * sizeof(ext_t));
    {
        // natural-order values, then coset_ifft(shift): c_i = shift^-i / len * sum_m v[m] w^-(i m)
        std::vector<ext_t> nat(fl);
        for (size_t j = 0; j < fl; j++) nat[gl_bitrev((uint32_t)j, cur_log)] = fvals[j];
        gl_t winv = gl_inv(gl_root_of_unity(cur_log)), linv = gl_inv((gl_t)fl), sinv = gl_inv(shift);
        std::vector<gl_t> wp(fl);
        wp[0] = 1; for (size_t i = 1; i < fl; i++) wp[i] = gl_mul(wp[i - 1], winv);
        gl_t sc = linv;
        for (size_t i = 0; i < fl; i++) {
            ext_t acc = ext_make(0, 0);
            for (size_t m = 0; m < fl; m++) acc = ext_add(acc, ext_mul_base(nat[m], wp[(i * m) & (fl - 1)]));
            fcoef[i] = ext_mul_base(acc, sc);
            sc = gl_mul(sc, sinv);
        }
    }
    const size_t final_len = fl >> cd.rate_bits;
    for (size_t i = final_len; i < fl; i++)
        if (fcoef[i].c0 || fcoef[i].c1) { ctx->err = "FRI final polynomial has non-zero high coefficients"; return P2G_E_UNSAT; }
    ch.observe_many((const gl_t*)fcoef.data(), 2 * final_len);
    tm.mark();

    DBG("fri committed");
    // ---- fri_proof_of_work: lowest nonce ----
    gl_t pow_witness = 0;
    {
        PowState ps; memcpy(ps.s, ch.state, sizeof(ps.s));
        const int pos = ch.in_len;
        for (int i = 0; i < pos; i++) ps.s[i] = ch.in_buf[i];
        unsigned long long* d_best;
        CU(S.alloc_bytes((void**)&d_best, 8));
        // Windows of candidates in increasing order keep the "lowest nonce" semantics.  The first window
        // has 2^pow_bits candidates (a hit with probability 1 - 1/e), then the windows double up to
        // 2^(pow_bits+2): about 2 x 2^pow_bits permutations on average instead of a fixed 4 x -- with
        // several proofs in flight the grind is throughput, not latency.
        unsigned long long win = 1ull << (d.pow_bits < 12 ? 12 : d.pow_bits);
        const unsigned long long win_max = win << 2;
        bool found = false;
        int windows = 0;     // bounded: 1024 windows without a hit cannot happen with a working kernel (e^-4000)
        for (unsigned long long base = 0; !found && windows < 1024; base += win, win = win < win_max ? win * 2 : win, windows++) {
            CU(cudaMemsetAsync(d_best, 0xFF, 8, st));
            P2G_COUNT_LAUNCH(1); pow_grind_kernel<<<(unsigned)(win / 256), 256, 0, st>>>(ps, pos, d.pow_bits, base, d_best);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(ctx->pinned, d_best, 8, cudaMemcpyDeviceToHost, st));
            CU(ctx_wait(ctx));
            unsigned long long best = *(unsigned long long*)ctx->pinned;
            if (best != ~0ull) { pow_witness = best; found = true; }
        }
        S.free(d_best);
        if (!found) { ctx->err = "proof of work failed"; return P2G_E_POW; }
        ch.observe(pow_witness);
        gl_t resp = ch.get();
        if ((resp >> (64 - d.pow_bits)) != 0) { ctx->err = "proof of work response mismatch"; return P2G_E_POW; }
    }
    tm.mark();

    DBG("pow done");
    // ---- query rounds ----
    const int nq = d.num_query_rounds;
    std::vector<unsigned long long> qidx(nq), qown(nq);
    for (int q = 0; q < nq; q++) {
        qidx[q] = ch.get() % N;
        const unsigned long long blk = qidx[q] >> logn;               // the coset block that holds this query's leaves
        qown[q] = (blk >= b0 && blk < b0 + bc) ? qidx[q] : ~0ull;     // another shard's query: skipped by the gather kernel
    }
    std::vector<GatherTree> gt;
    unsigned long long rec = 0;
    const int ocols[4] = {NC + R, W, zs_cols, nch * qdf};
    for (int o = 0; o < 4; o++) {
        // the preprocessed batch (o = 0) is whole on every rank; the per-proof batches hold this shard's blocks
        const bool whole = o == 0 || !sh;
        GatherTree T; T.data = oracles[o]->lde; T.digests = oracles[o]->digests; T.col_stride = whole ? N : N_loc;
        T.leaf_len = ocols[o]; T.log_leaves = whole ? logN : logn + (int)blk_log; T.path_len = logN - d.cap_height; T.index_shift = 0; T.out_offset = rec;
        T.leaf_first = whole ? 0 : j_off;
        rec += T.leaf_len + 1 + 4ull * T.path_len;
        gt.push_back(T);
    }
    {
        unsigned shift_bits = 0;
        for (int l = 0; l < nl; l++) {
            shift_bits += layers[l].arity_bits;
            const unsigned log_leaves = layers[l].log_len - layers[l].arity_bits;
            GatherTree T; T.data = layers[l].vals; T.digests = layers[l].digests; T.col_stride = 0;
            T.leaf_len = 2u << layers[l].arity_bits; T.log_leaves = sh ? log_leaves - cd.rate_bits + blk_log : log_leaves;
            T.path_len = log_leaves - d.cap_height; T.index_shift = shift_bits; T.out_offset = rec;
            T.leaf_first = sh ? ((unsigned long long)b0 << log_leaves) >> cd.rate_bits : 0;
            rec += T.leaf_len + 1 + 4ull * T.path_len;
            gt.push_back(T);
        }
    }
    GatherTree* d_gt; unsigned long long* d_qidx; gl_t* d_q;
    CU(S.alloc_bytes((void**)&d_gt, gt.size() * sizeof(GatherTree)));
    CU(S.alloc_bytes((void**)&d_qidx, nq * 8));
    if ((rc = S.alloc(&d_q, rec * nq))) return rc;
    CU(cudaMemcpyAsync(d_gt, gt.data(), gt.size() * sizeof(GatherTree), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_qidx, qown.data(), nq * 8, cudaMemcpyHostToDevice, st));
    if (sh) CU(cudaMemsetAsync(d_q, 0, rec * nq * sizeof(gl_t), st));
    P2G_COUNT_LAUNCH(1); query_gather_kernel<<<dim3(nq, (unsigned)gt.size()), 128, 0, st>>>(d_gt, (int)gt.size(), d_qidx, rec, d_q);
    CU(cudaGetLastError());
    if (!sh) {
        CU(cudaMemcpyAsync(w, d_q, rec * nq * sizeof(gl_t), cudaMemcpyDeviceToHost, st));
        CU(ctx_wait(ctx));
    } else {
        // every rank gathered the records of the queries that fall into its blocks; after the all-gather each
        // record is taken from its owner
        if ((rc = gather(22, d_q, rec * nq * sizeof(gl_t)))) return rc;
        for (int q = 0; q < nq; q++) {
            const size_t owner = (size_t)((qidx[q] >> logn) / bc);
            CU(cudaMemcpyAsync(w + (size_t)q * rec, sh->recv + (owner * nq + q) * rec, rec * sizeof(gl_t), cudaMemcpyDeviceToHost, st));
        }
        CU(ctx_wait(ctx));
    }
    w += rec * nq;
    memcpy(w, fcoef.data(), final_len * sizeof(ext_t)); w += 2 * final_len;
    *w++ = pow_witness;
    if (!fr) for (int i = 0; i < d.num_public_inputs; i++) *w++ = public_inputs[i];
    tm.mark();
    tm.finish(&ctx->timings);

    // transcript for parity tests
    if (!fr) {
        p2g_transcript& tr = ctx->transcript; memset(&tr, 0, sizeof(tr));
        for (int i = 0; i < nch; i++) { tr.betas[i] = pc_host->betas[i]; tr.gammas[i] = pc_host->gammas[i]; tr.alphas[i] = pc_host->alphas[i]; }
        memcpy(tr.deltas, deltas_flat, sizeof(gl_t) * 4 * nch);
        tr.zeta[0] = zeta.c0; tr.zeta[1] = zeta.c1; tr.fri_alpha[0] = fri_alpha.c0; tr.fri_alpha[1] = fri_alpha.c1;
        for (int l = 0; l < nl; l++) { tr.fri_betas[2 * l] = fri_betas[l].c0; tr.fri_betas[2 * l + 1] = fri_betas[l].c1; }
        tr.pow_witness = pow_witness;
        for (int q = 0; q < nq && q < 64; q++) tr.query_indices[q] = qidx[q];
    }
    // release
    for (int l = 0; l < nl; l++) { S.free(layers[l].digests); S.free(layers[l].cap); if (l > 0) S.free(layers[l].vals); }
    if (nl > 0) S.free(cur_vals);
    S.free(d_vals); S.free(d_comp); S.free(d_comp_lde); S.free(d_zp); S.free(d_apow); S.free(d_open);
    S.free(d_lut_evals);
    S.free(d_plist); S.free(d_gt); S.free(d_qidx); S.free(d_q); S.free(d_pc);
    if (!fr) { S.free_batch(wb); S.free_batch(zb); S.free_batch(qb); }
    if (d_wires) S.free(d_wires);
    if (fr) {
        memcpy(fr->challenger_io, ch.state, 12 * sizeof(gl_t));
        memcpy(fr->challenger_io + 12, ch.in_buf, 8 * sizeof(gl_t));
        memcpy(fr->challenger_io + 20, ch.out_buf, 8 * sizeof(gl_t));
        fr->challenger_io[28] = (uint64_t)ch.in_len; fr->challenger_io[29] = (uint64_t)ch.out_len;
    }
    if ((size_t)(w - proof_out) != (fr ? fri_part_words(C) : C->proof_words)) { ctx->err = "proof length mismatch"; return P2G_E_BADARG; }
    if (proof_words_out) *proof_words_out = (size_t)(w - proof_out);
    return P2G_OK;
}

extern "C" int32_t p2g_prove(p2g_ctx* ctx, const p2g_circuit* c, const uint64_t* wires_host, const uint64_t* public_inputs,
                             uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out) {
    return prove_impl(ctx, c, wires_host, true, public_inputs, proof_out, proof_cap_words, proof_words_out);
}
extern "C" int32_t p2g_prove_sharded(p2g_ctx* ctx, const p2g_circuit* c, const uint64_t* wires_host, const uint64_t* public_inputs,
                                     uint32_t rank, uint32_t world, uint64_t* send_dev, uint64_t* recv_dev, size_t buf_bytes,
                                     p2g_exchange_fn exchange, void* user, uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out) {
    if (!c || !world || rank >= world) return P2G_E_BADARG;
    const uint32_t nblk = 1u << c->cd.rate_bits;
    if (world > nblk || nblk % world) { if (ctx) ctx->err = "world size must divide the number of cosets"; return P2G_E_BADARG; }
    ShardHost sh = {rank * (nblk / world), nblk / world, world, rank, send_dev, recv_dev, buf_bytes, exchange, user};
    return prove_impl(ctx, c, wires_host, true, public_inputs, proof_out, proof_cap_words, proof_words_out, &sh);
}
extern "C" size_t p2g_shard_buffer_bytes(const p2g_circuit* c, uint32_t world) {
    if (!c || !world) return 0;
    const p2g_circuit_desc& d = c->d;
    const size_t N = (size_t)1 << (d.degree_bits + d.rate_bits);
    size_t q = (size_t)d.num_challenges * (N / world) * sizeof(gl_t);                     // quotient interpolants
    size_t per_q = 0;                                                                     // one query record
    {
        const int nlp = num_lookup_polys(d), NC = d.num_selectors + d.num_lookup_selectors + d.num_constants;
        const int cols[4] = {NC + d.num_routed_wires, d.num_wires, d.num_challenges * (1 + d.num_partial_products + nlp), d.num_challenges * d.quotient_degree_factor};
        const int logN = d.degree_bits + d.rate_bits;
        for (int o = 0; o < 4; o++) per_q += cols[o] + 1 + 4 * (size_t)(logN - d.cap_height);
        int lg = logN;
        for (int l = 0; l < d.num_reduction_arity_bits; l++) { lg -= d.reduction_arity_bits[l]; per_q += (2u << d.reduction_arity_bits[l]) + 1 + 4 * (size_t)(lg - d.cap_height); }
    }
    size_t r = per_q * d.num_query_rounds * sizeof(gl_t);
    size_t b = q > r ? q : r;
    return (b + 255) & ~(size_t)255;
}
// ---- staged entry points: the quotient stage and the openings on their own ------------
extern "C" int32_t p2g_quotient(p2g_ctx* ctx, const p2g_circuit* C, const p2g_batch* wires, const p2g_batch* zs,
                                const uint64_t* public_inputs, const uint64_t* betas, const uint64_t* gammas, const uint64_t* deltas,
                                const uint64_t* alphas, p2g_batch** quotient_out, uint64_t* cap_out) {
    if (!ctx || !C || !wires || !zs || !betas || !gammas || !alphas || !quotient_out) return P2G_E_BADARG;
    const CircuitDev& cd = C->cd;
    const p2g_circuit_desc& d = C->d;
    if ((cd.num_luts > 0 && !deltas) || (d.num_public_inputs > 0 && !public_inputs)) return P2G_E_BADARG;
    if (wires->ncols != (uint32_t)cd.W || zs->ncols != (uint32_t)cd.zs_cols || wires->log_n != (uint32_t)cd.logn || zs->log_n != (uint32_t)cd.logn ||
        wires->blk_log != (uint32_t)cd.rate_bits || zs->blk_log != (uint32_t)cd.rate_bits || wires->rate_bits != (uint32_t)cd.rate_bits) {
        ctx->err = "batches do not match the circuit (columns, degree, whole LDE domain)"; return P2G_E_BADARG;
    }
    CU(cudaSetDevice(ctx->device));
    const size_t N = (size_t)1 << (cd.logn + cd.rate_bits);
    Scratch S(ctx);
    int rc;
    gl_t pi_hash[4];
    host_hash_no_pad(public_inputs, (size_t)d.num_public_inputs, pi_hash);
    ProofConsts* pc_host = (ProofConsts*)(ctx->pinned + ctx->pinned_words / 2);
    fill_consts(pc_host, C, betas, gammas, deltas, pi_hash);
    set_alphas(pc_host, cd.nch, alphas);
    ProofConsts* d_pc; gl_t *d_lut_evals, *d_qv, *d_qa, *d_qc;
    CU(S.alloc_bytes((void**)&d_pc, sizeof(ProofConsts)));
    CU(cudaMemcpyAsync(d_pc, pc_host, sizeof(ProofConsts), cudaMemcpyHostToDevice, ctx->st));
    if ((rc = S.alloc(&d_lut_evals, MAX_CH * 8))) return rc;
    if (cd.num_luts > 0) {
        P2G_COUNT_LAUNCH(1); lut_eval_kernel<<<dim3(cd.num_luts, cd.nch), 1024, 0, ctx->st>>>(cd, d_pc, C->d_lut_data, C->d_lut_off, C->d_lut_len, d_lut_evals);
    }
    if ((rc = S.alloc(&d_qv, (size_t)cd.nch * N))) return rc;
    if ((rc = S.alloc(&d_qa, (size_t)cd.nch * N))) return rc;
    if ((rc = S.alloc(&d_qc, (size_t)cd.nch * N))) return rc;
    const ShardDev whole = {0u, 1u << cd.rate_bits};
    if ((rc = quotient_chunks(ctx, C, whole, d_pc, d_lut_evals, wires, zs, d_qv, d_qa, d_qc, nullptr))) return rc;
    if ((rc = commit_dev(ctx, d_qc, cd.nch * cd.qdf, cd.logn, cd.rate_bits, d.cap_height, false, quotient_out, true))) return rc;
    if (cap_out) memcpy(cap_out, (*quotient_out)->cap_host.data(), (*quotient_out)->cap_host.size() * sizeof(gl_t));
    return P2G_OK;
}
// OpeningSet::new for any set of batches of one degree: f(zeta) for every committed polynomial, batch by batch
extern "C" int32_t p2g_open(p2g_ctx* ctx, const p2g_batch* const* batches, uint32_t n_batches, const uint64_t zeta[2], uint64_t* openings_out) {
    if (!ctx || !batches || !n_batches || !zeta || !openings_out || zeta[0] >= GL_P || zeta[1] >= GL_P) return P2G_E_BADARG;
    std::vector<const gl_t*> plist;
    for (uint32_t b = 0; b < n_batches; b++) {
        if (!batches[b] || batches[b]->log_n != batches[0]->log_n) { ctx->err = "batches of different degree"; return P2G_E_BADARG; }
        for (uint32_t c = 0; c < batches[b]->ncols; c++) plist.push_back(batches[b]->coeffs + (size_t)c * batches[b]->n());
    }
    CU(cudaSetDevice(ctx->device));
    const size_t n = batches[0]->n(), tot = plist.size();
    Scratch S(ctx);
    const gl_t** d_plist; gl_t *d_zp, *d_open; int rc;
    CU(S.alloc_bytes((void**)&d_plist, tot * sizeof(gl_t*)));
    CU(cudaMemcpyAsync(d_plist, plist.data(), tot * sizeof(gl_t*), cudaMemcpyHostToDevice, ctx->st));
    if ((rc = S.alloc(&d_zp, 2 * n))) return rc;
    if ((rc = S.alloc(&d_open, 2 * tot))) return rc;
    Pow2Table t0;
    t0.p[0] = ext_make(zeta[0], zeta[1]);
    for (int b = 1; b < 32; b++) t0.p[b] = ext_mul(t0.p[b - 1], t0.p[b - 1]);
    P2G_COUNT_LAUNCH(1); ext_powers_kernel2<<<(unsigned)((n + 255) / 256), 256, 0, ctx->st>>>(t0, n, d_zp);
    P2G_COUNT_LAUNCH(1); eval_polys_kernel<<<(unsigned)tot, 256, 0, ctx->st>>>(d_plist, d_zp, n, d_open);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(openings_out, d_open, 2 * tot * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(ctx_wait(ctx));               // the host vector of pointers dies with this frame
    return P2G_OK;
}

// PolynomialBatch::prove_openings + fri_proof on their own: batch combination at zeta and g zeta, LDE, commit phase,
// proof of work (lowest nonce) and query rounds, with the caller's commitments and transcript
extern "C" size_t p2g_fri_proof_words(const p2g_circuit* C) { return C ? fri_part_words(C) : 0; }
extern "C" int32_t p2g_fri_prove(p2g_ctx* ctx, const p2g_circuit* C, const p2g_batch* wires, const p2g_batch* zs, const p2g_batch* quotient,
                                 const uint64_t zeta[2], uint64_t* challenger_io, uint64_t* fri_out, size_t fri_cap_words,
                                 size_t* fri_words_out) {
    if (!ctx || !C || !wires || !zs || !quotient || !zeta || !challenger_io || !fri_out || zeta[0] >= GL_P || zeta[1] >= GL_P) return P2G_E_BADARG;
    const CircuitDev& cd = C->cd;
    const p2g_batch* bs[3] = {wires, zs, quotient};
    const uint32_t cols[3] = {(uint32_t)cd.W, (uint32_t)cd.zs_cols, (uint32_t)(cd.nch * cd.qdf)};
    for (int i = 0; i < 3; i++)
        if (bs[i]->ncols != cols[i] || bs[i]->log_n != (uint32_t)cd.logn || bs[i]->rate_bits != (uint32_t)cd.rate_bits ||
            bs[i]->blk_log != (uint32_t)cd.rate_bits || bs[i]->cap_height != (uint32_t)C->d.cap_height) {
            ctx->err = "batches do not match the circuit (columns, degree, cap height, whole LDE domain)"; return P2G_E_BADARG;
        }
    if (challenger_io[28] > 7 || challenger_io[29] > 8) { ctx->err = "bad challenger buffer lengths"; return P2G_E_BADARG; }
    for (int i = 0; i < 28; i++) if (challenger_io[i] >= GL_P) { ctx->err = "non-canonical challenger word"; return P2G_E_BADARG; }
    FriResume fr = {wires, zs, quotient, ext_make(zeta[0], zeta[1]), challenger_io};
    return prove_impl(ctx, C, nullptr, false, nullptr, fri_out, fri_cap_words, fri_words_out, nullptr, &fr);
}

// p2g_prove for a batch: proof i runs on context i mod n_ctx, one host thread per context, so the latency-bound
// parts of one proof (Fiat-Shamir round trips, tree tops) overlap the heavy kernels of the others
extern "C" int32_t p2g_prove_batch(p2g_ctx* const* ctxs, const p2g_circuit* const* circuits, uint32_t n_ctx,
                                   const uint64_t* const* wires_host, const uint64_t* const* public_inputs, uint32_t n_proofs,
                                   uint64_t* const* proofs_out, size_t proof_cap_words, int32_t* status_out) {
    if (!ctxs || !circuits || !n_ctx || !wires_host || !proofs_out || (n_proofs && !status_out)) return P2G_E_BADARG;
    for (uint32_t t = 0; t < n_ctx; t++) if (!ctxs[t] || !circuits[t]) return P2G_E_BADARG;
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_ctx && t < n_proofs; t++)
        th.emplace_back([=]() {
            for (uint32_t i = t; i < n_proofs; i += n_ctx) {
                size_t got = 0;
                status_out[i] = prove_impl(ctxs[t], circuits[t], wires_host[i], true, public_inputs ? public_inputs[i] : nullptr,
                                           proofs_out[i], proof_cap_words, &got);
            }
        });
    for (auto& x : th) x.join();
    for (uint32_t i = 0; i < n_proofs; i++) if (status_out[i] != P2G_OK) return status_out[i];
    return P2G_OK;
}
extern "C" int32_t p2g_prove_dev(p2g_ctx* ctx, const p2g_circuit* c, const uint64_t* wires_dev, const uint64_t* public_inputs,
                                 uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out) {
    return prove_impl(ctx, c, wires_dev, false, public_inputs, proof_out, proof_cap_words, proof_words_out);
}
// ---- device-side PartitionWitness::full_witness (iop/witness.rs) --------------------------------
struct p2g_wmap {
    int32_t* d_map;      // [W][n] slot index or -1
    int64_t* d_fixed_pos; gl_t* d_fixed_val;
    uint32_t num_slots, num_fixed;
    size_t cells;
};
__global__ void __launch_bounds__(256)
wire_gather_kernel(const int32_t* __restrict__ map, const gl_t* __restrict__ slots, size_t cells, gl_t* __restrict__ wires) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const int32_t m = __ldg(map + i);
    wires[i] = m >= 0 ? __ldg(slots + m) : 0;
}
__global__ void __launch_bounds__(256)
wire_fixed_kernel(const int64_t* __restrict__ pos, const gl_t* __restrict__ val, uint32_t count, gl_t* __restrict__ wires) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) wires[pos[i]] = val[i];
}
extern "C" int32_t p2g_wmap_load(p2g_ctx* ctx, const p2g_circuit* c, const int32_t* wire_map, uint32_t num_slots,
                                 const int64_t* fixed_pos, const uint64_t* fixed_val, uint32_t num_fixed, p2g_wmap** out) {
    if (!ctx || !c || !wire_map || !out || !num_slots || (num_fixed && (!fixed_pos || !fixed_val))) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    const size_t cells = (size_t)c->cd.W << c->cd.logn;
    for (size_t i = 0; i < cells; i++) if (wire_map[i] >= (int32_t)num_slots || wire_map[i] < -1) { ctx->err = "wire map entry out of range"; return P2G_E_BADARG; }
    for (uint32_t i = 0; i < num_fixed; i++) if (fixed_pos[i] < 0 || (size_t)fixed_pos[i] >= cells || fixed_val[i] >= GL_P) { ctx->err = "fixed cell out of range"; return P2G_E_BADARG; }
    p2g_wmap* m = new p2g_wmap();
    m->num_slots = num_slots; m->num_fixed = num_fixed; m->cells = cells;
    m->d_map = nullptr; m->d_fixed_pos = nullptr; m->d_fixed_val = nullptr;
    bool ok = cudaMalloc(&m->d_map, cells * sizeof(int32_t)) == cudaSuccess;
    if (ok && num_fixed) ok = cudaMalloc(&m->d_fixed_pos, num_fixed * sizeof(int64_t)) == cudaSuccess && cudaMalloc(&m->d_fixed_val, num_fixed * sizeof(gl_t)) == cudaSuccess;
    if (ok) ok = cudaMemcpyAsync(m->d_map, wire_map, cells * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->st) == cudaSuccess;
    if (ok && num_fixed) ok = cudaMemcpyAsync(m->d_fixed_pos, fixed_pos, num_fixed * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st) == cudaSuccess &&
                              cudaMemcpyAsync(m->d_fixed_val, fixed_val, num_fixed * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st) == cudaSuccess;
    if (ok) ok = ctx_wait(ctx) == cudaSuccess;
    if (!ok) { cudaFree(m->d_map); cudaFree(m->d_fixed_pos); cudaFree(m->d_fixed_val); delete m; ctx->err = "wmap upload"; return P2G_E_CUDA; }
    *out = m;
    return P2G_OK;
}
extern "C" int32_t p2g_wmap_free(p2g_ctx* ctx, p2g_wmap* m) {
    if (!ctx || !m) return P2G_E_BADARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    cudaFree(m->d_map); cudaFree(m->d_fixed_pos); cudaFree(m->d_fixed_val);
    delete m;
    return P2G_OK;
}
// slots (device) -> wire matrix (device, from the context's pool)
static int wmap_gather_dev(p2g_ctx* ctx, const p2g_wmap* m, const gl_t* d_slots, gl_t** d_wires_out) {
    int rc; gl_t* d_wires;
    if ((rc = ctx_alloc(ctx, &d_wires, m->cells))) return rc;
    P2G_COUNT_LAUNCH(1); wire_gather_kernel<<<(unsigned)((m->cells + 255) / 256), 256, 0, ctx->st>>>(m->d_map, d_slots, m->cells, d_wires);
    if (m->num_fixed) { P2G_COUNT_LAUNCH(1); wire_fixed_kernel<<<(m->num_fixed + 255) / 256, 256, 0, ctx->st>>>(m->d_fixed_pos, m->d_fixed_val, m->num_fixed, d_wires); }
    if (cudaGetLastError() != cudaSuccess) { ctx_free(ctx, d_wires); ctx->err = "wire gather launch"; return P2G_E_CUDA; }
    *d_wires_out = d_wires;
    return P2G_OK;
}
// slots (host) -> wire matrix (device)
static int wmap_gather(p2g_ctx* ctx, const p2g_wmap* m, const uint64_t* slots_host, gl_t** d_wires_out) {
    int rc; gl_t* d_slots;
    if ((rc = ctx_alloc(ctx, &d_slots, m->num_slots))) return rc;
    if (cudaMemcpyAsync(d_slots, slots_host, (size_t)m->num_slots * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st) != cudaSuccess) {
        ctx_free(ctx, d_slots); ctx->err = "slot upload"; return P2G_E_CUDA;
    }
    rc = wmap_gather_dev(ctx, m, d_slots, d_wires_out);
    ctx_free(ctx, d_slots);                 // stream-ordered: released after the gather
    return rc;
}
// ---- witness generation on the device in front of the prover (witgen.cu) ---------------------------
int wprog_launch(p2g_ctx* ctx, const p2g_wprog* p, const gl_t* d_in, uint32_t count, gl_t* d_ext, int32_t* d_err);
uint32_t wprog_ext_total(const p2g_wprog* p);
uint32_t wprog_num_inputs(const p2g_wprog* p);
extern "C" int32_t p2g_prove_inputs(p2g_ctx* ctx, const p2g_circuit* c, const p2g_wmap* m, const p2g_wprog* prog,
                                    const uint64_t* input_vals_host, const uint64_t* public_inputs, uint64_t* proof_out,
                                    size_t proof_cap_words, size_t* proof_words_out) {
    if (!ctx || !c || !m || !prog || !input_vals_host) return P2G_E_BADARG;
    if (m->cells != ((size_t)c->cd.W << c->cd.logn) || m->num_slots != wprog_ext_total(prog)) {
        ctx->err = "wire map / witness program belong to another circuit"; return P2G_E_BADARG;
    }
    CU(cudaSetDevice(ctx->device));
    const uint32_t ni = wprog_num_inputs(prog);
    Scratch S(ctx);
    gl_t *d_in, *d_ext, *d_wires = nullptr; int32_t* d_err; int rc;
    if ((rc = S.alloc(&d_in, ni))) return rc;
    if ((rc = S.alloc(&d_ext, m->num_slots))) return rc;
    CU(S.alloc_bytes((void**)&d_err, sizeof(int32_t)));
    CU(cudaMemcpyAsync(d_in, input_vals_host, (size_t)ni * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st));
    if ((rc = wprog_launch(ctx, prog, d_in, 1, d_ext, d_err))) return rc;
    if ((rc = wmap_gather_dev(ctx, m, d_ext, &d_wires))) return rc;
    S.ptrs.push_back(d_wires);
    S.free(d_in); S.free(d_ext);
    rc = prove_impl(ctx, c, d_wires, false, public_inputs, proof_out, proof_cap_words, proof_words_out);
    if (rc) return rc;
    // generator errors (a looked-up value outside its table, a preset partition that disagrees with the
    // computed value) are flagged by the kernel; the proof of such a witness is discarded
    int32_t* flag = (int32_t*)ctx->pinned;
    CU(cudaMemcpyAsync(flag, d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(ctx_wait(ctx));
    if (*flag) {
        ctx->err = (*flag & 2) ? "lookup input not in table (P2W_E_LOOKUP)"
                 : (*flag & 4) ? "partition set twice with different values (P2W_E_CONFLICT)" : "non-canonical input value";
        return (*flag & 2) ? -11 : (*flag & 4) ? -10 : P2G_E_BADARG;
    }
    return P2G_OK;
}
extern "C" int32_t p2g_prove_slots(p2g_ctx* ctx, const p2g_circuit* c, const p2g_wmap* m, const uint64_t* slots_host,
                                   const uint64_t* public_inputs, uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out) {
    if (!ctx || !c || !m || !slots_host) return P2G_E_BADARG;
    if (m->cells != ((size_t)c->cd.W << c->cd.logn)) { ctx->err = "wire map belongs to another circuit"; return P2G_E_BADARG; }
    CU(cudaSetDevice(ctx->device));
    gl_t* d_wires; int rc;
    if ((rc = wmap_gather(ctx, m, slots_host, &d_wires))) return rc;
    rc = prove_impl(ctx, c, d_wires, false, public_inputs, proof_out, proof_cap_words, proof_words_out);
    ctx_free(ctx, d_wires);
    return rc;
}
extern "C" int32_t p2g_prove_slots_dev(p2g_ctx* ctx, const p2g_circuit* c, const p2g_wmap* m, const uint64_t* slots_dev,
                                       const uint64_t* public_inputs, uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out) {
    if (!ctx || !c || !m || !slots_dev) return P2G_E_BADARG;
    if (m->cells != ((size_t)c->cd.W << c->cd.logn)) { ctx->err = "wire map belongs to another circuit"; return P2G_E_BADARG; }
    CU(cudaSetDevice(ctx->device));
    gl_t* d_wires; int rc;
    if ((rc = wmap_gather_dev(ctx, m, slots_dev, &d_wires))) return rc;
    rc = prove_impl(ctx, c, d_wires, false, public_inputs, proof_out, proof_cap_words, proof_words_out);
    ctx_free(ctx, d_wires);
    return rc;
}
extern "C" int32_t p2g_wmap_fill(p2g_ctx* ctx, const p2g_circuit* c, const p2g_wmap* m, const uint64_t* slots_host, uint64_t* wires_out_host) {
    if (!ctx || !c || !m || !slots_host || !wires_out_host) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    gl_t* d_wires; int rc;
    if ((rc = wmap_gather(ctx, m, slots_host, &d_wires))) return rc;
    CU(cudaMemcpyAsync(wires_out_host, d_wires, m->cells * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(ctx_wait(ctx));
    ctx_free(ctx, d_wires);
    return P2G_OK;
}
extern "C" int32_t p2g_last_transcript(p2g_ctx* ctx, p2g_transcript* out) {
    if (!ctx || !out) return P2G_E_BADARG;
    *out = ctx->transcript; return P2G_OK;
}
extern "C" int32_t p2g_last_zs_values(p2g_ctx* ctx, uint64_t* out) {
    if (!ctx || !out || ctx->last_zs.empty()) return P2G_E_BADARG;
    memcpy(out, ctx->last_zs.data(), ctx->last_zs.size() * sizeof(gl_t)); return P2G_OK;
}
extern "C" int32_t p2g_last_quotient_chunks(p2g_ctx* ctx, uint64_t* out) {
    if (!ctx || !out || ctx->last_quotient_chunks.empty()) return P2G_E_BADARG;
    memcpy(out, ctx->last_quotient_chunks.data(), ctx->last_quotient_chunks.size() * sizeof(gl_t)); return P2G_OK;
}
extern "C" int32_t p2g_last_timings(p2g_ctx* ctx, p2g_timings* out) {
    if (!ctx || !out) return P2G_E_BADARG;
    *out = ctx->timings; return P2G_OK;
}
// enabled bit 0: per-stage CUDA-event timing; bit 1: keep stage dumps (zs values, quotient chunks)
extern "C" int32_t p2g_set_timing(p2g_ctx* ctx, int32_t enabled) {
    if (!ctx) return P2G_E_BADARG;
    ctx->timing = (enabled & 1) != 0; ctx->keep_debug = (enabled & 2) != 0; return P2G_OK;
}

extern "C" int32_t p2g_pow_grind(p2g_ctx* ctx, const uint64_t state[12], uint32_t pos, uint32_t pow_bits, uint64_t* nonce_out) {
    if (!ctx || !state || !nonce_out || pos >= 12 || pow_bits == 0 || pow_bits > 40) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    PowState ps; memcpy(ps.s, state, sizeof(ps.s));
    unsigned long long* d_best;
    CU(cudaMallocFromPoolAsync((void**)&d_best, 8, ctx->pool, ctx->st));
    const unsigned long long WIN = 1ull << 20;
    for (unsigned long long base = 0; base < (1ull << 44); base += WIN) {
        CU(cudaMemsetAsync(d_best, 0xFF, 8, ctx->st));
        P2G_COUNT_LAUNCH(1); pow_grind_kernel<<<(unsigned)(WIN / 256), 256, 0, ctx->st>>>(ps, (int)pos, (int)pow_bits, base, d_best);
        CU(cudaMemcpyAsync(ctx->pinned, d_best, 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(ctx_wait(ctx));
        unsigned long long best = *(unsigned long long*)ctx->pinned;
        if (best != ~0ull) { *nonce_out = best; cudaFreeAsync(d_best, ctx->st); return P2G_OK; }
    }
    cudaFreeAsync(d_best, ctx->st);
    return P2G_E_POW;
}

extern "C" int32_t p2g_fri_fold(p2g_ctx* ctx, const uint64_t* values_host, uint32_t log_len, uint32_t arity_bits, uint64_t shift,
                                const uint64_t beta[2], uint64_t* out_host) {
    if (!ctx || !values_host || !out_host || arity_bits == 0 || arity_bits > 4 || log_len < arity_bits || log_len > 26) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    const size_t len = (size_t)1 << log_len, chunks = len >> arity_bits;
    gl_t *d_in, *d_out; int rc;
    if ((rc = ctx_alloc(ctx, &d_in, 2 * len))) return rc;
    if ((rc = ctx_alloc(ctx, &d_out, 2 * chunks))) return rc;
    CU(cudaMemcpyAsync(d_in, values_host, 2 * len * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st));
    ext_t b = ext_make(beta[0], beta[1]);
    P2G_COUNT_LAUNCH(1); fri_fold_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, ctx->st>>>(d_in, (int)log_len, (int)arity_bits, gl_inv(shift),
                                                                         gl_inv(gl_root_of_unity((int)log_len)), gl_inv(gl_root_of_unity((int)arity_bits)),
                                                                         gl_inv((gl_t)1 << arity_bits), b, d_out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_host, d_out, 2 * chunks * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(ctx_wait(ctx));
    ctx_free(ctx, d_in); ctx_free(ctx, d_out);
    return P2G_OK;
}
