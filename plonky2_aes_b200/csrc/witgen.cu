// Device-side witness generation: plonky2's generate_partial_witness + set_lookup_wires
// (iop/generator.rs, plonk/prover.rs of the dependency pinned at /root/reference/Cargo.toml:12; entered
// from every `data.prove(pw)`, e.g. /root/reference/aes-gcm/src/circuit_gcm.rs:781) for the generator set
// the reference's circuits use: ArithmeticGate, LookupGate, equality, ConstantGate and PoseidonGate
// generators.  SURVEY.md section 8(f) row 4: with 8 GPUs proving at milliseconds per proof the host
// generators (0.3 M per AES-GCM proof) become the limiter of BASELINE config 5.
//
// The generator program (include/p2witness.h, the same description libp2witness.so interprets on the
// host) is LEVEL-SCHEDULED once at load: an op's level is one more than the highest level among the
// partitions it reads, so all ops of one level are independent.  The AES-GCM circuit of config 2 has
// 306 000 ops in 6 801 levels (median 16 ops per level: the GHASH and key-schedule chains are long and
// thin), so the unit of parallelism is the WITNESS, not the op: one warp evaluates one witness, lanes
// take the ops of the current level, __syncwarp() separates levels.  Many witnesses run side by side
// (a warp each); one witness takes a few milliseconds of one warp, i.e. nothing of the GPU's throughput.
// The result is the "extended slot vector" of p2w_generate_slots, bit for bit, which p2g_prove_slots'
// gather kernel turns into the wire matrix -- the witness never exists on the host.
#include "ctx.h"
#include "poseidon.cuh"
#include "../../include/p2witness.h"
#include <vector>
#include <algorithm>
#include <string.h>

#define PFAST_QUAL __constant__
#include "poseidon_fast.inc"
#undef PFAST_QUAL

enum { WG_CHECK0 = 0x80, WG_CHECK1 = 0x40 };   // output 0 / output 1 is already set: compare instead of write

struct p2g_wprog {
    uint32_t num_slots, ext_total, ext_mult, ext_pos, num_ops, num_levels, num_inputs, num_luts, num_poseidon, lut_entries;
    uint32_t lookups_total;
    struct WgOp* d_ops; uint32_t* d_level_off;                            // level-sorted packed records
    int32_t* d_in_slots;
    int32_t* d_key2entry;      // [num_luts][65536]: entry index of a key or -1
    int32_t* d_key2out;        // [num_luts][65536]: output of a key or -1 (LookupGenerator in one load)
    int32_t* d_lut_off;        // first entry of each LUT (num_luts + 1)
    int32_t* d_lookup_slots; int32_t* d_lookup_off; int32_t* d_lookup_padding;
    int32_t* d_poseidon_rows;
    int32_t* d_err;            // per-witness error flags of the last launch (lazily sized)
    uint32_t err_cap;
};

// one generator: 48 bytes, read as three 16-byte words.  kf = kind | WG_CHECK flags; s[4] = outputs / inputs as in
// p2w_program_desc.ops[1..4]; lut = ops[5]
struct __align__(16) WgOp { uint32_t kf; int32_t s[4]; int32_t lut; uint32_t pad[2]; gl_t c[2]; };
struct WgProg {
    uint32_t num_slots, ext_total, ext_mult, ext_pos, num_ops, num_levels, num_inputs, num_luts, num_poseidon;
    const WgOp* ops; const uint32_t* level_off;
    const int32_t* in_slots; const int32_t* key2entry; const int32_t* key2out; const int32_t* lut_off;
    const int32_t* lookup_slots; const int32_t* lookup_off; const int32_t* lookup_padding; const int32_t* poseidon_rows;
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ gl_t wg_sbox(gl_t x) { gl_t x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x, x2); return gl_mul(x3, x4); }
__device__ void wg_mds(gl_t s[12]) {
    const uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    gl_t o[12];
    for (int r = 0; r < 12; r++) {
        gl_t acc = 0;
        for (int i = 0; i < 12; i++) acc = gl_add(acc, gl_mul(s[(i + r) % 12], C[i]));
        if (r == 0) acc = gl_add(acc, gl_mul(s[0], 8));
        o[r] = acc;
    }
    for (int r = 0; r < 12; r++) s[r] = o[r];
}
// PoseidonGate generator with swap = 0 (gates/poseidon.rs): trace[c - 12] = wire c of the row, c = 12..134
__device__ __noinline__ void wg_poseidon_gate(const gl_t in[12], gl_t* trace) {
    gl_t st[12];
    for (int i = 0; i < 12; i++) st[i] = in[i];
    for (int i = 0; i < 123; i++) trace[i] = 0;            // swap (24) and deltas (25..28) are zero
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC_DEV[12 * r + i]);
        if (r != 0) for (int i = 0; i < 12; i++) trace[29 + 12 * (r - 1) + i - 12] = st[i];
        for (int i = 0; i < 12; i++) st[i] = wg_sbox(st[i]);
        wg_mds(st);
    }
    for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], PFAST_FIRST_C[i]);
    {
        gl_t t[11];
        for (int r = 0; r < 11; r++) { gl_t a = 0; for (int c = 0; c < 11; c++) a = gl_add(a, gl_mul(st[c + 1], PFAST_INIT[r * 11 + c])); t[r] = a; }
        for (int r = 0; r < 11; r++) st[r + 1] = t[r];
    }
    for (int r = 0; r < 22; r++) {
        trace[65 + r - 12] = st[0];
        st[0] = gl_add(wg_sbox(st[0]), PFAST_K[r]);
        gl_t s0 = gl_mul(st[0], 25);
        for (int j = 0; j < 11; j++) s0 = gl_add(s0, gl_mul(st[j + 1], PFAST_VROW[r * 11 + j]));
        for (int j = 0; j < 11; j++) st[j + 1] = gl_add(st[j + 1], gl_mul(st[0], PFAST_WCOL[r * 11 + j]));
        st[0] = s0;
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC_DEV[12 * (26 + r) + i]);
        for (int i = 0; i < 12; i++) trace[87 + 12 * r + i - 12] = st[i];
        for (int i = 0; i < 12; i++) st[i] = wg_sbox(st[i]);
        wg_mds(st);
    }
    for (int i = 0; i < 12; i++) trace[i] = st[i];         // outputs, wires 12..23
}
#endif

// one warp per witness; ext: [count][ext_total], zero-filled by the caller
__global__ void __launch_bounds__(128)
witgen_kernel(WgProg P, const gl_t* __restrict__ in_vals, uint32_t count, gl_t* ext_all, int32_t* __restrict__ err) {
#if defined(__CUDA_ARCH__)
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= count) return;
    // plain (L1-cached) accesses: a value written by another lane is read only after the __syncwarp() that ends its
    // level, which orders memory among the lanes of the warp (and is a compiler barrier); volatile accesses would go to
    // L2 every time and triple the latency of the level chain
#ifdef P2G_WITGEN_VOLATILE
    volatile gl_t* ext = ext_all + (size_t)w * P.ext_total;
#else
    gl_t* ext = ext_all + (size_t)w * P.ext_total;
#endif
    int bad = 0;
    for (uint32_t i = lane; i < P.num_inputs; i += 32) {
        const gl_t v = in_vals[(size_t)w * P.num_inputs + i];
        if (v >= GL_P) bad |= 1;
        ext[P.in_slots[i]] = v;
    }
    __syncwarp();
    // The record of the op this lane runs in the NEXT level is fetched before the current level's values are
    // awaited: a level then costs the round trips of its operand loads only (levels hold ~16 independent ops,
    // the chain of 6 801 levels is what takes the time).
    uint32_t k0 = P.level_off[0], k1 = P.num_levels ? P.level_off[1] : 0;
    WgOp nxt;
    if (k0 + lane < k1) nxt = P.ops[k0 + lane];
#pragma unroll 1
    for (uint32_t L = 0; L < P.num_levels; L++) {
        const uint32_t c0 = k0, c1 = k1;
        WgOp op = nxt;
        if (L + 1 < P.num_levels) {
            k0 = c1; k1 = P.level_off[L + 2];
            if (k0 + lane < k1) nxt = P.ops[k0 + lane];
        }
#pragma unroll 1
        for (uint32_t k = c0 + lane; k < c1; k += 32) {
            if (k != c0 + lane) op = P.ops[k];             // levels wider than a warp: the rest is read here
            const uint32_t kf = op.kf, kind = kf & 7;
            const int32_t s0 = op.s[0];
            gl_t r0 = 0, r1 = 0;
            bool two = false;
            if (kind == P2W_OP_ARITH) {
                const gl_t a = ext[op.s[1]], b = ext[op.s[2]], c = ext[op.s[3]];
                r0 = gl_add(gl_mul(op.c[0], gl_mul(a, b)), gl_mul(op.c[1], c));
            } else if (kind == P2W_OP_LOOKUP) {
                const gl_t x = ext[op.s[1]];
                const int32_t o = x <= 0xFFFF ? P.key2out[(size_t)op.lut * 65536 + (uint32_t)x] : -1;
                if (o < 0) bad |= 2; else r0 = (gl_t)o;
            } else if (kind == P2W_OP_EQ) {
                const gl_t diff = gl_sub(ext[op.s[2]], ext[op.s[3]]);
                r0 = diff == 0 ? 1 : 0;
                r1 = diff == 0 ? 0 : gl_inv(diff);
                two = true;
            } else if (kind == P2W_OP_CONST) {
                r0 = op.c[0];
            } else {                                       // P2W_OP_POSEIDON: s0 = row index
                const int32_t* pr = P.poseidon_rows + (size_t)25 * s0;
                gl_t in[12], trace[123];
                for (int i = 0; i < 12; i++) in[i] = ext[pr[1 + i]];
                wg_poseidon_gate(in, trace);
                for (int i = 0; i < 12; i++) {             // outputs (checked against values already set)
                    auto* o = ext + pr[13 + i];
                    if (kf & WG_CHECK0) { if (*o != trace[i]) bad |= 4; } else *o = trace[i];
                }
                auto* tr = ext + P.ext_pos + (size_t)111 * s0;
                for (int i = 0; i < 111; i++) tr[i] = trace[12 + i];
                continue;
            }
            if (kf & WG_CHECK0) { if (ext[s0] != r0) bad |= 4; } else ext[s0] = r0;
            if (two) { const int32_t s1 = op.s[1]; if (kf & WG_CHECK1) { if (ext[s1] != r1) bad |= 4; } else ext[s1] = r1; }
        }
        __syncwarp();
    }
    // set_lookup_wires: multiplicity of every table entry
    for (uint32_t l = 0; l < P.num_luts; l++) {
        unsigned long long* mult = (unsigned long long*)(ext_all + (size_t)w * P.ext_total + P.ext_mult + P.lut_off[l]);
        for (int32_t i = P.lookup_off[l] + (int32_t)lane; i < P.lookup_off[l + 1]; i += 32) {
            const gl_t x = ext[P.lookup_slots[i]];
            const int32_t e = x <= 0xFFFF ? P.key2entry[(size_t)l * 65536 + (uint32_t)x] : -1;
            if (e < 0) bad |= 2; else atomicAdd(mult + e, 1ull);
        }
        if (lane == 0 && P.lut_off[l + 1] > P.lut_off[l]) atomicAdd(mult, (unsigned long long)P.lookup_padding[l]);
    }
    if (bad) atomicOr(err + w, bad);
#endif
}

template <typename T>
static bool up(T** d, const std::vector<T>& h, cudaStream_t st) {
    const size_t bytes = (h.size() ? h.size() : 1) * sizeof(T);
    if (cudaMalloc((void**)d, bytes) != cudaSuccess) { *d = nullptr; return false; }
    return h.empty() || cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st) == cudaSuccess;
}

extern "C" int32_t p2g_wprog_free(p2g_ctx* ctx, p2g_wprog* p) {
    if (!ctx || !p) return P2G_E_BADARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    cudaFree(p->d_ops); cudaFree(p->d_level_off); cudaFree(p->d_in_slots);
    cudaFree(p->d_key2entry); cudaFree(p->d_key2out); cudaFree(p->d_lut_off); cudaFree(p->d_lookup_slots);
    cudaFree(p->d_lookup_off); cudaFree(p->d_lookup_padding); cudaFree(p->d_poseidon_rows); cudaFree(p->d_err);
    delete p;
    return P2G_OK;
}

extern "C" int32_t p2g_wprog_load(p2g_ctx* ctx, const p2w_program_desc* d, const int32_t* input_slots, uint32_t num_inputs,
                                  p2g_wprog** out) {
    if (!ctx || !d || !out || (num_inputs && !input_slots) || !d->num_slots) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    const uint32_t S = d->num_slots, K = d->num_ops;
    // ---- symbolic run on the host: which slots are set, which writes are re-writes, op levels ----
    std::vector<int32_t> level(S, 0);
    std::vector<uint8_t> has(S, 0);
    for (uint32_t i = 0; i < num_inputs; i++) {
        if (input_slots[i] < 0 || (uint32_t)input_slots[i] >= S) { ctx->err = "input slot out of range"; return P2G_E_BADARG; }
        has[input_slots[i]] = 1;
    }
    std::vector<uint8_t> kind(K);
    std::vector<int32_t> oplevel(K);
    int32_t max_level = 0;
    auto rd = [&](int32_t s, int32_t& lv) -> bool { if (s < 0 || (uint32_t)s >= S || !has[s]) return false; lv = std::max(lv, level[s]); return true; };
    for (uint32_t k = 0; k < K; k++) {
        const int32_t* op = d->ops + (size_t)6 * k;
        int32_t lv = 0;
        bool ok = true;
        uint8_t kf = (uint8_t)op[0];
        int32_t outs[13]; int nout = 0;
        switch (op[0]) {
        case P2W_OP_ARITH: ok = rd(op[2], lv) && rd(op[3], lv) && rd(op[4], lv); outs[nout++] = op[1]; break;
        case P2W_OP_LOOKUP: ok = rd(op[2], lv) && op[5] >= 0 && (uint32_t)op[5] < d->num_luts; outs[nout++] = op[1]; break;
        case P2W_OP_EQ: ok = rd(op[3], lv) && rd(op[4], lv); outs[nout++] = op[1]; outs[nout++] = op[2]; break;
        case P2W_OP_CONST: outs[nout++] = op[1]; break;
        case P2W_OP_POSEIDON: {
            ok = op[1] >= 0 && (uint32_t)op[1] < d->num_poseidon;
            if (ok) {
                const int32_t* pr = d->poseidon_rows + (size_t)25 * op[1];
                for (int i = 0; i < 12 && ok; i++) ok = rd(pr[1 + i], lv);
                for (int i = 0; i < 12; i++) outs[nout++] = pr[13 + i];
            }
            break; }
        default: ok = false;
        }
        if (!ok) { ctx->err = "generator program reads a partition that is never set (P2W_E_UNSET) or is malformed"; return P2G_E_BADARG; }
        lv += 1;
        bool any_set = false, all_set = true;
        for (int i = 0; i < nout; i++) {
            if (outs[i] < 0 || (uint32_t)outs[i] >= S) { ctx->err = "op output out of range"; return P2G_E_BADARG; }
            if (has[outs[i]]) { any_set = true; lv = std::max(lv, level[outs[i]] + 1); } else all_set = false;
        }
        if (op[0] == P2W_OP_POSEIDON) {
            if (any_set && !all_set) { ctx->err = "PoseidonGate outputs partially preset: not supported on the device"; return P2G_E_BADARG; }
            if (any_set) kf |= WG_CHECK0;
        } else {
            if (has[outs[0]]) kf |= WG_CHECK0;
            if (nout > 1 && has[outs[1]]) kf |= WG_CHECK1;
            if (nout > 1 && outs[0] == outs[1]) { ctx->err = "equality generator writes one partition twice"; return P2G_E_BADARG; }
        }
        for (int i = 0; i < nout; i++) if (!has[outs[i]]) { has[outs[i]] = 1; level[outs[i]] = lv; }
        kind[k] = kf; oplevel[k] = lv;
        max_level = std::max(max_level, lv);
    }
    size_t lookups_total = 0, lut_entries = 0;
    for (uint32_t l = 0; l < d->num_luts; l++) { lookups_total += d->lookup_counts[l]; lut_entries += d->lut_lens[l]; }
    for (size_t i = 0; i < lookups_total; i++)
        if (d->lookup_slots[i] < 0 || (uint32_t)d->lookup_slots[i] >= S || !has[d->lookup_slots[i]]) { ctx->err = "lookup input never set"; return P2G_E_BADARG; }
    // ---- level-sorted structure of arrays ----
    std::vector<uint32_t> order(K), level_off((size_t)max_level + 1, 0);
    for (uint32_t k = 0; k < K; k++) level_off[oplevel[k]]++;                 // levels are 1-based; slot 0 stays 0
    { uint32_t acc = 0; for (int32_t l = 1; l <= max_level; l++) { uint32_t c = level_off[l]; level_off[l] = acc; acc += c; } }
    std::vector<uint32_t> cursor(level_off);
    for (uint32_t k = 0; k < K; k++) order[cursor[oplevel[k]]++] = k;
    std::vector<uint32_t> offs((size_t)max_level + 1);
    for (int32_t l = 1; l <= max_level; l++) offs[l - 1] = level_off[l];
    offs[max_level] = K;
    std::vector<WgOp> ops_s(K);
    for (uint32_t j = 0; j < K; j++) {
        const uint32_t k = order[j];
        WgOp& o = ops_s[j];
        memset(&o, 0, sizeof(o));
        o.kf = kind[k];
        for (int f = 0; f < 4; f++) o.s[f] = d->ops[(size_t)6 * k + 1 + f];
        o.lut = d->ops[(size_t)6 * k + 5];
        o.c[0] = d->op_consts[(size_t)2 * k]; o.c[1] = d->op_consts[(size_t)2 * k + 1];
        if (o.c[0] >= GL_P || o.c[1] >= GL_P) { ctx->err = "non-canonical op constant"; return P2G_E_BADARG; }
    }
    // ---- lookup tables ----
    std::vector<int32_t> key2entry((size_t)d->num_luts * 65536, -1), key2out((size_t)d->num_luts * 65536, -1),
        lut_off(d->num_luts + 1, 0), lookup_off(d->num_luts + 1, 0);
    for (uint32_t l = 0; l < d->num_luts; l++) {
        lut_off[l + 1] = lut_off[l] + d->lut_lens[l];
        lookup_off[l + 1] = lookup_off[l] + d->lookup_counts[l];
        for (int32_t e = d->lut_lens[l] - 1; e >= 0; e--) {          // first occurrence of a key wins (as on the host)
            key2entry[(size_t)l * 65536 + d->lut_data[2 * ((size_t)lut_off[l] + e)]] = e;
            key2out[(size_t)l * 65536 + d->lut_data[2 * ((size_t)lut_off[l] + e)]] = d->lut_data[2 * ((size_t)lut_off[l] + e) + 1];
        }
    }
    p2g_wprog* p = new p2g_wprog();
    memset(p, 0, sizeof(*p));
    p->num_slots = S; p->num_ops = K; p->num_levels = (uint32_t)max_level; p->num_inputs = num_inputs; p->num_luts = d->num_luts;
    p->num_poseidon = d->num_poseidon; p->lut_entries = (uint32_t)lut_entries; p->lookups_total = (uint32_t)lookups_total;
    p->ext_mult = S; p->ext_pos = S + (uint32_t)lut_entries; p->ext_total = p->ext_pos + 111u * d->num_poseidon;
    std::vector<int32_t> in_s(input_slots, input_slots + num_inputs), lk(d->lookup_slots, d->lookup_slots + lookups_total),
        pad(d->lookup_padding, d->lookup_padding + d->num_luts),
        prow(d->num_poseidon ? d->poseidon_rows : nullptr, d->num_poseidon ? d->poseidon_rows + (size_t)25 * d->num_poseidon : nullptr);
    offs.push_back(K);            // the kernel reads level_off[L + 2] while it prefetches
    bool ok = up(&p->d_ops, ops_s, ctx->st) && up(&p->d_level_off, offs, ctx->st) &&
              up(&p->d_in_slots, in_s, ctx->st) && up(&p->d_key2entry, key2entry, ctx->st) && up(&p->d_key2out, key2out, ctx->st) &&
              up(&p->d_lut_off, lut_off, ctx->st) && up(&p->d_lookup_slots, lk, ctx->st) && up(&p->d_lookup_off, lookup_off, ctx->st) &&
              up(&p->d_lookup_padding, pad, ctx->st) && up(&p->d_poseidon_rows, prow, ctx->st);
    if (ok) ok = ctx_wait(ctx) == cudaSuccess;          // the host vectors die at the end of this function
    if (!ok) { cudaGetLastError(); p2g_wprog_free(ctx, p); ctx->err = "witness program upload"; return P2G_E_CUDA; }
    *out = p;
    return P2G_OK;
}

extern "C" uint32_t p2g_wprog_ext_slots(const p2g_wprog* p) { return p ? p->ext_total : 0; }
extern "C" uint32_t p2g_wprog_levels(const p2g_wprog* p) { return p ? p->num_levels : 0; }

// inputs (device, [count][num_inputs]) -> extended slot vectors (device, [count][ext_total]); d_err: [count]
int wprog_launch(p2g_ctx* ctx, const p2g_wprog* p, const gl_t* d_in, uint32_t count, gl_t* d_ext, int32_t* d_err) {
    CU(cudaMemsetAsync(d_ext, 0, (size_t)count * p->ext_total * sizeof(gl_t), ctx->st));
    CU(cudaMemsetAsync(d_err, 0, (size_t)count * sizeof(int32_t), ctx->st));
    WgProg P;
    P.num_slots = p->num_slots; P.ext_total = p->ext_total; P.ext_mult = p->ext_mult; P.ext_pos = p->ext_pos; P.num_ops = p->num_ops;
    P.num_levels = p->num_levels; P.num_inputs = p->num_inputs; P.num_luts = p->num_luts; P.num_poseidon = p->num_poseidon;
    P.ops = p->d_ops; P.level_off = p->d_level_off; P.in_slots = p->d_in_slots;
    P.key2entry = p->d_key2entry; P.key2out = p->d_key2out; P.lut_off = p->d_lut_off; P.lookup_slots = p->d_lookup_slots;
    P.lookup_off = p->d_lookup_off; P.lookup_padding = p->d_lookup_padding; P.poseidon_rows = p->d_poseidon_rows;
    // one warp per witness, one warp per block: the witnesses of a batch spread over the SMs
    witgen_kernel<<<count, 32, 0, ctx->st>>>(P, d_in, count, d_ext, d_err);
    P2G_COUNT_LAUNCH(1);
    CU(cudaGetLastError());
    return P2G_OK;
}
uint32_t wprog_ext_total(const p2g_wprog* p) { return p->ext_total; }
uint32_t wprog_num_inputs(const p2g_wprog* p) { return p->num_inputs; }

extern "C" int32_t p2g_wprog_generate(p2g_ctx* ctx, const p2g_wprog* p, const uint64_t* input_vals, uint32_t count, uint64_t* ext_out) {
    if (!ctx || !p || !input_vals || !ext_out || !count) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    gl_t *d_in, *d_ext; int32_t* d_err; int rc;
    if ((rc = ctx_alloc(ctx, &d_in, (size_t)count * p->num_inputs))) return rc;
    if ((rc = ctx_alloc(ctx, &d_ext, (size_t)count * p->ext_total))) { ctx_free(ctx, d_in); return rc; }
    if (cudaMallocFromPoolAsync((void**)&d_err, count * sizeof(int32_t), ctx->pool, ctx->st) != cudaSuccess) { ctx_free(ctx, d_in); ctx_free(ctx, d_ext); return P2G_E_CUDA; }
    std::vector<int32_t> err(count);
    cudaError_t e = cudaMemcpyAsync(d_in, input_vals, (size_t)count * p->num_inputs * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st);
    if (e == cudaSuccess) rc = wprog_launch(ctx, p, d_in, count, d_ext, d_err);
    if (e == cudaSuccess && rc == P2G_OK) e = cudaMemcpyAsync(ext_out, d_ext, (size_t)count * p->ext_total * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st);
    if (e == cudaSuccess && rc == P2G_OK) e = cudaMemcpyAsync(err.data(), d_err, count * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->st);
    if (e == cudaSuccess && rc == P2G_OK) e = ctx_wait(ctx);
    ctx_free(ctx, d_in); ctx_free(ctx, d_ext); ctx_free(ctx, d_err);
    if (rc) return rc;
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return P2G_E_CUDA; }
    for (uint32_t i = 0; i < count; i++)
        if (err[i]) {
            // a value outside its table usually also derails later generators: report the lookup first
            ctx->err = (err[i] & 2) ? "lookup input not in table (P2W_E_LOOKUP)"
                     : (err[i] & 4) ? "partition set twice with different values (P2W_E_CONFLICT)" : "non-canonical input value";
            return (err[i] & 2) ? P2W_E_LOOKUP : (err[i] & 4) ? P2W_E_CONFLICT : P2G_E_BADARG;
        }
    return P2G_OK;
}

// Batch form for pipelines that keep the witnesses in HBM: `count` witnesses are generated into caller-owned device
// memory (ext_dev: [count][ext_slots] words, flags_dev: [count] int32, 0 = ok or a P2W_E-style bit set: 1 non-canonical
// input, 2 lookup, 4 conflict); nothing is waited for -- the work is ordered on the context's stream, and
// p2g_prove_slots_dev on the SAME context (or after p2g_ctx_sync) consumes the vectors.
extern "C" int32_t p2g_wprog_generate_dev(p2g_ctx* ctx, const p2g_wprog* p, const uint64_t* input_vals_host, uint32_t count,
                                          uint64_t* ext_dev, int32_t* flags_dev) {
    if (!ctx || !p || !input_vals_host || !ext_dev || !flags_dev || !count) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    gl_t* d_in; int rc;
    if ((rc = ctx_alloc(ctx, &d_in, (size_t)count * p->num_inputs))) return rc;
    cudaError_t e = cudaMemcpyAsync(d_in, input_vals_host, (size_t)count * p->num_inputs * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st);
    if (e == cudaSuccess) rc = wprog_launch(ctx, p, d_in, count, ext_dev, flags_dev);
    ctx_free(ctx, d_in);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return P2G_E_CUDA; }
    return rc;
}
