// Device kernels of the prove() hot path beyond commitment: Z / partial products, lookup
// running sums, quotient evaluation, openings, FRI batch combination, FRI folding, proof-of-work
// grind and query gathering.  Each kernel cites the plonky2 function it replaces (dependency
// pinned at /root/reference/Cargo.toml:12, entered through `data.prove(pw)`,
// /root/reference/aes-gcm/src/circuit_gcm.rs:781).
#pragma once
#include "common.h"
#include "poseidon.cuh"
#include "../../include/p2gpu.h"

#define MAX_CH 2
#define MAX_ROUTED 80


// ---------------------------------------------------------------------------------------------
// affine maps x -> a*x + b and their block-wide exclusive scan (prefix products, running sums,
// the RE recurrence of the lookup argument are all instances)
// ---------------------------------------------------------------------------------------------
struct Aff { gl_t a, b; };
__device__ __forceinline__ Aff aff_id() { Aff r; r.a = 1; r.b = 0; return r; }
// f first, then g
__device__ __forceinline__ Aff aff_then(Aff f, Aff g) {
    Aff r; r.a = gl_mul(g.a, f.a); r.b = gl_add(gl_mul(g.a, f.b), g.b); return r;
}
__device__ __forceinline__ gl_t aff_apply(Aff f, gl_t x) { return gl_add(gl_mul(f.a, x), f.b); }
__device__ __forceinline__ Aff aff_shfl_up(Aff v, int d) {
    Aff r; r.a = __shfl_up_sync(0xffffffffu, v.a, d); r.b = __shfl_up_sync(0xffffffffu, v.b, d); return r;
}
// exclusive scan over the threads of a block (thread 0 gets the identity); sm: 32 entries
__device__ __forceinline__ Aff block_scan_exclusive(Aff mine, Aff* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    Aff cur = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Aff o = aff_shfl_up(cur, d);
        if (lane >= d) cur = aff_then(o, cur);
    }
    if (lane == 31) sm[warp] = cur;
    __syncthreads();
    if (warp == 0) {
        Aff w = lane < nwarps ? sm[lane] : aff_id();
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Aff o = aff_shfl_up(w, d);
            if (lane >= d) w = aff_then(o, w);
        }
        sm[lane] = w;
    }
    __syncthreads();
    Aff prev = aff_shfl_up(cur, 1);
    if (lane == 0) prev = aff_id();
    Aff base = warp > 0 ? sm[warp - 1] : aff_id();
    Aff r = aff_then(base, prev);
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------
// per-proof constants shared by several kernels (uploaded once per proof)
// ---------------------------------------------------------------------------------------------
struct ProofConsts {
    gl_t betas[MAX_CH], gammas[MAX_CH], alphas[MAX_CH];
    gl_t deltas[MAX_CH][4];                 // (a, b, alpha, delta) per challenge
    gl_t beta_kis[MAX_CH][MAX_ROUTED];      // beta_c * k_j
    gl_t k_is[MAX_ROUTED];
    gl_t pi_hash[4];
    gl_t alpha_pows[MAX_CH][256];           // alpha_c^k for the reduce_with_powers of the vanishing terms
    gl_t zh[16], zh_inv[16];                // Z_H on coset s (natural coset index), and inverse
    gl_t delta_pow_slots[MAX_CH];           // delta^(num_lut_slots)
};

struct CircuitDev {
    int logn, rate_bits, nch, R, W, NC, num_sel, num_lsel, num_consts, num_prods, qdf;
    int nlp, num_sldc, lut_degree, lu_degree, lu_slots, lut_slots, num_luts, num_gates, num_gate_constraints;
    int zs_cols;
};
// Coset shard of one proof (multi-GPU split of a single proof): this rank holds the leaf
// blocks [blk_first, blk_first + blk_count) of the 2^rate_bits cosets; its per-proof buffers (wires / zs / quotient
// LDE, FRI layers) cover only those blocks, the per-circuit tables (constants_sigmas LDE, domain) all of them.
struct ShardDev { unsigned int blk_first, blk_count; };

// ---------------------------------------------------------------------------------------------
// wires_permutation_partial_products_and_zs (plonk/prover.rs), step 1: per row and challenge
// the quotient of every chunk of 8 routed wires  prod(num)/prod(den).
// Output layout (temporary, finished by zs_scan_kernel): chunk ck < num_prods goes to the
// partial-product column ck, the last chunk to the Z column.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
zs_chunk_kernel(CircuitDev cd, const ProofConsts* __restrict__ pc, const gl_t* __restrict__ wires,
                const gl_t* __restrict__ sigmas, const gl_t* __restrict__ subgroup, gl_t* __restrict__ zs,
                gl_t* __restrict__ rowprod /*[nch][n]*/) {
    const size_t n = (size_t)1 << cd.logn;
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    if (r >= n) return;
    const gl_t x = subgroup[r], beta = pc->betas[ch], gamma = pc->gammas[ch];
    gl_t nums[16], dens[16];
    const int nchunks = cd.num_prods + 1;
#pragma unroll 1
    for (int ck = 0; ck < nchunks; ck++) {
        gl_t np = 1, dp = 1;
        int lo = ck * cd.qdf, hi = min(lo + cd.qdf, cd.R);
#pragma unroll 1
        for (int j = lo; j < hi; j++) {
            gl_t w = wires[(size_t)j * n + r];
            gl_t num = gl_add(gl_add(w, gl_mul(pc->beta_kis[ch][j], x)), gamma);
            gl_t den = gl_add(gl_add(w, gl_mul(beta, sigmas[(size_t)j * n + r])), gamma);
            np = gl_mul(np, num); dp = gl_mul(dp, den);
        }
        nums[ck] = np; dens[ck] = dp;
    }
    // batch inverse of the chunk denominators (one field inversion per thread)
    gl_t pre[16];
    gl_t acc = 1;
#pragma unroll
    for (int ck = 0; ck < 16; ck++) if (ck < nchunks) { pre[ck] = acc; acc = gl_mul(acc, dens[ck]); }
    gl_t inv = gl_inv(acc);
    gl_t rp = inv;                       // row quotient = prod(nums) / prod(dens)
#pragma unroll
    for (int ck = 15; ck >= 0; ck--) if (ck < nchunks) {
        gl_t q = gl_mul(nums[ck], gl_mul(inv, pre[ck]));
        inv = gl_mul(inv, dens[ck]);
        rp = gl_mul(rp, nums[ck]);
        int col = ck < cd.num_prods ? cd.nch + ch * cd.num_prods + ck : ch;
        zs[(size_t)col * n + r] = q;
    }
    rowprod[(size_t)ch * n + r] = rp;
}

// step 2: exclusive prefix product of the row quotients, one block per challenge
// (in place: rowprod[r] <- Z(g^r), Z(1) = 1)
__global__ void __launch_bounds__(1024)
zs_scan_kernel(CircuitDev cd, gl_t* __restrict__ rowprod) {
    __shared__ Aff sm[32];
    const size_t n = (size_t)1 << cd.logn;
    gl_t* P = rowprod + (size_t)blockIdx.x * n;
    const size_t per = (n + blockDim.x - 1) / blockDim.x;
    const size_t r0 = min(n, (size_t)threadIdx.x * per), r1 = min(n, r0 + per);
    Aff mine = aff_id();
    for (size_t r = r0; r < r1; r++) mine.a = gl_mul(mine.a, P[r]);
    Aff pre = block_scan_exclusive(mine, sm);
    gl_t z = pre.a;
    for (size_t r = r0; r < r1; r++) { gl_t q = P[r]; P[r] = z; z = gl_mul(z, q); }
}
// step 3: per row, Z and the running partial products  Z * chunk_0 * ... * chunk_t
__global__ void __launch_bounds__(128)
zs_apply_kernel(CircuitDev cd, const gl_t* __restrict__ rowprod, gl_t* __restrict__ zs) {
    const size_t n = (size_t)1 << cd.logn;
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    if (r >= n) return;
    gl_t z = rowprod[(size_t)ch * n + r];
    gl_t* PP = zs + ((size_t)cd.nch + (size_t)ch * cd.num_prods) * n;
    gl_t acc = z;
    for (int ck = 0; ck < cd.num_prods; ck++) {
        acc = gl_mul(acc, PP[(size_t)ck * n + r]);
        PP[(size_t)ck * n + r] = acc;
    }
    zs[(size_t)ch * n + r] = z;
}

// ---------------------------------------------------------------------------------------------
// compute_lookup_polys (plonk/prover.rs), step 1: per active row the row-local contributions.
// row_kind: 0 = none, 1 = LookupGate row, 2 = LookupTableGate row.
//   RE column      <- sum_s (in_s + b*out_s) * delta^(slots-1-s)            (LUT rows)
//   SLDC column k  <- cumulative sum over slot groups 0..k of
//                       mult_s / (alpha - looked_s)      (LUT rows)
//                     -1 / (alpha - looking_s)            (LookupGate rows)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
lookup_rows_kernel(CircuitDev cd, const ProofConsts* __restrict__ pc, const gl_t* __restrict__ wires,
                   const uint8_t* __restrict__ row_kind, gl_t* __restrict__ zs) {
    const size_t n = (size_t)1 << cd.logn;
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    if (r >= n) return;
    gl_t* base = zs + ((size_t)cd.nch * (1 + cd.num_prods) + (size_t)ch * cd.nlp) * n;
    const int kind = row_kind[r];
    if (kind == 0) {
        for (int k = 0; k < cd.nlp; k++) base[(size_t)k * n + r] = 0;
        return;
    }
    const gl_t da = pc->deltas[ch][0], db = pc->deltas[ch][1], dalpha = pc->deltas[ch][2], ddelta = pc->deltas[ch][3];
    gl_t cum = 0;
    if (kind == 2) {
        gl_t re = 0;
        for (int k = 0; k < cd.num_sldc; k++) {
            int s0 = k * cd.lut_degree, s1 = min(s0 + cd.lut_degree, cd.lut_slots);
            gl_t den[8], pre[8], mult[8];
            gl_t acc = 1;
            for (int s = s0; s < s1; s++) {
                gl_t in = wires[(size_t)(3 * s) * n + r], o = wires[(size_t)(3 * s + 1) * n + r];
                mult[s - s0] = wires[(size_t)(3 * s + 2) * n + r];
                re = gl_add(gl_mul(re, ddelta), gl_add(in, gl_mul(db, o)));
                den[s - s0] = gl_sub(dalpha, gl_add(in, gl_mul(da, o)));
                pre[s - s0] = acc; acc = gl_mul(acc, den[s - s0]);
            }
            gl_t inv = gl_inv(acc);         // one inversion per slot group (Montgomery's trick)
            for (int s = s1 - 1; s >= s0; s--) {
                cum = gl_add(cum, gl_mul(mult[s - s0], gl_mul(inv, pre[s - s0])));
                inv = gl_mul(inv, den[s - s0]);
            }
            base[(size_t)(k + 1) * n + r] = cum;
        }
        base[r] = re;
    } else {
        for (int k = 0; k < cd.num_sldc; k++) {
            int s0 = k * cd.lu_degree, s1 = min(s0 + cd.lu_degree, cd.lu_slots);
            gl_t den[8], pre[8];
            gl_t acc = 1;
            for (int s = s0; s < s1; s++) {
                gl_t in = wires[(size_t)(2 * s) * n + r], o = wires[(size_t)(2 * s + 1) * n + r];
                den[s - s0] = gl_sub(dalpha, gl_add(in, gl_mul(da, o)));
                pre[s - s0] = acc; acc = gl_mul(acc, den[s - s0]);
            }
            gl_t inv = gl_inv(acc);
            for (int s = s1 - 1; s >= s0; s--) {
                cum = gl_sub(cum, gl_mul(inv, pre[s - s0]));
                inv = gl_mul(inv, den[s - s0]);
            }
            base[(size_t)(k + 1) * n + r] = cum;
        }
        base[r] = 0;
    }
}

// step 2: backward recurrences over the rows (the lookup rows are deliberately upside down):
//   SLDC_k(row) = X(row+1) + cum_k(row),  X(row) = SLDC_last(row)     on active rows, 0 elsewhere
//   RE(row)     = RE(row+1) * delta^slots + B(row)                     on LUT rows, 0 elsewhere
__global__ void __launch_bounds__(1024)
lookup_scan_kernel(CircuitDev cd, const ProofConsts* __restrict__ pc, const uint8_t* __restrict__ row_kind,
                   gl_t* __restrict__ zs) {
    __shared__ Aff sm[32];
    const size_t n = (size_t)1 << cd.logn;
    const int ch = blockIdx.x;
    gl_t* base = zs + ((size_t)cd.nch * (1 + cd.num_prods) + (size_t)ch * cd.nlp) * n;
    const gl_t dpow = pc->delta_pow_slots[ch];
    // thread t owns rows [n - (t+1)*per, n - t*per) and walks them downwards
    const size_t per = (n + blockDim.x - 1) / blockDim.x;
    const size_t hi = n > (size_t)threadIdx.x * per ? n - (size_t)threadIdx.x * per : 0;
    const size_t lo = hi > per ? hi - per : 0;
    Aff fx = aff_id(), fre = aff_id();
    for (size_t r = hi; r-- > lo;) {
        int kind = row_kind[r];
        if (kind == 0) { fx.a = 0; fx.b = 0; } else fx.b = gl_add(fx.b, base[(size_t)cd.num_sldc * n + r]);
        if (kind == 2) { Aff g; g.a = dpow; g.b = base[r]; fre = aff_then(fre, g); } else { fre.a = 0; fre.b = 0; }
    }
    Aff px = block_scan_exclusive(fx, sm);
    Aff pre_ = block_scan_exclusive(fre, sm);
    gl_t X = px.b, RE = pre_.b;      // values entering from row `hi` (initial state 0)
    for (size_t r = hi; r-- > lo;) {
        int kind = row_kind[r];
        if (kind == 0) { X = 0; RE = 0; continue; }
        gl_t last = 0;
        for (int k = 0; k < cd.num_sldc; k++) {
            last = gl_add(X, base[(size_t)(k + 1) * n + r]);
            base[(size_t)(k + 1) * n + r] = last;
        }
        X = last;
        if (kind == 2) { RE = gl_add(gl_mul(RE, dpow), base[r]); base[r] = RE; } else { RE = 0; }
    }
}

// ---------------------------------------------------------------------------------------------
// compute_quotient_polys + eval_vanishing_poly_base_batch (plonk/prover.rs, vanishing_poly.rs).
// One thread per LDE point, reading the column-major bit-reversed LDE buffers directly
// (coalesced: consecutive threads = consecutive leaf indices).  "next row" = natural index + 8
// = same coset block, position bitrev(k + 1).  The result is scattered to natural order inside
// each coset block so the inverse NTT that follows can take natural input.
// ---------------------------------------------------------------------------------------------
#define PFAST_QUAL __constant__
#include "poseidon_fast.inc"
#undef PFAST_QUAL

// PoseidonGate::eval_unfiltered (gates/poseidon.rs): 123 constraints, evaluated in their upstream
// order and handed to `emit(k, value)`.  Wires: 0..11 input, 12..23 output, 24 swap, 25..28 delta,
// 29..64 / 65..86 / 87..134 the S-box inputs of the first full, the partial (sparse form) and the
// last full rounds.  Rare path (only circuits with Poseidon rows), kept out of line.
template <typename Emit>
__device__ __noinline__ void poseidon_gate_constraints(const gl_t* __restrict__ wl, size_t N, size_t j, Emit emit) {
#if defined(__CUDA_ARCH__)
    auto W = [&](int c) { return wl[(size_t)c * N + j]; };
    gl_t st[12]; int k = 0;
    const gl_t swap = W(24);
    emit(k++, gl_mul(swap, gl_sub(swap, 1)));
    for (int i = 0; i < 4; i++) {
        gl_t d = W(25 + i), a = W(i), b = W(i + 4);
        emit(k++, gl_sub(gl_mul(swap, gl_sub(b, a)), d));
        st[i] = gl_add(a, d); st[i + 4] = gl_sub(b, d);
    }
    for (int i = 8; i < 12; i++) st[i] = W(i);
    for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC_DEV[i]);
    for (int r = 0; r < 4; r++) {
        if (r != 0) for (int i = 0; i < 12; i++) { gl_t in = W(29 + 12 * (r - 1) + i); emit(k++, gl_sub(gl_canon(st[i]), in)); st[i] = in; }
        for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
        poseidon_mds_rc(st, r < 3 ? r + 1 : 30);           // next round constants; none before the partial rounds
    }
    for (int i = 0; i < 12; i++) st[i] = gl_add(gl_canon(st[i]), PFAST_FIRST_C[i]);
    {
        gl_t t[11];
        for (int r = 0; r < 11; r++) { gl_t a = 0; for (int c = 0; c < 11; c++) a = gl_add(a, gl_mul(st[c + 1], PFAST_INIT[r * 11 + c])); t[r] = a; }
        for (int r = 0; r < 11; r++) st[r + 1] = t[r];
    }
    for (int r = 0; r < 22; r++) {
        gl_t in = W(65 + r);
        emit(k++, gl_sub(st[0], in));
        st[0] = gl_add(gl_canon(poseidon_sbox(in)), PFAST_K[r]);
        gl_t s0 = gl_mul(st[0], 25);
        for (int q = 0; q < 11; q++) s0 = gl_add(s0, gl_mul(st[q + 1], PFAST_VROW[r * 11 + q]));
        for (int q = 0; q < 11; q++) st[q + 1] = gl_add(st[q + 1], gl_mul(st[0], PFAST_WCOL[r * 11 + q]));
        st[0] = s0;
    }
    for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC_DEV[12 * 26 + i]);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) { gl_t in = W(87 + 12 * r + i); emit(k++, gl_sub(gl_canon(st[i]), in)); st[i] = in; }
        for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
        poseidon_mds_rc(st, 27 + r);                          // rows 27..29, then 30 = zero
    }
    for (int i = 0; i < 12; i++) emit(k++, gl_sub(gl_canon(st[i]), W(12 + i)));
#endif
}

// lazy helpers of the quotient kernel (device only): results are any-u64 residues that only feed
// further multiplications or a final gl_canon
#if defined(__CUDA_ARCH__)
// a (any) + b (canonical) -> any u64, same residue: on carry add 2^64 = EPS, as the all-ones mask
// -carry on the low word.  (The carry is materialised with addc: an add chain must not be continued
// with subc -- ptxas keeps the hardware flag, where borrow = !carry.)
__device__ __forceinline__ gl_t qadd_lazy(gl_t a, gl_t b) {
    gl_t d;
    asm("{\n\t.reg .u32 a0, a1, b0, b1, m;\n\t"
        "mov.b64 {a0,a1}, %1;\n\tmov.b64 {b0,b1}, %2;\n\t"
        "add.cc.u32 a0, a0, b0;\n\taddc.cc.u32 a1, a1, b1;\n\taddc.u32 m, 0, 0;\n\tneg.s32 m, m;\n\t"
        "add.cc.u32 a0, a0, m;\n\taddc.u32 a1, a1, 0;\n\t"
        "mov.b64 %0, {a0,a1};\n\t}" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ gl_t pmul2(gl_t a, gl_t b, gl_t c, gl_t d) { return gl_mul2_lazy(a, b, c, d); }
#else
__device__ __forceinline__ gl_t qadd_lazy(gl_t a, gl_t b) { return gl_add_lazy(a, b); }
__device__ __forceinline__ gl_t pmul2(gl_t a, gl_t b, gl_t c, gl_t d) { return gl_add(gl_mul(a, b), gl_mul(c, d)); }
__device__ __forceinline__ gl_t pmul(gl_t a, gl_t b) { return gl_mul(a, b); }
__device__ __forceinline__ gl_t pmul_add(gl_t a, gl_t b, gl_t c) { return gl_add(gl_mul(a, b), gl_canon(c)); }
struct Acc160 { gl_t v; };
__device__ __forceinline__ void acc_mul(Acc160& A, gl_t a, gl_t b) { A.v = gl_add(A.v, gl_mul(a, b)); }
__device__ __forceinline__ gl_t acc_fold(const Acc160& A) { return A.v; }
#endif

__device__ __forceinline__ gl_t gate_filter(int row, int gs, int ge, gl_t s, bool many) {
    gl_t f = 1;
#pragma unroll 1
    for (int i = gs; i < ge; i++) if (i != row) f = gl_mul(f, gl_sub((gl_t)i, s));
    if (many) f = gl_mul(f, gl_sub(0xFFFFFFFFULL, s));
    return f;
}

// get_lut_poly(lut, deltas).eval(delta) (plonk/vanishing_poly.rs): sum_e (in_e + b*out_e) * delta^(degree-1-e)
// with the table zero-padded to `degree` = rows * slots entries.  One block per (lut, challenge).
__global__ void __launch_bounds__(1024)
lut_eval_kernel(CircuitDev cd, const ProofConsts* __restrict__ pc, const uint16_t* __restrict__ lut_data,
                const int* __restrict__ lut_off, const int* __restrict__ lut_len, gl_t* __restrict__ out /*[nch][8]*/) {
    __shared__ gl_t sm[1024];
    const int lut = blockIdx.x, ch = blockIdx.y;
    const int len = lut_len[lut];
    const int degree = ((len + cd.lut_slots - 1) / cd.lut_slots) * cd.lut_slots;
    const gl_t b = pc->deltas[ch][1], delta = pc->deltas[ch][3];
    const uint16_t* data = lut_data + 2 * (size_t)lut_off[lut];
    const int chunk = (degree + blockDim.x - 1) / blockDim.x;
    const int e0 = min(degree, (int)threadIdx.x * chunk), e1 = min(degree, e0 + chunk);
    gl_t acc = 0;
    for (int e = e0; e < e1; e++) {
        gl_t combo = e < len ? gl_add((gl_t)data[2 * e], gl_mul(b, (gl_t)data[2 * e + 1])) : 0;
        acc = gl_add(gl_mul(acc, delta), combo);
    }
    sm[threadIdx.x] = gl_mul(acc, gl_pow(delta, (uint64_t)(degree - e1)));
    __syncthreads();
    for (int d = blockDim.x / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) sm[threadIdx.x] = gl_add(sm[threadIdx.x], sm[threadIdx.x + d]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[ch * 8 + lut] = sm[0];
}

template <bool HAS_POSEIDON>
__global__ void __launch_bounds__(128)
quotient_kernel(CircuitDev cd, ShardDev sh, const ProofConsts* __restrict__ pc, const gl_t* __restrict__ lut_evals, const p2g_gate* __restrict__ gates,
                const gl_t* __restrict__ cs, const gl_t* __restrict__ wl, const gl_t* __restrict__ zl,
                const gl_t* __restrict__ domain, const gl_t* __restrict__ l0inv, gl_t* __restrict__ out) {
    const int logn = cd.logn;
    const size_t n = (size_t)1 << logn;
    // per-proof arrays (wl, zl, out) hold this shard's blocks only: N = leaves held, j = local leaf index; the
    // per-circuit arrays are addressed through the pointer offsets below
    const size_t N = (size_t)sh.blk_count << logn, NG = n << cd.rate_bits, j_off = (size_t)sh.blk_first << logn;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    cs += j_off; domain += j_off; l0inv += j_off;
    const uint32_t blk_l = (uint32_t)(j >> logn), blk = blk_l + sh.blk_first, within = (uint32_t)(j & (n - 1));
    const uint32_t k = gl_bitrev(within, logn);
    const size_t jn = ((size_t)blk_l << logn) + gl_bitrev((k + 1) & (uint32_t)(n - 1), logn);
    const uint32_t coset = gl_bitrev(blk, cd.rate_bits);
    const gl_t x = domain[j];
    const gl_t zh = pc->zh[coset];
    const gl_t l0 = gl_mul(zh, l0inv[j]);       // L_0(x) = Z_H(x) / (n (x - 1)), inverse tabulated at circuit load
    const int nch = cd.nch, R = cd.R;
    // sum_k alpha^k term_k per challenge: 128-bit products accumulated in 160 bits, one fold at the end
    Acc160 acc[MAX_CH];
#pragma unroll
    for (int c = 0; c < MAX_CH; c++) acc[c] = Acc160{};
    int t = 0;   // running term index
#define ADD_TERM(idx, val) do { gl_t v_ = (val); _Pragma("unroll") for (int c_ = 0; c_ < MAX_CH; c_++) if (c_ < nch) \
        acc_mul(acc[c_], v_, pc->alpha_pows[c_][(idx)]); } while (0)
    // Z(x) - 1 terms
    for (int c = 0; c < nch; c++) ADD_TERM(t + c, gl_mul(l0, gl_sub(zl[(size_t)c * N + j], 1)));
    t += nch;
    // partial products: both challenges share the wire / sigma loads
    {
        gl_t prev[MAX_CH];
        for (int c = 0; c < nch; c++) prev[c] = zl[(size_t)c * N + j];
        // the wire / sigma values of routed wire w+1 are loaded while wire w is multiplied in
        gl_t wv_n = wl[j], sv_n = cs[(size_t)cd.NC * NG + j];
#pragma unroll 1
        for (int ck = 0; ck <= cd.num_prods; ck++) {
            gl_t np[MAX_CH], dp[MAX_CH];
#pragma unroll
            for (int c = 0; c < MAX_CH; c++) { np[c] = 1; dp[c] = 1; }
            int lo = ck * cd.qdf, hi = min(lo + cd.qdf, R);
#pragma unroll 1
            for (int w = lo; w < hi; w++) {
                const gl_t wv = wv_n, sv = sv_n;
                if (w + 1 < R) { wv_n = wl[(size_t)(w + 1) * N + j]; sv_n = cs[(size_t)(cd.NC + w + 1) * NG + j]; }
#pragma unroll
                for (int c = 0; c < MAX_CH; c++) if (c < nch) {
                    // lazy residues: they only feed the running products
                    gl_t num = qadd_lazy(pmul_add(pc->beta_kis[c][w], x, wv), pc->gammas[c]);
                    gl_t den = qadd_lazy(pmul_add(pc->betas[c], sv, wv), pc->gammas[c]);
                    np[c] = pmul(np[c], num); dp[c] = pmul(dp[c], den);
                }
            }
#pragma unroll
            for (int c = 0; c < MAX_CH; c++) if (c < nch) {
                gl_t next = ck == cd.num_prods ? zl[(size_t)c * N + jn] : zl[(size_t)(nch + c * cd.num_prods + ck) * N + j];
                gl_t term = gl_sub(gl_mul(prev[c], np[c]), gl_mul(next, dp[c]));
                ADD_TERM(t + c * (cd.num_prods + 1) + ck, term);
                prev[c] = next;
            }
        }
        t += nch * (cd.num_prods + 1);
    }
    // lookup terms (check_lookup_constraints_batch)
    if (cd.num_luts > 0) {
        const int zpp = nch * (1 + cd.num_prods);
        const gl_t* lsel = cs + (size_t)cd.num_sel * NG;       // lookup selector columns
        const gl_t s_trans_sre = lsel[0 * NG + j], s_trans_ldc = lsel[1 * NG + j], s_init = lsel[2 * NG + j], s_last = lsel[3 * NG + j];
        const int n_lookup_terms = 4 + cd.num_luts + 2 * cd.num_sldc;
#pragma unroll 1
        for (int c = 0; c < nch; c++) {
            const gl_t da = pc->deltas[c][0], db = pc->deltas[c][1], dalpha = pc->deltas[c][2], ddelta = pc->deltas[c][3];
            const gl_t* lz = zl + (size_t)(zpp + c * cd.nlp) * N;
            const int tb = t + c * n_lookup_terms;
            const gl_t z_re = lz[j], next_z_re = lz[jn];
            ADD_TERM(tb + 0, gl_mul(s_last, lz[(size_t)cd.num_sldc * N + j]));
            ADD_TERM(tb + 1, gl_mul(s_init, lz[(size_t)1 * N + j]));
            ADD_TERM(tb + 2, gl_mul(s_init, z_re));
#pragma unroll 1
            for (int r = 0; r < cd.num_luts; r++)
                ADD_TERM(tb + 3 + r, gl_mul(lsel[(size_t)(4 + r) * NG + j], gl_sub(z_re, lut_evals[c * 8 + r])));
            gl_t re_cur = next_z_re;
            const int tt = tb + 4 + cd.num_luts;   // index of the first per-poly term (after RE transition at tt-1)
#pragma unroll 1
            for (int poly = 0; poly < cd.num_sldc; poly++) {
                int a0 = poly * cd.lut_degree, a1 = min(a0 + cd.lut_degree, cd.lut_slots);
                int b0 = poly * cd.lu_degree, b1 = min(b0 + cd.lu_degree, cd.lu_slots);
                // prod_i f_i and sum_i m_i prod_{j != i} f_j by the running pair (S, P) <- (S f + m P, P f)
                gl_t lut_prod = 1, lut_sum = 0, lu_prod = 1, lu_sum = 0;
                gl_t in_n = 0, o_n = 0, mult_n = 0;
                if (a0 < a1) { in_n = wl[(size_t)(3 * a0) * N + j]; o_n = wl[(size_t)(3 * a0 + 1) * N + j]; mult_n = wl[(size_t)(3 * a0 + 2) * N + j]; }
#pragma unroll 1
                for (int s = a0; s < a1; s++) {
                    const gl_t in = in_n, o = o_n, mult = mult_n;
                    if (s + 1 < a1) { in_n = wl[(size_t)(3 * s + 3) * N + j]; o_n = wl[(size_t)(3 * s + 4) * N + j]; mult_n = wl[(size_t)(3 * s + 5) * N + j]; }
                    gl_t f = gl_sub(dalpha, gl_canon(pmul_add(da, o, in)));
                    re_cur = pmul_add(re_cur, ddelta, pmul_add(db, o, in));
                    lut_sum = pmul2(lut_sum, f, mult, lut_prod);
                    lut_prod = pmul(lut_prod, f);
                }
                if (b0 < b1) { in_n = wl[(size_t)(2 * b0) * N + j]; o_n = wl[(size_t)(2 * b0 + 1) * N + j]; }
#pragma unroll 1
                for (int s = b0; s < b1; s++) {
                    const gl_t in = in_n, o = o_n;
                    if (s + 1 < b1) { in_n = wl[(size_t)(2 * s + 2) * N + j]; o_n = wl[(size_t)(2 * s + 3) * N + j]; }
                    gl_t f = gl_sub(dalpha, gl_canon(pmul_add(da, o, in)));
                    lu_sum = pmul_add(lu_sum, f, lu_prod);
                    lu_prod = pmul(lu_prod, f);
                }
                lut_sum = gl_canon(lut_sum); lu_sum = gl_canon(lu_sum);
                gl_t cur = lz[(size_t)(poly + 1) * N + j];
                gl_t prev = poly == 0 ? lz[(size_t)cd.num_sldc * N + jn] : lz[(size_t)poly * N + j];
                gl_t diff = gl_sub(cur, prev);
                ADD_TERM(tt + 2 * poly, gl_mul(s_trans_sre, gl_sub(gl_mul(lut_prod, diff), lut_sum)));
                ADD_TERM(tt + 2 * poly + 1, gl_mul(s_trans_ldc, gl_add(gl_mul(lu_prod, diff), lu_sum)));
            }
            ADD_TERM(tt - 1, gl_mul(s_trans_sre, gl_sub(z_re, gl_canon(re_cur))));
        }
        t += nch * n_lookup_terms;
    }
    // gate constraints: slot k accumulates filter * constraint_k over all gates
    {
        const gl_t* gconst = cs + (size_t)(cd.num_sel + cd.num_lsel) * NG;
        const bool many = cd.num_sel > 1;
        gl_t f_arith = 0, f_const = 0, f_pi = 0, f_pos = 0;
        int arith_ops = 0, nconst = 0;
        bool has_poseidon = false;
#pragma unroll 1
        for (int g = 0; g < cd.num_gates; g++) {
            const p2g_gate G = gates[g];
            if (G.num_constraints == 0) continue;
            gl_t f = gate_filter(g, G.group_start, G.group_end, cs[(size_t)G.selector_index * NG + j], many);
            if (G.kind == P2G_GATE_ARITHMETIC) { f_arith = f; arith_ops = G.param0; }
            else if (G.kind == P2G_GATE_CONSTANT) { f_const = f; nconst = G.param0; }
            else if (G.kind == P2G_GATE_PUBLIC_INPUT) f_pi = f;
            else if (G.kind == P2G_GATE_POSEIDON) { f_pos = f; has_poseidon = true; }
        }
        const gl_t c0 = cd.num_consts > 0 ? gconst[j] : 0, c1 = cd.num_consts > 1 ? gconst[NG + j] : 0;
        // contribution of the cheap gates to constraint slot k
        auto small_gates = [&](int k) -> gl_t {
            gl_t v = 0;
            if (k < arith_ops) {
                gl_t m0 = wl[(size_t)(4 * k) * N + j], m1 = wl[(size_t)(4 * k + 1) * N + j];
                gl_t ad = wl[(size_t)(4 * k + 2) * N + j], o = wl[(size_t)(4 * k + 3) * N + j];
                gl_t comp = gl_add(gl_mul(gl_mul(m0, m1), c0), gl_mul(ad, c1));
                v = gl_mul(f_arith, gl_sub(o, comp));
            }
            if (k < nconst) v = gl_add(v, gl_mul(f_const, gl_sub(k == 0 ? c0 : c1, wl[(size_t)k * N + j])));
            if (k < 4 && f_pi) v = gl_add(v, gl_mul(f_pi, gl_sub(wl[(size_t)k * N + j], pc->pi_hash[k])));
            return v;
        };
        if (HAS_POSEIDON && has_poseidon) {
            Acc160 a0 = acc[0], a1 = acc[1];
            const int tbase = t;
            poseidon_gate_constraints(wl, N, j, [&](int k, gl_t cval) {
                gl_t v = gl_add(gl_mul(f_pos, cval), small_gates(k));
                acc_mul(a0, v, pc->alpha_pows[0][tbase + k]);
                if (nch > 1) acc_mul(a1, v, pc->alpha_pows[1][tbase + k]);
            });
            acc[0] = a0; acc[1] = a1;
        } else {
#pragma unroll 1
            for (int k = 0; k < cd.num_gate_constraints; k++) ADD_TERM(t + k, small_gates(k));
        }
    }
#undef ADD_TERM
    const gl_t zi = pc->zh_inv[coset];
    for (int c = 0; c < nch; c++)
        out[(size_t)c * N + ((size_t)blk_l << logn) + k] = gl_mul(acc_fold(acc[c]), zi);
}

// Recover the quotient chunk coefficients from the 8 per-coset inverse NTTs:
//   t_c[j] = 7^(-n c)/8 * sum_s w_8^(-s c) * h_s^(-j) * a_s[j],  h_s = 7 w_N^s
// in: [nch][8 blocks (bit-reversed coset order)][n]; table: [8 (natural coset s)][n] = h_s^(-j)/8
__global__ void __launch_bounds__(256)
quotient_combine_kernel(int logn, int nch, unsigned int blk_count, const gl_t* __restrict__ in, const gl_t* __restrict__ table,
                        const gl_t* __restrict__ w8inv_pows /*[8]*/, const gl_t* __restrict__ shift_n_inv_pows /*[8]*/,
                        gl_t* __restrict__ out /*[nch*8][n]*/) {
    const size_t n = (size_t)1 << logn;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    if (j >= n) return;
    gl_t v[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        uint32_t blk = gl_bitrev((uint32_t)s, 3);
        // `in` is the all-gather of the shards' [nch][blk_count][n] arrays (blk_count = 8: one shard, plain layout)
        const uint32_t rk = blk / blk_count, bl = blk % blk_count;
        v[s] = gl_mul(in[(((size_t)rk * nch + ch) * blk_count + bl) * n + j], table[(size_t)s * n + j]);
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
        gl_t acc = 0;
#pragma unroll
        for (int s = 0; s < 8; s++) acc = gl_add(acc, gl_mul(v[s], w8inv_pows[(s * c) & 7]));
        out[((size_t)ch * 8 + c) * n + j] = gl_mul(acc, shift_n_inv_pows[c]);
    }
}

// ---------------------------------------------------------------------------------------------
// OpeningSet::new (plonk/proof.rs): f(zeta) for every committed polynomial.
// powers: zp[j] = zeta^j (ext, interleaved c0,c1).  One block per polynomial.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ext_powers_kernel(ext_t z, const ext_t* __restrict__ z_pow2 /*[32]: z^(2^b)*/, size_t count, gl_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    ext_t r = ext_make(1, 0);
    size_t e = j;
    for (int b = 0; e; b++, e >>= 1) if (e & 1) r = ext_mul(r, z_pow2[b]);
    out[2 * j] = r.c0; out[2 * j + 1] = r.c1;
    (void)z;
}
__global__ void __launch_bounds__(256)
eval_polys_kernel(const gl_t* const* __restrict__ polys, const gl_t* __restrict__ zp, size_t n, gl_t* __restrict__ out) {
    __shared__ gl_t s0[256], s1[256];
    const gl_t* c = polys[blockIdx.x];
    // 160-bit lazy sums; four coefficients per trip keep 12 loads in flight per thread
    Acc160 A0 = Acc160{}, A1 = Acc160{};
#pragma unroll 4
    for (size_t j = threadIdx.x; j < n; j += blockDim.x) {
        gl_t cv = c[j];
        acc_mul(A0, cv, zp[2 * j]); acc_mul(A1, cv, zp[2 * j + 1]);
    }
    gl_t a0 = gl_canon(acc_fold(A0)), a1 = gl_canon(acc_fold(A1));
    s0[threadIdx.x] = a0; s1[threadIdx.x] = a1;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) { s0[threadIdx.x] = gl_add(s0[threadIdx.x], s0[threadIdx.x + d]); s1[threadIdx.x] = gl_add(s1[threadIdx.x], s1[threadIdx.x + d]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = s0[0]; out[2 * blockIdx.x + 1] = s1[0]; }
}

// ---------------------------------------------------------------------------------------------
// PolynomialBatch::prove_openings (fri/oracle.rs): composition polynomial of one batch,
// comp[k] = sum_j alpha^j f_j[k], output as two base columns.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fri_compose_kernel(const gl_t* const* __restrict__ polys, int npolys, const gl_t* __restrict__ apow /*[npolys][2]*/, size_t n,
                   gl_t* __restrict__ out_c0, gl_t* __restrict__ out_c1) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    // sum_j alpha^j f_j[k] with tabulated powers: two base multiplications per polynomial,
    // accumulated lazily in 160 bits (the Horner form costs a full extension multiply each)
    Acc160 A0 = Acc160{}, A1 = Acc160{};
#pragma unroll 4
    for (int j = 0; j < npolys; j++) {
        const gl_t v = polys[j][k];
        acc_mul(A0, v, apow[2 * j]); acc_mul(A1, v, apow[2 * j + 1]);
    }
    out_c0[k] = gl_canon(acc_fold(A0)); out_c1[k] = gl_canon(acc_fold(A1));
}
// final(x) = alpha^{|b1|} (comp0(x) - comp0(zeta)) / (x - zeta) + (comp1(x) - comp1(g zeta)) / (x - g zeta)
// lde: [4][N] = comp0.c0, comp0.c1, comp1.c0, comp1.c1 on the LDE domain (bit-reversed); out [N][2]
__global__ void __launch_bounds__(256)
fri_final_values_kernel(const gl_t* __restrict__ lde, size_t N, const gl_t* __restrict__ domain /* at this shard's first point */,
                        ext_t zeta, ext_t zeta_next, ext_t comp0_at, ext_t comp1_at, ext_t shift0, gl_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const gl_t x = domain[j];
    ext_t v0 = ext_sub(ext_make(lde[j], lde[N + j]), comp0_at);
    ext_t v1 = ext_sub(ext_make(lde[2 * N + j], lde[3 * N + j]), comp1_at);
    ext_t d0 = ext_make(gl_sub(x, zeta.c0), gl_neg(zeta.c1));
    ext_t d1 = ext_make(gl_sub(x, zeta_next.c0), gl_neg(zeta_next.c1));
    ext_t r = ext_add(ext_mul(shift0, ext_mul(v0, ext_inv(d0))), ext_mul(v1, ext_inv(d1)));
    out[2 * j] = r.c0; out[2 * j + 1] = r.c1;
}

// ---------------------------------------------------------------------------------------------
// fri_committed_trees (fri/prover.rs), fold step in the VALUE domain: the reference folds
// coefficients (coeffs'[k] = sum_i coeffs[16k+i] beta^i) and re-FFTs; evaluating that on the
// next coset gives, for chunk k of 16 bit-reversed values on {x0 * w_16^m}:
//   r = iDFT16(values)  (coefficients of Q(x0 Y)),   out[k] = sum_i r_i (beta / x0)^i
// values: [len][2] bit-reversed order on coset shift*<w_len>;  out: [len/arity][2]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fri_fold_kernel(const gl_t* __restrict__ values, int log_len, int arity_bits, gl_t shift_inv, gl_t w_len_inv,
                gl_t w_arity_inv, gl_t arity_inv, ext_t beta, gl_t* __restrict__ out, size_t chunk_first = 0, size_t chunk_count = 0) {
    // log_len: size of the WHOLE layer; a coset shard holds the chunks [chunk_first, chunk_first + chunk_count)
    // (values / out point at its first chunk); chunk_count = 0: the whole layer
    const int arity = 1 << arity_bits;
    const size_t chunks = chunk_count ? chunk_count : (size_t)1 << (log_len - arity_bits);
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= chunks) return;
    ext_t u[16];
    // u[m] = value at x0 * w_arity^m = chunk entry bitrev(m)
    for (int m = 0; m < arity; m++) {
        int tpos = gl_bitrev((uint32_t)m, arity_bits);
        u[m] = ext_make(values[2 * (k * arity + tpos)], values[2 * (k * arity + tpos) + 1]);
    }
    // x0^-1 = shift^-1 * w_len^-(bitrev_{log_len - arity_bits}(k))
    uint32_t e = gl_bitrev((uint32_t)(k + chunk_first), log_len - arity_bits);
    gl_t x0_inv = gl_mul(shift_inv, gl_pow(w_len_inv, e));
    ext_t y = ext_mul_base(beta, x0_inv);
    // out = sum_i r_i y^i with r_i = (1/arity) sum_m u[m] w^-(i m): evaluate directly (arity <= 16)
    gl_t wp[16];
    wp[0] = 1;
    for (int i = 1; i < arity; i++) wp[i] = gl_mul(wp[i - 1], w_arity_inv);
    ext_t acc = ext_make(0, 0);
    for (int i = arity - 1; i >= 0; i--) {
        ext_t r = ext_make(0, 0);
        for (int m = 0; m < arity; m++) r = ext_add(r, ext_mul_base(u[m], wp[(i * m) & (arity - 1)]));
        acc = ext_add(ext_mul(acc, y), r);
    }
    acc = ext_mul_base(acc, arity_inv);
    out[2 * k] = acc.c0; out[2 * k + 1] = acc.c1;
}

// ---------------------------------------------------------------------------------------------
// fri_proof_of_work (fri/prover.rs): lowest nonce in [base, base + count) whose response
// (state[7] after the permutation) has pow_bits leading zeros; atomicMin keeps the minimum.
// ---------------------------------------------------------------------------------------------
struct PowState { gl_t s[12]; };
__global__ void __launch_bounds__(256)
pow_grind_kernel(PowState st, int pos, int pow_bits, unsigned long long base, unsigned long long* __restrict__ best) {
    const unsigned long long cand = base + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    gl_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {          // select, not an indexed store: the state stays in registers
        const gl_t m = (gl_t)0 - (gl_t)(i == pos);
        s[i] = (st.s[i] & ~m) | (cand & m);
    }
    poseidon_permute(s);
    if ((s[7] >> (64 - pow_bits)) == 0) atomicMin(best, cand);
}

// ---------------------------------------------------------------------------------------------
// fri_prover_query_rounds: gather leaf rows and Merkle paths of all queries into one buffer.
// One block per (query, tree).  Tree descriptor: column-major (LDE batches) or row-major (FRI).
// ---------------------------------------------------------------------------------------------
struct GatherTree {
    const gl_t* data; const gl_t* digests;
    unsigned long long col_stride;   // column-major stride (N); 0 = row-major
    unsigned int leaf_len, log_leaves, path_len, index_shift;  // leaf index = query_index >> index_shift
    unsigned long long out_offset;   // offset inside one query's record
    unsigned long long leaf_first;   // coset shard: first leaf held (data / digests are local); 0 otherwise
};
__global__ void __launch_bounds__(128)
query_gather_kernel(const GatherTree* __restrict__ trees, int ntrees, const unsigned long long* __restrict__ qidx,
                    unsigned long long record_words, gl_t* __restrict__ out) {
    const int q = blockIdx.x, t = blockIdx.y;
    const GatherTree T = trees[t];
    if (qidx[q] == ~0ull) return;                    // a query another shard owns: its record stays zero
    const size_t leaf = (size_t)(qidx[q] >> T.index_shift) - (size_t)T.leaf_first;
    gl_t* o = out + (size_t)q * record_words + T.out_offset;
    for (unsigned int c = threadIdx.x; c < T.leaf_len; c += blockDim.x)
        o[c] = T.col_stride ? T.data[(size_t)c * T.col_stride + leaf] : T.data[leaf * T.leaf_len + c];
    if (threadIdx.x == 0) o[T.leaf_len] = T.path_len;
    for (unsigned int w = threadIdx.x; w < 4 * T.path_len; w += blockDim.x) {
        unsigned int lvl = w >> 2;
        size_t off = 0;
        for (unsigned int k = 0; k < lvl; k++) off += ((size_t)4 << (T.log_leaves - k));
        size_t node = (leaf >> lvl) ^ 1;
        o[T.leaf_len + 1 + w] = T.digests[off + 4 * node + (w & 3)];
    }
}

// 1 / (n (x_j - 1)) on the LDE domain: the point-dependent factor of L_0(x) = Z_H(x) / (n (x - 1))
__global__ void l0inv_kernel(size_t N, gl_t n, const gl_t* __restrict__ domain, gl_t* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    out[j] = gl_inv(gl_mul(n, gl_sub(domain[j], 1)));
}
// domain points in leaf order: x_j = 7 * w_N^bitrev(j);  subgroup g^i
__global__ void domain_kernel(int logN, gl_t wN, gl_t shift, gl_t* __restrict__ out, int bitrev_order) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ((size_t)1 << logN)) return;
    uint32_t e = bitrev_order ? gl_bitrev((uint32_t)j, logN) : (uint32_t)j;
    out[j] = gl_mul(shift, gl_pow(wN, e));
}
