// Goldilocks NTT kernels for sm_100a: inverse NTT and coset low-degree extension behind
// PolynomialBatch::from_values / from_coeffs (plonky2 fri/oracle.rs, field/src/fft.rs of the
// dependency pinned at /root/reference/Cargo.toml:12).
//
// Design (B200-first, not a translation of the CPU radix-2 loop):
//  * one thread block owns one (column, coset, chunk) and keeps its 2^LOG_M points in shared
//    memory for the whole transform: HBM sees one read of the coefficients (L2-served for the
//    8 cosets of a column) and one coalesced write of the result;
//  * the coset shift 7*w_N^s, the zero padding of lde(), the 1/n of the inverse and the first
//    log2(R) decimation stages are folded into a single table multiply on load;
//  * decimation in frequency leaves the output in bit-reversed order, which IS the leaf order
//    of MerkleTree::new after reverse_index_bits_in_place, and coset s lands in the contiguous
//    block bitrev3(s) -- so the reference's transpose + bit-reverse pass disappears;
//  * radix-16 register passes (4 butterfly stages per shared-memory round trip), padded
//    shared-memory indexing (idx + idx/16) to stay bank-conflict free at every stride.
#include "common.h"
#include <vector>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smpad(uint32_t i) { return i + (i >> 4); }

// K butterfly stages on 2^K register-resident points.  DIF: (a,b) -> (a+b, (a-b)*w).
// log_b = log2 of the block size at the first stage of this pass.
// tw: this pass's twiddles in shared memory, tw[p] = w_B^p for p < B/2 (B = block size at the
// first stage of the pass); stage u needs w_{B>>u}^p = tw[p << u].
template <int K>
__device__ __forceinline__ void dif_pass(gl_t* sm, int log_m, int log_b, const gl_t* __restrict__ tw,
                                         uint32_t tid, uint32_t nthreads) {
    constexpr int E = 1 << K;
    const uint32_t M = 1u << log_m;
    const int log_s = log_b - K;           // stride between a thread's points
    const uint32_t S = 1u << log_s;
    for (uint32_t g = tid; g < (M >> K); g += nthreads) {
        uint32_t blk = g >> log_s, lowpos = g & (S - 1);
        uint32_t base = (blk << log_b) + lowpos;
        gl_t a[E];
#pragma unroll
        for (int i = 0; i < E; i++) a[i] = sm[smpad(base + ((uint32_t)i << log_s))];
#pragma unroll
        for (int u = 0; u < K; u++) {
            const int half = E >> (u + 1);            // in units of i
#pragma unroll
            for (int i = 0; i < E; i++) {
                if ((i & half) == 0) {
                    uint32_t p = lowpos + ((uint32_t)(i & (half - 1)) << log_s);  // position of the lower point in its block
                    gl_t w = tw[p << u];
                    gl_t x = a[i], y = a[i + half];
                    a[i] = gl_add(x, y);
                    a[i + half] = gl_mul(gl_sub(x, y), w);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < E; i++) sm[smpad(base + ((uint32_t)i << log_s))] = a[i];
    }
}

// DIF butterfly with the twiddle w_16^(+-E0) = 2^(156 E mod 192) as a compile-time shift; shifts of
// 96 and more are -2^(s-96), absorbed by swapping the operands of the subtraction.
template <int E0, bool INV>
__device__ __forceinline__ void bfly16(gl_t& x, gl_t& y) {
    const gl_t a = x, b = y;
    x = gl_add(a, b);
    constexpr int E = INV ? (16 - E0) % 16 : E0;     // inverse transform: w_16^-E
    constexpr int SH = (156 * E) % 192;
    if (SH == 0) y = gl_sub(a, b);
    else if (SH < 96) y = gl_mul_pow2<(SH > 0 && SH < 96) ? SH : 1>(gl_sub(a, b));
    else y = gl_mul_pow2<(SH >= 97) ? SH - 96 : 1>(gl_sub(b, a));
}
// The last pass (block size 16, points contiguous): every twiddle is one of w_16^0..7, fixed by the
// register index, so the 32 general multiplications and 32 twiddle loads of dif_pass<4> become
// 15 plain butterflies and 17 shift-multiplies.
template <bool INV>
__device__ __forceinline__ void dif_pass_last16(gl_t* sm, int log_m, uint32_t tid, uint32_t nthreads) {
    const uint32_t M = 1u << log_m;
    for (uint32_t g = tid; g < (M >> 4); g += nthreads) {
        const uint32_t base = g << 4;
        gl_t a[16];
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = sm[smpad(base + i)];
        bfly16<0, INV>(a[0], a[8]);  bfly16<1, INV>(a[1], a[9]);  bfly16<2, INV>(a[2], a[10]); bfly16<3, INV>(a[3], a[11]);
        bfly16<4, INV>(a[4], a[12]); bfly16<5, INV>(a[5], a[13]); bfly16<6, INV>(a[6], a[14]); bfly16<7, INV>(a[7], a[15]);
#pragma unroll
        for (int h = 0; h < 16; h += 8) {
            bfly16<0, INV>(a[h + 0], a[h + 4]); bfly16<2, INV>(a[h + 1], a[h + 5]);
            bfly16<4, INV>(a[h + 2], a[h + 6]); bfly16<6, INV>(a[h + 3], a[h + 7]);
        }
#pragma unroll
        for (int h = 0; h < 16; h += 4) { bfly16<0, INV>(a[h], a[h + 2]); bfly16<4, INV>(a[h + 1], a[h + 3]); }
#pragma unroll
        for (int h = 0; h < 16; h += 2) bfly16<0, INV>(a[h], a[h + 1]);
#pragma unroll
        for (int i = 0; i < 16; i++) sm[smpad(base + i)] = a[i];
    }
}

// Pre-folded plans, first kernel: for one column, one coset variant and 32 consecutive t it loads the R values
// x[t + kM] (32 consecutive words per k: coalesced), scales them by shift^(kM) and runs the R-point decimation in
// frequency over k in shared memory; entry q of the result, Z[q][t], is what block q of the size-M kernel folds in,
// and it is written where that block will read it (z + blk_local * n + q * M + t).  in == z is allowed for a
// single variant (every (col, t) reads its R words before it writes them).
__global__ void __launch_bounds__(256)
ntt_outer_kernel(const gl_t* in, size_t in_stride, gl_t* z, size_t z_stride, const gl_t* __restrict__ C,
                 int log_n, int log_m, int log_variants, int out_mode, uint32_t blk_first, int scale) {
    extern __shared__ gl_t sm[];
    const int log_r = log_n - log_m;
    const uint32_t R = 1u << log_r;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * 32 + lane;
    const uint32_t blk_local = blockIdx.y;
    const uint32_t variant = out_mode == 0 ? gl_bitrev(blk_first + blk_local, log_variants) : blk_local;
    const gl_t* x = in + (size_t)blockIdx.z * in_stride;
    const gl_t* Cv = C + (size_t)variant * R;
    const gl_t* twr = C + ((size_t)R << log_variants);
    for (uint32_t k = w; k < R; k += 8) {
        gl_t v = x[t + ((size_t)k << log_m)];
        if (scale && k) v = gl_mul(v, __ldg(Cv + k));
        sm[k * 33 + lane] = v;
    }
    __syncthreads();
    for (int s = 0; s < log_r; s++) {
        const int log_half = log_r - 1 - s;
        const uint32_t half = 1u << log_half;
        for (uint32_t b = w; b < (R >> 1); b += 8) {
            const uint32_t lowp = b & (half - 1);
            const uint32_t i = ((b >> log_half) << (log_half + 1)) | lowp, j = i + half;
            const gl_t a = sm[i * 33 + lane], c = sm[j * 33 + lane];
            sm[i * 33 + lane] = gl_add(a, c);
            const gl_t d = gl_sub(a, c);
            sm[j * 33 + lane] = lowp ? gl_mul(d, __ldg(twr + ((size_t)lowp << s))) : d;
        }
        __syncthreads();
    }
    gl_t* o = z + (size_t)blockIdx.z * z_stride + ((size_t)blk_local << log_n) + t;
    for (uint32_t q = w; q < R; q += 8) o[(size_t)q << log_m] = sm[q * 33 + lane];
}

__global__ void __launch_bounds__(512, 1)
ntt_dif_kernel(const gl_t* __restrict__ in, size_t in_stride, gl_t* __restrict__ out, size_t out_stride,
               const gl_t* __restrict__ T, const gl_t* __restrict__ tw, const gl_t* __restrict__ D,
               int log_n, int log_m, int log_variants, int out_mode, uint32_t blk_first, int inverse, uint32_t tw_skip, int prefold) {
    extern __shared__ gl_t sm[];
    const int log_r = log_n - log_m;
    const uint32_t M = 1u << log_m, R = 1u << log_r;
    // output block b (a contiguous run of n leaves) holds coset variant bitrev(b); a launch may cover
    // only blocks [blk_first, blk_first + gridDim.x / R) (multi-GPU coset sharding)
    const uint32_t blk_local = blockIdx.x >> log_r, q = blockIdx.x & (R - 1);
    const uint32_t variant = out_mode == 0 ? gl_bitrev(blk_first + blk_local, log_variants) : blk_local;
    const uint32_t col = blockIdx.y;
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    // pre-folded plans: `in` holds Z[q][t] at blk_local * n + q * M + t (possibly the output buffer itself: the block
    // has read its M words before anyone writes) and the table is base_q^t only
    const gl_t* x = in + (size_t)col * in_stride + (prefold ? ((size_t)blk_local << log_n) + ((size_t)q << log_m) : 0);
    const gl_t* Tq = T + ((size_t)(variant * R + q) << log_m);       // base_q^t, t < M (times 1/n for the inverse)
    const gl_t* Dq = D + ((size_t)(variant * R + q) << log_r);       // direct plans: base_q^(kM), k < R
    // per-pass twiddle tables live behind the data in shared memory (global/L2 twiddle loads were
    // the dominant stall of this kernel: long_scoreboard 3.0 per issue, profiles/)
    // tw_skip: leading table words left in global memory (the block-size-2^13 table of the two-blocks-
    // per-SM configuration, which would not fit twice)
    gl_t* tws = sm + M + (M >> 4) + 1;
    {
        uint32_t total = 0;
        int lb = log_m, rem0 = log_m & 3;
        if (rem0) { total += 1u << (lb - 1); lb -= rem0; }
        while (lb > 0) { total += 1u << (lb - 1); lb -= 4; }
        for (uint32_t i = tw_skip + tid; i < total; i += nth) tws[i - tw_skip] = __ldg(tw + i);
    }

    // load + fold: y[t] = sum_k x[t + kM] * T[k][t].  All loads of a batch of 8 points are issued
    // before any arithmetic so that one block per SM still keeps ~32 loads per thread in flight
    // (the kernel was stalled on these loads: long_scoreboard 3.0 per issue).
    if (prefold) {
        if ((M % (8 * nth)) == 0) {
            for (uint32_t t0 = tid; t0 < M; t0 += 8 * nth) {
                gl_t xv[8], tv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) { xv[u] = x[t0 + u * nth]; tv[u] = __ldg(Tq + t0 + u * nth); }
#pragma unroll
                for (int u = 0; u < 8; u++) sm[smpad(t0 + u * nth)] = gl_mul(xv[u], tv[u]);
            }
        } else {
            for (uint32_t t = tid; t < M; t += nth) sm[smpad(t)] = gl_mul(x[t], __ldg(Tq + t));
        }
    } else if (R <= 2 && (M % (8 * nth)) == 0) {
        // y[t] = base^t (x[t] + base^M x[t + M]): one table word per point, the per-block constant in a register
        const gl_t c1 = R == 2 ? __ldg(Dq + 1) : 0;
        for (uint32_t t0 = tid; t0 < M; t0 += 8 * nth) {
            gl_t xv[8][2], tv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t t = t0 + u * nth;
                xv[u][0] = __ldg(x + t); tv[u] = __ldg(Tq + t);
                if (R == 2) xv[u][1] = __ldg(x + t + M);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                gl_t acc = xv[u][0];
                if (R == 2) acc = gl_mad_lazy(xv[u][1], c1, acc);
                sm[smpad(t0 + u * nth)] = gl_mul(acc, tv[u]);
            }
        }
    } else if (R == 4 && (M % (4 * nth)) == 0) {
        // y[t] = base^t (x[t] + c1 x[t + M] + c2 x[t + 2M] + c3 x[t + 3M]), c_k = base^(kM): the three products are
        // summed unreduced (gl_mad3_lazy: one fold), then one multiplication by the table word -- 5 loads per point
        // instead of 8 and a table of V R M words instead of V R n (the load phase was waiting on L2: 52 % of the
        // kernel's samples for 32 % of its instructions)
        const gl_t c1 = __ldg(Dq + 1), c2 = __ldg(Dq + 2), c3 = __ldg(Dq + 3);
        // two half-batches of two points in flight: the loads of one are issued before the arithmetic of the other
        gl_t xa[2][4], ta[2], xb[2][4], tb[2];
#define NTT_LD4(XV, TV, T0) do {                                                          \
            _Pragma("unroll") for (int u = 0; u < 2; u++) {                               \
                const uint32_t t = (T0) + u * nth;                                        \
                TV[u] = __ldg(Tq + t);                                                    \
                _Pragma("unroll") for (int k = 0; k < 4; k++) XV[u][k] = __ldg(x + t + k * M); \
            } } while (0)
#define NTT_FOLD4(XV, TV, T0) do {                                                        \
            _Pragma("unroll") for (int u = 0; u < 2; u++)                                 \
                sm[smpad((T0) + u * nth)] = gl_mul(gl_mad3_lazy(XV[u][1], c1, XV[u][2], c2, XV[u][3], c3, XV[u][0]), TV[u]); \
            } while (0)
        NTT_LD4(xa, ta, tid);
        for (uint32_t t0 = tid; t0 < M; t0 += 4 * nth) {
            NTT_LD4(xb, tb, t0 + 2 * nth);
            NTT_FOLD4(xa, ta, t0);
            if (t0 + 4 * nth < M) NTT_LD4(xa, ta, t0 + 4 * nth);
            NTT_FOLD4(xb, tb, t0 + 2 * nth);
        }
#undef NTT_LD4
#undef NTT_FOLD4
    } else {
        for (uint32_t t = tid; t < M; t += nth) {
            gl_t acc = __ldg(x + t);
            for (uint32_t k = 1; k < R; k++)
                acc = gl_add(acc, gl_mul(__ldg(x + t + ((size_t)k << log_m)), __ldg(Dq + k)));
            sm[smpad(t)] = gl_mul(acc, __ldg(Tq + t));
        }
    }
    __syncthreads();

    int log_b = log_m;
    int rem = log_m & 3;
    const gl_t* twp = tws;      // always a shared-memory pointer, so the passes below compile to LDS
    if (rem) {
        if (tw_skip) {
            // two-blocks-per-SM configuration: the first (radix-2) pass reads its table from global memory
            dif_pass<1>(sm, log_m, log_b, tw, tid, nth);
        } else {
            if (rem == 1) dif_pass<1>(sm, log_m, log_b, twp, tid, nth);
            else if (rem == 2) dif_pass<2>(sm, log_m, log_b, twp, tid, nth);
            else dif_pass<3>(sm, log_m, log_b, twp, tid, nth);
            twp += 1u << (log_b - 1);
        }
        log_b -= rem; __syncthreads();
    }
    while (log_b > 0) {
        if (log_b == 4) { if (inverse) dif_pass_last16<true>(sm, log_m, tid, nth); else dif_pass_last16<false>(sm, log_m, tid, nth); }
        else dif_pass<4>(sm, log_m, log_b, twp, tid, nth);
        twp += 1u << (log_b - 1);
        log_b -= 4;
        __syncthreads();
    }

    gl_t* o = out + (size_t)col * out_stride;
    if (out_mode == 0) {
        size_t base = ((size_t)blk_local << log_n) + ((size_t)q << log_m);
        for (uint32_t m = tid; m < M; m += nth) o[base + m] = sm[smpad(m)];
    } else {
        uint32_t qr = gl_bitrev(q, log_r);
        for (uint32_t j = tid; j < M; j += nth) o[((size_t)j << log_r) + qr] = sm[smpad(gl_bitrev(j, log_m))];
    }
}

int ntt_init_device() {
    return cudaFuncSetAttribute(ntt_dif_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess ? 0 : -1;
}

// returns 0, -1 (CUDA error) or -2 (out of device memory)
int ntt_plan_build(NttPlan* plan, int kind, int log_n, int rate_bits, cudaStream_t st) {
    plan->kind = kind; plan->log_n = log_n; plan->T = plan->tw = plan->C = nullptr; plan->prefold = 0;
    if (log_n > P2G_MAX_LOG_N) return -1;
    plan->log_m = log_n < P2G_MAX_LOG_M ? log_n : P2G_MAX_LOG_M;
    if (log_n == 13) plan->log_m = 13;
    // n = 2^14, 2^15: 2^13-point chunks, two 256-thread blocks per SM, so one block's load / store phases
    // overlap the other's butterflies (LDE of the 135 x 2^15 wires 0.68 -> 0.65 ms); larger n keep
    // 2^14-point chunks.  P2G_NTT_LOG_M overrides (A/B knob).
    if (log_n == 14 || log_n == 15) plan->log_m = 13;
    // n >= 2^17 (pre-folded shape): 2^13-point chunks as well (64 x 2^17: LDE 1.72 -> 1.46 ms)
    if (log_n >= P2G_MAX_LOG_M + P2G_PREFOLD_MIN_LOG_R) plan->log_m = 13;
    {
        const char* e = getenv("P2G_NTT_LOG_M");
        if (e && atoi(e) >= 5 && atoi(e) <= P2G_MAX_LOG_M && log_n > atoi(e) && log_n - atoi(e) <= P2G_MAX_LOG_R) plan->log_m = atoi(e);
    }
    plan->log_r = log_n - plan->log_m;
    // R >= 8: pre-folded shape (table linear in n, one multiplication per point on load); P2G_NTT_PREFOLD=1 forces it
    // for every R >= 2 (tests), =0 forbids it below the table-size limit
    plan->prefold = plan->log_r >= P2G_PREFOLD_MIN_LOG_R;
    {
        const char* e = getenv("P2G_NTT_PREFOLD");
        if (e && atoi(e) == 1 && plan->log_r >= 1) plan->prefold = 1;
    }
    if (!plan->prefold && plan->log_r > 2) return -1;
    plan->log_variants = kind == NTT_KIND_LDE ? rate_bits : 0;
    const size_t n = (size_t)1 << log_n, M = (size_t)1 << plan->log_m, R = (size_t)1 << plan->log_r;
    const size_t V = (size_t)1 << plan->log_variants;
    std::vector<gl_t> T(V * R * M), tw(M / 2 ? M / 2 : 1), Cs;
    gl_t wn = gl_root_of_unity(log_n);
    gl_t wN = gl_root_of_unity(log_n + plan->log_variants);
    gl_t wm = gl_root_of_unity(plan->log_m);
    if (kind == NTT_KIND_INV) { wn = gl_inv(wn); wm = gl_inv(wm); }
    gl_t ninv = gl_inv((gl_t)1 << log_n);
    for (size_t v = 0; v < V; v++) {
        gl_t shift = kind == NTT_KIND_LDE ? gl_mul(7, gl_pow(wN, v)) : 1;
        for (size_t q = 0; q < R; q++) {
            gl_t base = gl_mul(shift, gl_pow(wn, gl_bitrev((uint32_t)q, plan->log_r)));
            gl_t x = kind == NTT_KIND_INV ? ninv : 1;
            gl_t* dst = T.data() + (v * R + q) * M;
            for (size_t j = 0; j < M; j++) { dst[j] = x; x = gl_mul(x, base); }
        }
    }
    if (!plan->prefold) {
        // direct plans: D[v][q][k] = base_q^(kM), the per-block constants of the load phase
        Cs.resize(V * R * R);
        for (size_t v = 0; v < V; v++) {
            gl_t shift = kind == NTT_KIND_LDE ? gl_mul(7, gl_pow(wN, v)) : 1;
            for (size_t q = 0; q < R; q++) {
                gl_t step = gl_pow(gl_mul(shift, gl_pow(wn, gl_bitrev((uint32_t)q, plan->log_r))), M), x = 1;
                for (size_t k = 0; k < R; k++) { Cs[(v * R + q) * R + k] = x; x = gl_mul(x, step); }
            }
        }
    }
    if (plan->prefold) {
        // C[v][k] = shift_v^(kM) (inverse plans: 1, the 1/n sits in T), then w_R^p for p < R/2 with w_R = w_n^M
        Cs.resize(V * R + R / 2 + 1);
        for (size_t v = 0; v < V; v++) {
            gl_t shift = kind == NTT_KIND_LDE ? gl_mul(7, gl_pow(wN, v)) : 1;
            gl_t step = gl_pow(shift, M), x = 1;
            for (size_t k = 0; k < R; k++) { Cs[v * R + k] = x; x = gl_mul(x, step); }
        }
        gl_t wr = gl_pow(wn, M), x = 1;
        for (size_t p2 = 0; p2 < R / 2; p2++) { Cs[V * R + p2] = x; x = gl_mul(x, wr); }
    }
    {   // per-pass compact twiddle tables, in pass order: w_B^p, p < B/2
        tw.clear();
        int lb = plan->log_m, rem0 = plan->log_m & 3;
        auto push = [&](int logb) {
            gl_t wb = gl_root_of_unity(logb);
            if (kind == NTT_KIND_INV) wb = gl_inv(wb);
            gl_t x = 1;
            for (size_t k = 0; k < ((size_t)1 << (logb - 1)); k++) { tw.push_back(x); x = gl_mul(x, wb); }
        };
        if (rem0) { push(lb); lb -= rem0; }
        while (lb > 0) { push(lb); lb -= 4; }
        if (tw.empty()) tw.push_back(1);
        plan->tw_words = (int)tw.size();
        (void)wm;
    }
    if (cudaMalloc(&plan->T, T.size() * sizeof(gl_t)) != cudaSuccess) { cudaGetLastError(); return -2; }
    if (cudaMalloc(&plan->tw, tw.size() * sizeof(gl_t)) != cudaSuccess) { cudaGetLastError(); ntt_plan_free(plan); return -2; }
    {
        if (cudaMalloc(&plan->C, Cs.size() * sizeof(gl_t)) != cudaSuccess) { cudaGetLastError(); ntt_plan_free(plan); return -2; }
        if (cudaMemcpyAsync(plan->C, Cs.data(), Cs.size() * sizeof(gl_t), cudaMemcpyHostToDevice, st) != cudaSuccess) { ntt_plan_free(plan); return -1; }
    }
    if (cudaMemcpyAsync(plan->T, T.data(), T.size() * sizeof(gl_t), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(plan->tw, tw.data(), tw.size() * sizeof(gl_t), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {             // host vectors die here
        ntt_plan_free(plan); return -1;
    }
    return 0;
}
void ntt_plan_free(NttPlan* plan) {
    if (plan->T) cudaFree(plan->T);
    if (plan->tw) cudaFree(plan->tw);
    if (plan->C) cudaFree(plan->C);
    plan->T = plan->tw = plan->C = nullptr;
}
size_t ntt_scratch_words(const NttPlan* plan, int ncols, int out_mode) {
    return plan->prefold && out_mode == 1 ? (size_t)ncols << plan->log_n : 0;
}

int ntt_launch(const NttPlan* plan, const gl_t* in, size_t in_stride, gl_t* out, size_t out_stride,
               int ncols, int out_mode, cudaStream_t st, uint32_t blk_first, uint32_t blk_count, gl_t* scratch) {
    const size_t M = (size_t)1 << plan->log_m;
    if (plan->prefold && out_mode == 1 && !scratch) return -1;
    // a 2^13-point chunk with a non-trivial fold (n > 2^13) runs two blocks per SM: 256 threads each and
    // the first pass's table (4096 words) stays in global memory
    const bool two_per_sm = plan->log_m == 13 && plan->log_r > 0;
    const uint32_t tw_skip = two_per_sm ? 4096u : 0u;
    size_t smem = (M + (M >> 4) + 1 + (size_t)plan->tw_words - tw_skip) * sizeof(gl_t);
    // (the > 48 KB dynamic shared memory opt-in is per device: ntt_init_device, called by p2g_ctx_create)
    uint32_t threads = (uint32_t)(M >> 4);
    if (threads < 32) threads = 32;
    if (threads > 512) threads = 512;
    if (two_per_sm) threads = 256;
    for (int c0 = 0; c0 < ncols; c0 += 65535) {
        int nc = ncols - c0 < 65535 ? ncols - c0 : 65535;
        if (blk_count == 0) blk_count = 1u << plan->log_variants;
        dim3 grid(blk_count << plan->log_r, nc);
        const gl_t* src = in + (size_t)c0 * in_stride;
        size_t src_stride = in_stride;
        if (plan->prefold) {
            // Z goes where the size-M blocks will read it: the output buffer itself (out_mode 0, same block layout)
            // or the scratch buffer (out_mode 1, whose output is scattered over the column)
            gl_t* z = out_mode == 0 ? out + (size_t)c0 * out_stride : scratch + ((size_t)c0 << plan->log_n);
            const size_t z_stride = out_mode == 0 ? out_stride : (size_t)1 << plan->log_n;
            dim3 og((unsigned)(M / 32), blk_count, nc);
            const size_t osm = ((size_t)33 << plan->log_r) * sizeof(gl_t);
            ntt_outer_kernel<<<og, 256, osm, st>>>(src, src_stride, z, z_stride, plan->C, plan->log_n, plan->log_m,
                                                  plan->log_variants, out_mode, blk_first, plan->kind == NTT_KIND_LDE ? 1 : 0);
            P2G_COUNT_LAUNCH(1);
            src = z; src_stride = z_stride;
        }
        ntt_dif_kernel<<<grid, threads, smem, st>>>(src, src_stride, out + (size_t)c0 * out_stride,
                                                    out_stride, plan->T, plan->tw, plan->C, plan->log_n, plan->log_m,
                                                    plan->log_variants, out_mode, blk_first, plan->kind == NTT_KIND_INV ? 1 : 0, tw_skip, plan->prefold);
        P2G_COUNT_LAUNCH(1);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
