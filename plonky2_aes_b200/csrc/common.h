// Internal declarations shared by the translation units of libp2gpu.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include "gl64.cuh"

// number of kernels launched by this library since load (reported by bench.py as gpu_launches);
// incremented from every prover host thread, hence atomic
extern std::atomic<unsigned long long> g_p2g_launches;
#define P2G_COUNT_LAUNCH(k) (g_p2g_launches.fetch_add((k), std::memory_order_relaxed))

#define P2G_MAX_LOG_M 14   // largest sub-transform held in shared memory (2^14 * 8 B = 128 KB)

// ---- NTT (ntt.cu) -------------------------------------------------------------------------
// A plan = device tables for one transform shape.  The transform of size n = R * M runs as
// R * variants thread blocks per column: block (variant, q) folds the first log2(R) DIF stages,
// the coset/inverse scaling and the variant shift into its load phase
//     y_q[t] = sum_k x[t + k*M] * base_q^(t + kM)
// and then runs a size-M decimation-in-frequency transform entirely in shared memory.
enum { NTT_KIND_LDE = 0, NTT_KIND_INV = 1 };
// Two shapes, one table T[variant][q][t] = base_q^t (t < M; times 1/n for the inverse) -- V n words, linear in n.
// Direct (R <= 4): block (variant, q) computes y_q[t] = base_q^t * sum_k x[t + kM] * base_q^(kM) with the R per-block
// constants base_q^(kM) from C ([variants][R][R]): R - 1 products summed unreduced + one table multiplication per point.
// Pre-folded (R >= 8, prefold = 1): a first kernel (ntt_outer_kernel) multiplies x[t + kM] by shift^(kM) (C,
// [variants][R]) and runs the R-point transform over k for every t, leaving Z[q][t] where block q will read it; the
// size-M kernel does one multiplication per point on load.
struct NttPlan {
    int kind, log_n, log_m, log_r, log_variants, prefold;
    gl_t* T;    // [variants][R][M]
    gl_t* tw;   // per-pass compact twiddle tables (forward or inverse roots), tw_words entries
    gl_t* C;    // direct: [variants][R][R] constants base_q^(kM); pre-folded: [variants][R] input scale factors, then R/2 twiddles of the R-point transform
    int tw_words;
};
// largest transform: 2^21 points (R = 2^7 chunks of 2^14); the provers accept degree_bits <= 20
#define P2G_MAX_LOG_R 7
#define P2G_MAX_LOG_N 20
#define P2G_PREFOLD_MIN_LOG_R 3
// per-device kernel attributes (dynamic shared memory opt-in); called by p2g_ctx_create after cudaSetDevice
int ntt_init_device();
int ntt_plan_build(NttPlan* plan, int kind, int log_n, int rate_bits, cudaStream_t st);
void ntt_plan_free(NttPlan* plan);
// out_mode 0: block (variant,q) writes M contiguous words at out[col*out_stride + bitrev(variant)*n + q*M]
//             (bit-reversed order: the layout MerkleTree leaves and the quotient kernel use)
// out_mode 1: natural order, out[col*out_stride + f] (variants must be 1)
// blk_first / blk_count: only the output blocks (cosets in leaf order) [blk_first, blk_first + blk_count) are
// produced, into out[col*out_stride + (b - blk_first)*n ...]; blk_count = 0 means all
// scratch: pre-folded plans with out_mode 1 need ncols * n words of scratch (out_mode 0 stages Z in `out` itself);
// ntt_scratch_words tells how much
size_t ntt_scratch_words(const NttPlan* plan, int ncols, int out_mode);
int ntt_launch(const NttPlan* plan, const gl_t* in, size_t in_stride, gl_t* out, size_t out_stride,
               int ncols, int out_mode, cudaStream_t st, uint32_t blk_first = 0, uint32_t blk_count = 0, gl_t* scratch = nullptr);

// ---- Merkle (merkle.cu) -------------------------------------------------------------------
// Leaves are read either column-major (element c of leaf j at data[c*col_stride + j]) or
// row-major (data[j*leaf_len + c]).  digests: levels 0..L-1 concatenated, then the cap is
// written separately.  L = log2(num_leaves) - cap_height.
int merkle_build(const gl_t* data, int col_major, size_t col_stride, uint32_t leaf_len, uint32_t log_leaves,
                 uint32_t cap_height, gl_t* digests, gl_t* cap, cudaStream_t st);
size_t merkle_digest_words(uint32_t log_leaves, uint32_t cap_height);
size_t merkle_level_offset(uint32_t log_leaves, uint32_t level);  // in words

// poseidon microbenchmark (roofline denominator for the INT pipe): runs `iters` chained
// permutations per thread
int poseidon_bench_launch(gl_t* out, uint32_t nthreads_total, uint32_t iters, cudaStream_t st);
