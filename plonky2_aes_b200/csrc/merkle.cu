// Poseidon Merkle commitment for sm_100a: MerkleTree::new(leaves, cap_height) of the pinned
// plonky2 dependency (hash/merkle_tree.rs; /root/reference/Cargo.toml:12).
//
// Leaves are hashed straight out of the column-major, bit-reversed LDE buffer the NTT kernel
// wrote (thread j reads word j of every column: fully coalesced, no transposed copy exists).
// One thread = one leaf sponge; the block then folds log2(blockDim) tree levels through shared
// memory, writing every level's digests (needed later for Merkle paths).  The few levels above
// that are finished by a single block.
#include "common.h"
#include "poseidon.cuh"
#include <stdlib.h>

#ifndef POS_MINB
#define POS_MINB 3
#endif
#ifndef POS_BLOCK
#define POS_BLOCK 256          // threads per block of the thread-per-permutation kernels (power of two)
#endif
#define MERKLE_BLOCK POS_BLOCK
#define MERKLE_BLOCK_LOG (POS_BLOCK == 64 ? 6 : POS_BLOCK == 128 ? 7 : POS_BLOCK == 256 ? 8 : 9)

size_t merkle_level_offset(uint32_t log_leaves, uint32_t level) {
    size_t off = 0;
    for (uint32_t k = 0; k < level; k++) off += ((size_t)4 << (log_leaves - k));
    return off;
}
size_t merkle_digest_words(uint32_t log_leaves, uint32_t cap_height) {
    uint32_t L = log_leaves > cap_height ? log_leaves - cap_height : 0;
    size_t w = merkle_level_offset(log_leaves, L);
    return w ? w : 4;
}

// level pointer: levels < L live in `digests`, level L is the cap
__device__ __forceinline__ gl_t* level_ptr(gl_t* digests, gl_t* cap, uint32_t log_leaves, uint32_t L, uint32_t level) {
    if (level >= L) return cap;
    size_t off = 0;
    for (uint32_t k = 0; k < level; k++) off += ((size_t)4 << (log_leaves - k));
    return digests + off;
}

template <bool COL_MAJOR>
__global__ void __launch_bounds__(MERKLE_BLOCK, POS_MINB)
merkle_leaves_kernel(const gl_t* __restrict__ data, size_t col_stride, uint32_t leaf_len, uint32_t log_leaves,
                     uint32_t L, uint32_t levels_here, gl_t* __restrict__ digests, gl_t* __restrict__ cap) {
    __shared__ gl_t sh[MERKLE_BLOCK][4];
    const uint32_t tid = threadIdx.x;
    const size_t j = (size_t)blockIdx.x * blockDim.x + tid;
    const size_t num_leaves = (size_t)1 << log_leaves;
    const bool leaf_ok = j < num_leaves;
    gl_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    // One loop drives both phases so the permutation is inlined exactly once (the kernel was
    // instruction-fetch bound with two copies): iterations [0, absorbs) are the leaf sponge,
    // iterations [absorbs, absorbs + levels_here) fold tree levels through shared memory.
    const uint32_t absorbs = leaf_len <= 4 ? 0 : (leaf_len + 7) / 8;
    if (leaf_len <= 4 && leaf_ok) {                // hash_or_noop: the leaf is its own digest
#pragma unroll
        for (uint32_t c = 0; c < 4; c++)
            if (c < leaf_len) s[c] = COL_MAJOR ? __ldg(data + (size_t)c * col_stride + j) : __ldg(data + j * leaf_len + c);
    }
    if (absorbs == 0) {
        if (leaf_ok) {
            gl_t* d0 = level_ptr(digests, cap, log_leaves, L, 0);
#pragma unroll
            for (int i = 0; i < 4; i++) d0[4 * j + i] = s[i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) sh[tid][i] = s[i];
    }
    for (uint32_t it = 0; it < absorbs + levels_here; it++) {
        bool active;
        uint32_t lv = 0;
        if (it < absorbs) {
            active = leaf_ok;
            if (active) {
                const uint32_t c0 = it * 8;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint32_t c = c0 + i;
                    if (c < leaf_len)
                        s[i] = COL_MAJOR ? __ldg(data + (size_t)c * col_stride + j) : __ldg(data + j * leaf_len + c);
                }
            }
        } else {
            lv = it - absorbs + 1;
            __syncthreads();
            // a block padded to a full warp (fewer than 32 leaves) has more threads than nodes
            active = tid < (blockDim.x >> lv) && (((size_t)blockIdx.x * blockDim.x) >> lv) + tid < (num_leaves >> lv);
            if (active) {
#pragma unroll
                for (int i = 0; i < 4; i++) { s[i] = sh[2 * tid][i]; s[4 + i] = sh[2 * tid + 1][i]; s[8 + i] = 0; }
            }
            __syncthreads();
        }
        if (active) poseidon_permute_lazy(s);
        if (it + 1 == absorbs) {                   // leaf digest complete
#pragma unroll
            for (int i = 0; i < 4; i++) s[i] = gl_canon(s[i]);
            if (leaf_ok) {
                gl_t* d0 = level_ptr(digests, cap, log_leaves, L, 0);
#pragma unroll
                for (int i = 0; i < 4; i++) d0[4 * j + i] = s[i];
            }
#pragma unroll
            for (int i = 0; i < 4; i++) sh[tid][i] = s[i];
        } else if (it >= absorbs && active) {
            size_t node = ((size_t)blockIdx.x * blockDim.x >> lv) + tid;
            gl_t* d = level_ptr(digests, cap, log_leaves, L, lv);
#pragma unroll
            for (int i = 0; i < 4; i++) { gl_t v = gl_canon(s[i]); sh[tid][i] = v; d[4 * node + i] = v; }
        }
    }
}

// one grid-wide level, thread per node
__global__ void __launch_bounds__(POS_BLOCK, POS_MINB)
merkle_level_kernel(const gl_t* __restrict__ src, gl_t* __restrict__ dst, size_t cnt) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    gl_t l[4], r[4], o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { l[i] = src[8 * t + i]; r[i] = src[8 * t + 4 + i]; }
    poseidon_two_to_one(l, r, o);
#pragma unroll
    for (int i = 0; i < 4; i++) dst[4 * t + i] = o[i];
}

// one tree level, two nodes per warp (12 lanes each): the low-latency path for narrow levels
__global__ void __launch_bounds__(256)
merkle_level_coop_kernel(const gl_t* __restrict__ src, gl_t* __restrict__ dst, uint32_t cnt) {
    const uint32_t lane = threadIdx.x & 31, l = lane & 15, g = lane >> 4;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t node = warp * 2 + g;
#if defined(__CUDA_ARCH__)
    gl_t x = 0;
    if (node < cnt && l < 8) x = src[8 * (size_t)node + l];       // [left digest | right digest]
    x = poseidon_coop(x, l, g << 4, POSEIDON_RC_GLOBAL);
    if (node < cnt && l < 4) dst[4 * (size_t)node + l] = gl_canon(x);
#else
    (void)node; (void)l; (void)g; (void)src; (void)dst;
#endif
}
// Up to 7 consecutive narrow levels in one launch: a block of 1024 threads (64 groups of 16 lanes) owns 128 nodes of
// level lv_in and hashes its subtree upwards, one 12-lane permutation deep per level, children passed through
// shared memory; every level is also written to `digests` / `cap` (Merkle paths need them).  Replaces one launch
// per level for the levels of at most 4096 nodes: 9 launches -> 2 for the tree tops of a 2^18-leaf commitment.
#define TOP_NODES 128
__global__ void __launch_bounds__(1024)
merkle_top_coop_kernel(gl_t* __restrict__ digests, gl_t* __restrict__ cap, uint32_t log_leaves, uint32_t L,
                       uint32_t lv_in, uint32_t levels, uint32_t nodes_per_block) {
    __shared__ gl_t sh[2][TOP_NODES][4];
#if defined(__CUDA_ARCH__)
    const uint32_t lane = threadIdx.x & 31, l = lane & 15, g = threadIdx.x >> 4;      // group g of 16 lanes
    const gl_t* src = level_ptr(digests, cap, log_leaves, L, lv_in) + (size_t)blockIdx.x * nodes_per_block * 4;
    for (uint32_t i = threadIdx.x; i < nodes_per_block * 4; i += blockDim.x) sh[0][i >> 2][i & 3] = src[i];
    __syncthreads();
    uint32_t cnt = nodes_per_block >> 1;                    // parents of this block at the current level
    for (uint32_t p = 0; p < levels; p++, cnt >>= 1) {
        const uint32_t in = p & 1;
        if ((g & ~1u) < ((cnt + 1) & ~1u)) {                // whole warps stay together for the shuffles
            gl_t x = 0;
            if (g < cnt && l < 8) x = sh[in][2 * g + (l >> 2)][l & 3];
            x = poseidon_coop(x, l, (lane & 16), POSEIDON_RC_GLOBAL);
            if (g < cnt && l < 4) {
                const gl_t v = gl_canon(x);
                sh[in ^ 1][g][l] = v;
                gl_t* d = level_ptr(digests, cap, log_leaves, L, lv_in + p + 1);
                d[((size_t)blockIdx.x * cnt + g) * 4 + l] = v;
            }
        }
        __syncthreads();
    }
#else
    (void)digests; (void)cap; (void)log_leaves; (void)L; (void)lv_in; (void)levels; (void)nodes_per_block; (void)sh;
#endif
}
// row-major leaves hashed cooperatively (FRI layers with few leaves): two leaves per warp
__global__ void __launch_bounds__(256)
merkle_leaves_coop_kernel(const gl_t* __restrict__ data, uint32_t leaf_len, uint32_t cnt, gl_t* __restrict__ dst) {
    const uint32_t lane = threadIdx.x & 31, l = lane & 15, g = lane >> 4;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t leaf = warp * 2 + g;
#if defined(__CUDA_ARCH__)
    const bool ok = leaf < cnt;
    gl_t x = 0;
    for (uint32_t c0 = 0; c0 < leaf_len; c0 += 8) {
        if (ok && l < 8 && c0 + l < leaf_len) x = data[(size_t)leaf * leaf_len + c0 + l];
        x = poseidon_coop(x, l, g << 4, POSEIDON_RC_GLOBAL);
    }
    if (ok && l < 4) dst[4 * (size_t)leaf + l] = gl_canon(x);
#else
    (void)leaf; (void)l; (void)g; (void)data; (void)dst; (void)leaf_len;
#endif
}

int merkle_build(const gl_t* data, int col_major, size_t col_stride, uint32_t leaf_len, uint32_t log_leaves,
                 uint32_t cap_height, gl_t* digests, gl_t* cap, cudaStream_t st) {
    if (cap_height > log_leaves) return -2;
    const uint32_t L = log_leaves - cap_height;
    const size_t num_leaves = (size_t)1 << log_leaves;
    // levels (or leaf sets) this narrow use the 12-lane permutation; P2G_COOP_MAX: A/B knob
    static const uint32_t COOP_MAX = [] { const char* e = getenv("P2G_COOP_MAX"); return e ? (uint32_t)atoi(e) : 4096u; }();
    static const bool FUSE_TOP = [] { const char* e = getenv("P2G_FUSE_TOP"); return e ? atoi(e) != 0 : true; }();    // A/B knob
    uint32_t lv;
    if (!col_major && leaf_len > 4 && num_leaves <= COOP_MAX) {
        gl_t* d0 = L == 0 ? cap : digests;
        uint32_t warps = (uint32_t)((num_leaves + 1) / 2);
        merkle_leaves_coop_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(data, leaf_len, (uint32_t)num_leaves, d0);
        P2G_COUNT_LAUNCH(1);
        lv = 0;
    } else {
        uint32_t threads = num_leaves < MERKLE_BLOCK ? (uint32_t)num_leaves : MERKLE_BLOCK;
        if (threads < 32) threads = 32;   // keep full warps; surplus threads hash nothing
        uint32_t block_log = 0; while ((1u << block_log) < threads) block_log++;
        uint32_t real_log = log_leaves < block_log ? log_leaves : block_log;
        uint32_t levels_here = real_log < L ? real_log : L;
        {   // levels folded inside the leaf kernel: every folded level halves the busy threads of a block
            // that still pins its registers, so only one level is folded (measured on config 2:
            // 0/1/2/3 levels -> 86.2/86.6/85.9/85.4 proofs/s) and the per-level kernels finish the tree
            static const int cap_levels = [] { const char* e = getenv("P2G_MERKLE_BLOCK_LEVELS"); return e ? atoi(e) : 1; }();   // thread-safe init
            if ((int)levels_here > cap_levels) levels_here = (uint32_t)cap_levels;
        }
        uint32_t blocks = (uint32_t)((num_leaves + threads - 1) / threads);
        if (col_major)
            merkle_leaves_kernel<true><<<blocks, threads, 0, st>>>(data, col_stride, leaf_len, log_leaves, L, levels_here, digests, cap);
        else
            merkle_leaves_kernel<false><<<blocks, threads, 0, st>>>(data, col_stride, leaf_len, log_leaves, L, levels_here, digests, cap);
        P2G_COUNT_LAUNCH(1);
        lv = levels_here;
    }
    // remaining levels, one launch each: thread-per-node while the level is wide, 12-lane form below
    while (lv < L) {
        size_t cnt = (size_t)1 << (log_leaves - lv - 1);
        const gl_t* src = digests + merkle_level_offset(log_leaves, lv);
        gl_t* dst = (lv + 1 >= L) ? cap : digests + merkle_level_offset(log_leaves, lv + 1);
        if (cnt > COOP_MAX) {
            merkle_level_kernel<<<(uint32_t)((cnt + POS_BLOCK - 1) / POS_BLOCK), POS_BLOCK, 0, st>>>(src, dst, cnt);
            P2G_COUNT_LAUNCH(1);
            lv++;
        } else if (FUSE_TOP) {
            // up to 7 levels per launch: blocks of 128 nodes (fewer when the level is narrower)
            const uint32_t nodes = (uint32_t)(2 * cnt);
            const uint32_t per_block = nodes < TOP_NODES ? nodes : TOP_NODES;
            uint32_t k = 0; while ((per_block >> (k + 1)) >= 1 && k < L - lv) k++;       // levels below the block's root, capped at the tree's cap
            const uint32_t threads = per_block * 8 < 32 ? 32 : per_block * 8;              // 16 lanes per parent of the first level
            merkle_top_coop_kernel<<<nodes / per_block, threads, 0, st>>>(digests, cap, log_leaves, L, lv, k, per_block);
            P2G_COUNT_LAUNCH(1);
            lv += k;
        } else {
            uint32_t warps = (uint32_t)((cnt + 1) / 2);
            merkle_level_coop_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(src, dst, (uint32_t)cnt);
            P2G_COUNT_LAUNCH(1);
            lv++;
        }
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---- INT-pipe roofline microbenchmark: chained permutations, no memory traffic ------------
__global__ void __launch_bounds__(POS_BLOCK, POS_MINB)
poseidon_bench_kernel(gl_t* out, uint32_t iters) {
    gl_t s[12];
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = (gl_t)g * 12 + i;
    for (uint32_t k = 0; k < iters; k++) poseidon_permute_lazy(s);
    gl_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc ^= s[i];
    out[g] = acc;
}
int poseidon_bench_launch(gl_t* out, uint32_t nthreads_total, uint32_t iters, cudaStream_t st) {
    poseidon_bench_kernel<<<nthreads_total / POS_BLOCK, POS_BLOCK, 0, st>>>(out, iters);
    P2G_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
