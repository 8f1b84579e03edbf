// Internal context/batch structures of libp2gpu.so.
#pragma once
#include <sched.h>
#include <time.h>
#include <sys/prctl.h>
#include "common.h"
#include "../../include/p2gpu.h"
#include <map>
#include <string>
#include <vector>
#include <tuple>

struct p2g_batch {
    uint32_t ncols, log_n, rate_bits, cap_height;
    uint32_t blk_first, blk_log;   // leaf blocks held: [blk_first, blk_first + 2^blk_log) of the 2^rate_bits cosets
    gl_t* coeffs;    // [ncols][n]  natural coefficient order
    gl_t* lde;       // [ncols][N]  column-major, bit-reversed point order (== Merkle leaf order)
    gl_t* digests;   // tree levels 0..L-1
    gl_t* cap;       // [2^cap_height][4] device
    std::vector<gl_t> cap_host;
    size_t n() const { return (size_t)1 << log_n; }
    size_t N() const { return (size_t)1 << (log_n + blk_log); }          // leaves held by this batch
    uint32_t log_N() const { return log_n + blk_log; }
    uint32_t path_len() const { return log_N() - cap_height; }
};

struct ProveScratch;  // prover.cu

struct p2g_ctx {
    int device;
    cudaStream_t st;
    cudaMemPool_t pool;     // private pool: blocks freed on this stream are never handed to another
                            // context's stream, so proofs in flight on one GPU stay independent
    std::string err;
    std::map<std::tuple<int, int, int>, const NttPlan*> plans;   // pointers into the per-device plan cache (api.cu)
    gl_t* pinned; size_t pinned_words;     // small pinned staging buffer for D2H results
    bool timing;
    p2g_timings timings;
    p2g_transcript transcript;
    std::vector<gl_t> last_zs, last_quotient_chunks;
    bool keep_debug;
    float commit_ms[3];     // last commit: inverse NTT, coset LDE, Merkle (when timing is on)
    cudaEvent_t wait_ev;    // blocking-sync event: host threads sleep while they wait for the stream
    int wait_mode;          // 0 spin (cudaStreamSynchronize), 1 blocking-sync event, 2 poll + sched_yield, 3 poll + 15 us sleeps
    // P2G_CANARY=1 (debug): every ctx_alloc block gets a guard band behind it that ctx_free checks -- the
    // out-of-bounds-write detector used where compute-sanitizer is not available (tests/test_gpu_prove.py)
    bool canary;
    std::map<void*, size_t> canary_words;
    unsigned long long canary_failures, canary_checked;
};
#define P2G_CANARY_WORDS 64

// Host wait for everything queued on the context's stream.  A proof has ~10 such waits (Fiat-Shamir
// round trips) and a process keeps several proofs in flight on separate host threads, times one
// process per GPU, so the threads can outnumber the cores (8 ranks x 8 threads on a 32-core box).
// Default (P2G_SYNC unset or "yield"): poll cudaStreamQuery and sched_yield between polls -- as fast
// as spinning when cores are free (132 proofs/s) and the best mode when they are not (8 threads on
// 4 cores: 132 vs 127 spinning vs 126 blocking).  P2G_SYNC=spin: cudaStreamSynchronize.
// P2G_SYNC=block: sleep on a blocking-sync event (no CPU burnt while waiting).
static inline cudaError_t ctx_wait(p2g_ctx* ctx) {
    if (ctx->wait_mode == 0) return cudaStreamSynchronize(ctx->st);
    if (ctx->wait_mode == 2) {          // P2G_SYNC=yield: poll, giving the core away between polls
        cudaError_t e;
        while ((e = cudaStreamQuery(ctx->st)) == cudaErrorNotReady) sched_yield();
        return e;
    }
    if (ctx->wait_mode == 3) {          // P2G_SYNC=sleep: poll with short sleeps -- the core is free between polls
        static thread_local bool slack_set = false;
        if (!slack_set) { prctl(PR_SET_TIMERSLACK, 2000UL, 0, 0, 0); slack_set = true; }   // 2 us instead of the default 50 us
        cudaError_t e;
        while ((e = cudaStreamQuery(ctx->st)) == cudaErrorNotReady) { struct timespec ts = {0, 15000}; nanosleep(&ts, nullptr); }
        return e;
    }
    cudaError_t e = cudaEventRecord(ctx->wait_ev, ctx->st);
    return e != cudaSuccess ? e : cudaEventSynchronize(ctx->wait_ev);
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return P2G_E_CUDA; } } while (0)

int ctx_get_plan(p2g_ctx* ctx, int kind, int log_n, int rate_bits, const NttPlan** out);
int ctx_alloc(p2g_ctx* ctx, gl_t** p, size_t words);
void ctx_free(p2g_ctx* ctx, void* p);
// builds coefficients (optional inverse NTT), LDE and Merkle tree for device-resident columns
// blk_count = 0: the whole LDE domain.  Otherwise only leaf blocks [blk_first, blk_first + blk_count) are
// extended and hashed; cap_height then counts the levels below this shard's subtree roots.
int commit_dev(p2g_ctx* ctx, const gl_t* cols_dev, uint32_t ncols, uint32_t log_n, uint32_t rate_bits,
               uint32_t cap_height, bool from_values, p2g_batch** out, bool sync_cap,
               uint32_t blk_first = 0, uint32_t blk_count = 0);
