// C ABI of libp2gpu.so, part 1: context, PolynomialBatch commitment, Merkle/hash helpers.
// See include/p2gpu.h for the reference interfaces each entry point replaces.
#include "ctx.h"
#include "poseidon.cuh"
#include <string.h>
#include <stdio.h>

#include <mutex>

std::atomic<unsigned long long> g_p2g_launches{0};
extern "C" int32_t p2g_version(void) { return 2; }
extern "C" uint64_t p2g_launch_count(void) { return g_p2g_launches.load(std::memory_order_relaxed); }

// NTT plans (twiddle / fold tables) are per device, shared by every context of that device: bench.py
// keeps 8 contexts per GPU and the Rust side one per prover thread.  Freed when the last context
// of the device goes away.
static std::mutex g_plan_mu;
static std::map<std::tuple<int, int, int, int>, NttPlan> g_plans;     // (device, kind, log_n, rate_bits)
static std::map<int, int> g_ctx_per_device;

extern "C" int32_t p2g_ctx_create(int32_t device, p2g_ctx** out) {
    if (!out) return P2G_E_BADARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return P2G_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return P2G_E_CUDA;
    // function attributes belong to one device: set them for every device a context is created on
    if (ntt_init_device() != 0) return P2G_E_CUDA;
    p2g_ctx* ctx = new p2g_ctx();
    ctx->device = device; ctx->timing = false; ctx->keep_debug = false;
    ctx->st = nullptr; ctx->pool = nullptr; ctx->pinned = nullptr; ctx->wait_ev = nullptr;
    { const char* e = getenv("P2G_CANARY"); ctx->canary = e && atoi(e) != 0; }
    ctx->canary_failures = ctx->canary_checked = 0;
    memset(&ctx->timings, 0, sizeof(ctx->timings));
    memset(&ctx->transcript, 0, sizeof(ctx->transcript));
    bool ok = cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) == cudaSuccess;
    if (ok) {
        const char* mode = getenv("P2G_SYNC");
        ctx->wait_mode = !mode ? 2 : strcmp(mode, "block") == 0 ? 1 : strcmp(mode, "spin") == 0 ? 0 : strcmp(mode, "sleep") == 0 ? 3 : 2;
        ok = cudaEventCreateWithFlags(&ctx->wait_ev, cudaEventBlockingSync | cudaEventDisableTiming) == cudaSuccess;
    }
    if (ok) {
        cudaMemPoolProps props; memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        ok = cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess;
        if (ok) { uint64_t thr = UINT64_MAX; cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    }
    ctx->pinned_words = 1 << 20;
    if (ok) ok = cudaMallocHost(&ctx->pinned, ctx->pinned_words * sizeof(gl_t)) == cudaSuccess;
    if (!ok) {                                   // release whatever was created before the failure
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
        if (ctx->wait_ev) cudaEventDestroy(ctx->wait_ev);
        if (ctx->st) cudaStreamDestroy(ctx->st);
        delete ctx;
        return P2G_E_CUDA;
    }
    { std::lock_guard<std::mutex> lk(g_plan_mu); g_ctx_per_device[device]++; }
    *out = ctx;
    return P2G_OK;
}
extern "C" void p2g_ctx_destroy(p2g_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        if (--g_ctx_per_device[ctx->device] == 0) {         // last context of this device: drop its plans
            cudaDeviceSynchronize();
            for (auto it = g_plans.begin(); it != g_plans.end();) {
                if (std::get<0>(it->first) == ctx->device) { ntt_plan_free(&it->second); it = g_plans.erase(it); } else ++it;
            }
        }
    }
    cudaFreeHost(ctx->pinned);
    cudaEventDestroy(ctx->wait_ev);
    cudaMemPoolDestroy(ctx->pool);
    cudaStreamDestroy(ctx->st);
    delete ctx;
}
extern "C" const char* p2g_last_error(p2g_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" int32_t p2g_ctx_sync(p2g_ctx* ctx) { CU(ctx_wait(ctx)); return P2G_OK; }
extern "C" void* p2g_ctx_stream(p2g_ctx* ctx) { return (void*)ctx->st; }

int ctx_get_plan(p2g_ctx* ctx, int kind, int log_n, int rate_bits, const NttPlan** out) {
    auto lkey = std::make_tuple(kind, log_n, kind == NTT_KIND_LDE ? rate_bits : 0);
    auto lit = ctx->plans.find(lkey);                      // per-context cache of pointers: no lock on the hot path
    if (lit != ctx->plans.end()) { *out = lit->second; return P2G_OK; }
    if (log_n > P2G_MAX_LOG_N) { ctx->err = "transform larger than 2^20 points is not supported"; return P2G_E_BADARG; }
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto key = std::make_tuple(ctx->device, kind, log_n, kind == NTT_KIND_LDE ? rate_bits : 0);
    auto it = g_plans.find(key);
    if (it == g_plans.end()) {
        NttPlan p;
        int rc = ntt_plan_build(&p, kind, log_n, rate_bits, ctx->st);
        if (rc != 0) { ctx->err = rc == -2 ? "out of device memory for the NTT tables" : "ntt_plan_build failed"; return rc == -2 ? P2G_E_NOMEM : P2G_E_CUDA; }
        it = g_plans.emplace(key, p).first;
    }
    ctx->plans[lkey] = &it->second;                         // std::map nodes are address-stable
    *out = &it->second;
    return P2G_OK;
}
int ctx_alloc(p2g_ctx* ctx, gl_t** p, size_t words) {
    if (!words) words = 1;
    if (!ctx->canary) {
        CU(cudaMallocFromPoolAsync((void**)p, words * sizeof(gl_t), ctx->pool, ctx->st));
        return P2G_OK;
    }
    CU(cudaMallocFromPoolAsync((void**)p, (words + P2G_CANARY_WORDS) * sizeof(gl_t), ctx->pool, ctx->st));
    CU(cudaMemsetAsync(*p + words, 0xA5, P2G_CANARY_WORDS * sizeof(gl_t), ctx->st));
    ctx->canary_words[*p] = words;
    return P2G_OK;
}
void ctx_free(p2g_ctx* ctx, void* p) {
    if (!p) return;
    if (ctx->canary) {
        auto it = ctx->canary_words.find(p);
        if (it != ctx->canary_words.end()) {            // everything queued before this free has run once the copy is back
            gl_t* band = ctx->pinned + ctx->pinned_words - P2G_CANARY_WORDS;
            if (cudaMemcpyAsync(band, (gl_t*)p + it->second, P2G_CANARY_WORDS * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st) == cudaSuccess &&
                cudaStreamSynchronize(ctx->st) == cudaSuccess) {
                ctx->canary_checked++;
                for (int i = 0; i < P2G_CANARY_WORDS; i++)
                    if (band[i] != 0xA5A5A5A5A5A5A5A5ULL) {
                        ctx->canary_failures++;
                        fprintf(stderr, "[p2g] guard band behind a %zu-word block overwritten at word +%d\n", it->second, i);
                        break;
                    }
            }
            ctx->canary_words.erase(it);
        }
    }
    cudaFreeAsync(p, ctx->st);
}
extern "C" int32_t p2g_debug_canary(p2g_ctx* ctx, uint64_t* blocks_checked, uint64_t* failures) {
    if (!ctx || !ctx->canary) return P2G_E_BADARG;
    if (blocks_checked) *blocks_checked = ctx->canary_checked;
    if (failures) *failures = ctx->canary_failures;
    return P2G_OK;
}

int commit_dev(p2g_ctx* ctx, const gl_t* cols_dev, uint32_t ncols, uint32_t log_n, uint32_t rate_bits,
               uint32_t cap_height, bool from_values, p2g_batch** out, bool sync_cap, uint32_t blk_first, uint32_t blk_count) {
    if (blk_count == 0) { blk_first = 0; blk_count = 1u << rate_bits; }
    uint32_t blk_log = 0; while ((1u << blk_log) < blk_count) blk_log++;
    if ((1u << blk_log) != blk_count || blk_first % blk_count || blk_first + blk_count > (1u << rate_bits)) return P2G_E_BADARG;
    if (!ncols || log_n > P2G_MAX_LOG_N || log_n + rate_bits > 26 || cap_height > log_n + blk_log) {
        ctx->err = "unsupported commitment shape (log_n <= 20, cap_height <= log of the leaves held)"; return P2G_E_BADARG;
    }
    struct BatchGuard {                 // an early return releases the half-built batch
        p2g_ctx* ctx; p2g_batch* b;
        ~BatchGuard() { if (b) p2g_batch_free(ctx, b); }
    } guard{ctx, nullptr};
    p2g_batch* b = new p2g_batch();
    guard.b = b;
    b->ncols = ncols; b->log_n = log_n; b->rate_bits = rate_bits; b->cap_height = cap_height;
    b->blk_first = blk_first; b->blk_log = blk_log;
    b->coeffs = b->lde = b->digests = b->cap = nullptr;
    const size_t n = b->n(), N = b->N();
    int rc;
    if ((rc = ctx_alloc(ctx, &b->coeffs, (size_t)ncols * n))) return rc;
    if ((rc = ctx_alloc(ctx, &b->lde, (size_t)ncols * N))) return rc;
    if ((rc = ctx_alloc(ctx, &b->digests, merkle_digest_words(b->log_N(), cap_height)))) return rc;
    if ((rc = ctx_alloc(ctx, &b->cap, (size_t)4 << cap_height))) return rc;
    const NttPlan *inv, *lde;
    cudaEvent_t ev[4];
    const bool tim = ctx->timing;
    if (tim) { for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], ctx->st); }
    if (from_values) {
        if ((rc = ctx_get_plan(ctx, NTT_KIND_INV, log_n, 0, &inv))) return rc;
        // large transforms stage their outer stages in scratch (the caller's values must stay intact)
        gl_t* scratch = nullptr;
        const size_t sw = ntt_scratch_words(inv, (int)ncols, 1);
        if (sw && (rc = ctx_alloc(ctx, &scratch, sw))) return rc;
        const int lrc = ntt_launch(inv, cols_dev, n, b->coeffs, n, ncols, 1, ctx->st, 0, 0, scratch);
        if (scratch) ctx_free(ctx, scratch);
        if (lrc) { ctx->err = "intt launch"; return P2G_E_CUDA; }
    } else {
        CU(cudaMemcpyAsync(b->coeffs, cols_dev, (size_t)ncols * n * sizeof(gl_t), cudaMemcpyDeviceToDevice, ctx->st));
    }
    if (tim) cudaEventRecord(ev[1], ctx->st);
    if ((rc = ctx_get_plan(ctx, NTT_KIND_LDE, log_n, rate_bits, &lde))) return rc;
    if (ntt_launch(lde, b->coeffs, n, b->lde, N, ncols, 0, ctx->st, blk_first, blk_count)) { ctx->err = "lde launch"; return P2G_E_CUDA; }
    if (tim) cudaEventRecord(ev[2], ctx->st);
    if (merkle_build(b->lde, 1, N, ncols, b->log_N(), cap_height, b->digests, b->cap, ctx->st)) { ctx->err = "merkle launch"; return P2G_E_CUDA; }
    if (tim) {
        cudaEventRecord(ev[3], ctx->st); cudaEventSynchronize(ev[3]);
        for (int i = 0; i < 3; i++) cudaEventElapsedTime(&ctx->commit_ms[i], ev[i], ev[i + 1]);
        for (auto& e : ev) cudaEventDestroy(e);
    }
    b->cap_host.resize((size_t)4 << cap_height);
    if (sync_cap) {
        CU(cudaMemcpyAsync(ctx->pinned, b->cap, b->cap_host.size() * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
        CU(ctx_wait(ctx));
        memcpy(b->cap_host.data(), ctx->pinned, b->cap_host.size() * sizeof(gl_t));
    }
    guard.b = nullptr;
    *out = b;
    return P2G_OK;
}

static int32_t commit_any(p2g_ctx* ctx, const uint64_t* cols, bool host, bool from_values, uint32_t ncols, uint32_t log_n,
                          uint32_t rate_bits, uint32_t cap_height, p2g_batch** out, uint64_t* cap_out) {
    if (!ctx || !cols || !out) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    gl_t* tmp = nullptr;
    const gl_t* src = cols;
    int rc;
    if (host) {
        size_t words = (size_t)ncols << log_n;
        if ((rc = ctx_alloc(ctx, &tmp, words))) return rc;
        CU(cudaMemcpyAsync(tmp, cols, words * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st));
        src = tmp;
    }
    rc = commit_dev(ctx, src, ncols, log_n, rate_bits, cap_height, from_values, out, true);
    if (tmp) ctx_free(ctx, tmp);
    if (rc) return rc;
    if (cap_out) memcpy(cap_out, (*out)->cap_host.data(), (*out)->cap_host.size() * sizeof(gl_t));
    return P2G_OK;
}
extern "C" int32_t p2g_commit_from_values(p2g_ctx* ctx, const uint64_t* c, uint32_t ncols, uint32_t log_n, uint32_t rate_bits,
                                          uint32_t cap_height, p2g_batch** out, uint64_t* cap_out) {
    return commit_any(ctx, c, true, true, ncols, log_n, rate_bits, cap_height, out, cap_out);
}
extern "C" int32_t p2g_commit_from_coeffs(p2g_ctx* ctx, const uint64_t* c, uint32_t ncols, uint32_t log_n, uint32_t rate_bits,
                                          uint32_t cap_height, p2g_batch** out, uint64_t* cap_out) {
    return commit_any(ctx, c, true, false, ncols, log_n, rate_bits, cap_height, out, cap_out);
}
extern "C" int32_t p2g_commit_from_values_dev(p2g_ctx* ctx, const uint64_t* c, uint32_t ncols, uint32_t log_n, uint32_t rate_bits,
                                              uint32_t cap_height, p2g_batch** out, uint64_t* cap_out) {
    return commit_any(ctx, c, false, true, ncols, log_n, rate_bits, cap_height, out, cap_out);
}
extern "C" int32_t p2g_commit_from_coeffs_dev(p2g_ctx* ctx, const uint64_t* c, uint32_t ncols, uint32_t log_n, uint32_t rate_bits,
                                              uint32_t cap_height, p2g_batch** out, uint64_t* cap_out) {
    return commit_any(ctx, c, false, false, ncols, log_n, rate_bits, cap_height, out, cap_out);
}
extern "C" int32_t p2g_commit_blocks_from_values_dev(p2g_ctx* ctx, const uint64_t* cols_dev, uint32_t ncols, uint32_t log_n,
                                                     uint32_t rate_bits, uint32_t cap_height, uint32_t blk_first, uint32_t blk_count,
                                                     p2g_batch** out, uint64_t* cap_part_out) {
    if (!ctx || !cols_dev || !out || blk_count == 0) return P2G_E_BADARG;
    uint32_t blk_log = 0; while ((1u << blk_log) < blk_count) blk_log++;
    if (cap_height + blk_log < rate_bits) return P2G_E_BADARG;      // a shard must own whole cap entries
    CU(cudaSetDevice(ctx->device));
    int rc = commit_dev(ctx, cols_dev, ncols, log_n, rate_bits, cap_height + blk_log - rate_bits, true, out, true, blk_first, blk_count);
    if (rc) return rc;
    if (cap_part_out) memcpy(cap_part_out, (*out)->cap_host.data(), (*out)->cap_host.size() * sizeof(gl_t));
    return P2G_OK;
}
extern "C" int32_t p2g_last_commit_timings(p2g_ctx* ctx, float out[3]) {
    if (!ctx || !out) return P2G_E_BADARG;
    for (int i = 0; i < 3; i++) out[i] = ctx->commit_ms[i];
    return P2G_OK;
}
extern "C" int32_t p2g_batch_free(p2g_ctx* ctx, p2g_batch* b) {
    if (!ctx || !b) return P2G_E_BADARG;
    ctx_free(ctx, b->coeffs); ctx_free(ctx, b->lde); ctx_free(ctx, b->digests); ctx_free(ctx, b->cap);
    delete b;
    return P2G_OK;
}
extern "C" int32_t p2g_batch_get_coeffs(p2g_ctx* ctx, const p2g_batch* b, uint64_t* out) {
    CU(cudaMemcpyAsync(out, b->coeffs, (size_t)b->ncols * b->n() * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(ctx_wait(ctx));
    return P2G_OK;
}
extern "C" int32_t p2g_batch_get_lde(p2g_ctx* ctx, const p2g_batch* b, uint64_t* out) {
    CU(cudaMemcpyAsync(out, b->lde, (size_t)b->ncols * b->N() * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(ctx_wait(ctx));
    return P2G_OK;
}
extern "C" int32_t p2g_batch_get_level(p2g_ctx* ctx, const p2g_batch* b, uint32_t level, uint64_t* out) {
    uint32_t L = b->path_len();
    if (level > L) return P2G_E_BADARG;
    const gl_t* src = level == L ? b->cap : b->digests + merkle_level_offset(b->log_N(), level);
    size_t words = (size_t)4 << (b->log_N() - level);
    CU(cudaMemcpyAsync(out, src, words * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(ctx_wait(ctx));
    return P2G_OK;
}
extern "C" int32_t p2g_batch_open_leaf(p2g_ctx* ctx, const p2g_batch* b, uint64_t leaf, uint64_t* row_out, uint64_t* sib_out) {
    if (leaf >= b->N()) return P2G_E_BADARG;
    CU(cudaMemcpy2DAsync(row_out, sizeof(gl_t), b->lde + leaf, b->N() * sizeof(gl_t), sizeof(gl_t), b->ncols,
                         cudaMemcpyDeviceToHost, ctx->st));
    uint64_t idx = leaf;
    for (uint32_t k = 0; k < b->path_len(); k++) {
        const gl_t* lvl = b->digests + merkle_level_offset(b->log_N(), k);
        CU(cudaMemcpyAsync(sib_out + 4 * k, lvl + 4 * (idx ^ 1), 4 * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
        idx >>= 1;
    }
    CU(ctx_wait(ctx));
    return P2G_OK;
}

extern "C" int32_t p2g_merkle_cap(p2g_ctx* ctx, const uint64_t* leaves_host, uint32_t log_leaves, uint32_t leaf_len,
                                  uint32_t cap_height, uint64_t* cap_out, uint64_t* leaf_digests_out) {
    if (!ctx || !leaves_host || !cap_out || cap_height > log_leaves) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    size_t words = (size_t)leaf_len << log_leaves;
    gl_t *d_leaves, *d_dig, *d_cap; int rc;
    if ((rc = ctx_alloc(ctx, &d_leaves, words))) return rc;
    if ((rc = ctx_alloc(ctx, &d_dig, merkle_digest_words(log_leaves, cap_height)))) return rc;
    if ((rc = ctx_alloc(ctx, &d_cap, (size_t)4 << cap_height))) return rc;
    CU(cudaMemcpyAsync(d_leaves, leaves_host, words * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st));
    if (merkle_build(d_leaves, 0, 0, leaf_len, log_leaves, cap_height, d_dig, d_cap, ctx->st)) { ctx->err = "merkle launch"; return P2G_E_CUDA; }
    CU(cudaMemcpyAsync(cap_out, d_cap, ((size_t)4 << cap_height) * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    if (leaf_digests_out) {
        const gl_t* src = log_leaves == cap_height ? d_cap : d_dig;
        CU(cudaMemcpyAsync(leaf_digests_out, src, ((size_t)4 << log_leaves) * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    }
    CU(ctx_wait(ctx));
    ctx_free(ctx, d_leaves); ctx_free(ctx, d_dig); ctx_free(ctx, d_cap);
    return P2G_OK;
}

__global__ void hash_no_pad_many_kernel(const gl_t* __restrict__ in, uint32_t count, uint32_t len, gl_t* __restrict__ out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    gl_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    for (uint32_t c0 = 0; c0 < len; c0 += 8) {
#pragma unroll
        for (int i = 0; i < 8; i++) if (c0 + i < len) s[i] = in[(size_t)j * len + c0 + i];
        poseidon_permute_lazy(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[4 * (size_t)j + i] = gl_canon(s[i]);
}
extern "C" int32_t p2g_hash_no_pad_many(p2g_ctx* ctx, const uint64_t* in_host, uint32_t count, uint32_t len, uint64_t* out_host) {
    if (!ctx || !in_host || !out_host || !count) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    gl_t *d_in, *d_out; int rc;
    if ((rc = ctx_alloc(ctx, &d_in, (size_t)count * len))) return rc;
    if ((rc = ctx_alloc(ctx, &d_out, (size_t)count * 4))) return rc;
    CU(cudaMemcpyAsync(d_in, in_host, (size_t)count * len * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st));
    hash_no_pad_many_kernel<<<(count + 127) / 128, 128, 0, ctx->st>>>(d_in, count, len, d_out);
    P2G_COUNT_LAUNCH(1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_host, d_out, (size_t)count * 4 * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(ctx_wait(ctx));
    ctx_free(ctx, d_in); ctx_free(ctx, d_out);
    return P2G_OK;
}

extern "C" int32_t p2g_poseidon_peak(p2g_ctx* ctx, uint32_t iters, double* perms_per_sec) {
    if (!ctx || !perms_per_sec || !iters) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    const uint32_t nthreads = 148 * 2048;   // every SM full of resident warps
    gl_t* d_out; int rc;
    if ((rc = ctx_alloc(ctx, &d_out, nthreads))) return rc;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    poseidon_bench_launch(d_out, nthreads, 2, ctx->st);   // warm-up
    CU(cudaEventRecord(e0, ctx->st));
    if (poseidon_bench_launch(d_out, nthreads, iters, ctx->st)) { ctx->err = "poseidon bench launch"; return P2G_E_CUDA; }
    CU(cudaEventRecord(e1, ctx->st));
    CU(ctx_wait(ctx));
    float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
    *perms_per_sec = (double)nthreads * iters / (ms * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx_free(ctx, d_out);
    return P2G_OK;
}

// ---- device field arithmetic behind a test entry point ---------------------------------------
template <int K> struct Pow2Dispatch {
    static __device__ __forceinline__ gl_t run(gl_t x, int k) { return k == K ? gl_mul_pow2<K>(x) : Pow2Dispatch<K - 1>::run(x, k); }
};
template <> struct Pow2Dispatch<0> { static __device__ __forceinline__ gl_t run(gl_t x, int) { return x; } };

__global__ void field_ops_kernel(const gl_t* __restrict__ a, const gl_t* __restrict__ b, const gl_t* __restrict__ la,
                                 const gl_t* __restrict__ lb, size_t n, gl_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = gl_add(a[i], b[i]);
    out[n + i] = gl_sub(a[i], b[i]);
    out[2 * n + i] = gl_mul(a[i], b[i]);
    out[3 * n + i] = gl_canon(la[i]);
    out[4 * n + i] = gl_canon(gl_mul_lazy(la[i], lb[i]));
    out[5 * n + i] = Pow2Dispatch<95>::run(a[i], (int)(i % 96));
}

extern "C" int32_t p2g_field_ops(p2g_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* la, const uint64_t* lb,
                                 size_t n, uint64_t* out) {
    if (!ctx || !a || !b || !la || !lb || !out || !n) return P2G_E_BADARG;
    CU(cudaSetDevice(ctx->device));
    gl_t *d_in, *d_out; int rc;
    if ((rc = ctx_alloc(ctx, &d_in, 4 * n))) return rc;
    if ((rc = ctx_alloc(ctx, &d_out, 6 * n))) return rc;
    const uint64_t* src[4] = {a, b, la, lb};
    for (int k = 0; k < 4; k++) CU(cudaMemcpyAsync(d_in + k * n, src[k], n * sizeof(gl_t), cudaMemcpyHostToDevice, ctx->st));
    field_ops_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->st>>>(d_in, d_in + n, d_in + 2 * n, d_in + 3 * n, n, d_out);
    P2G_COUNT_LAUNCH(1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, 6 * n * sizeof(gl_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(ctx_wait(ctx));
    ctx_free(ctx, d_in); ctx_free(ctx, d_out);
    return P2G_OK;
}
