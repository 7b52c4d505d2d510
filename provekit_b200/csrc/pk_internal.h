// provekit_b200/csrc/pk_internal.h — private state behind the opaque handles of include/pkwhir.h
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pkwhir.h"
#include "kernels.cuh"

struct pk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // Private stream-ordered memory pool.  With the device transcript the host enqueues a whole proof ahead of the GPU; on the
    // shared default pool the allocator then recycles blocks ACROSS the streams of the proofs in flight by inserting
    // inter-stream dependencies, which serialises them.  One pool per ctx (= per stream) keeps reuse stream-local.
    cudaMemPool_t pool = nullptr;
    bool shared_pool = false;
    std::string err;
    uint64_t launches = 0;
    // small fixed work areas
    void* d_partials = nullptr;   // REDUCE_MAX_BLOCKS * 3 field elements
    void* d_result = nullptr;     // 64 field elements
    uint64_t* h_result = nullptr; // pinned, 64 field elements
    unsigned long long* d_best = nullptr;
    // grow-on-demand work areas
    void* d_twiddles = nullptr;      // omega_M^e, e < M/2, Montgomery form
    void* d_twiddles_can = nullptr;  // the same values as canonical integers (coset twist of a canonical-output RS-encode)
    int twiddle_log_m = 0;
    // pass plans of the TMA-staged RS-encode per column length 2^L: stage split and per-pass twiddle slices
    struct NttPlan {
        int L = 0, npass = 0;
        int S[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
        void* tw[4] = {nullptr, nullptr, nullptr, nullptr};
    };
    std::vector<NttPlan> ntt_plans;
    void* d_scratch = nullptr;    // NTT scratch
    size_t scratch_elems = 0;
    void* d_tables = nullptr;     // tensor tables / small uploads
    size_t tables_elems = 0;
    void* d_small = nullptr;      // points / scalars / indexes staging (bytes)
    size_t small_bytes = 0;
    void* h_stage = nullptr;      // pinned staging for uploads/downloads
    size_t h_stage_bytes = 0;
    // sharded sumchecks (pk_shard_group_set): peer mailboxes and the lock-step sequence number of the exchange
    pk::ShardGroup shard = {};
    uint32_t shard_seq = 0;
    uint32_t* d_shard_status = nullptr;
    void* d_shard_gather = nullptr;  // 3 * SHARD_MAX_WORLD elements: the all-gathered payloads of pk_shard_barrier / pk_shard_allgather
    // optional per-kernel-class CUDA-event timing (pk_profile_begin/end)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    struct EvSpan { int cls; size_t a, b; };
    std::vector<EvSpan> ev_spans;
    size_t ev_used = 0;
};

struct pk_buf {
    void* d = nullptr;
    size_t n = 0;
    bool plain = false;  // cudaMalloc'd (IPC-exportable) instead of stream-ordered pool memory
};

struct pk_commitment {
    void* leaves = nullptr;  // L * w field elements, Montgomery
    void* nodes = nullptr;   // 2L field elements, canonical digests, heap order
    size_t L = 0, w = 0;
    int depth = 0;
    bool canonical_leaves = false;  // leaves hold canonical integers instead of Montgomery-form elements
    bool owns = true;               // false: a view over the caller's buffers (pk_commit_wrap), freed by the caller
};

namespace pk {
enum ProfClass { PROF_NTT = 0, PROF_MERKLE_LEAVES, PROF_MERKLE_UPPER, PROF_ZK_SUMCHECK, PROF_WHIR_SUMCHECK, PROF_WAVELET,
                 PROF_POW, PROF_OTHER, PROF_NCLASS };
// records a CUDA-event pair around the launches issued during its lifetime (only while profiling)
struct ProfScope {
    pk_ctx* ctx;
    size_t a = 0;
    int cls;
    bool on;
    ProfScope(pk_ctx* c, int cls_) : ctx(c), cls(cls_), on(c->profiling) {
        if (!on) return;
        a = next();
        cudaEventRecord(ctx->ev_pool[a], ctx->stream);
    }
    ~ProfScope() {
        if (!on) return;
        size_t b = next();
        cudaEventRecord(ctx->ev_pool[b], ctx->stream);
        ctx->ev_spans.push_back({cls, a, b});
    }
    size_t next() {
        if (ctx->ev_used == ctx->ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ctx->ev_pool.push_back(e);
        }
        return ctx->ev_used++;
    }
};
int set_err(pk_ctx* ctx, int code, const char* fmt, ...);
inline cudaError_t ctx_malloc(pk_ctx* ctx, void** p, size_t bytes) {
    return ctx->pool ? cudaMallocFromPoolAsync(p, bytes, ctx->pool, ctx->stream) : cudaMallocAsync(p, bytes, ctx->stream);
}
int ensure_twiddles(pk_ctx* ctx, int log_m);
int ensure_scratch(pk_ctx* ctx, size_t elems);
int ensure_tables(pk_ctx* ctx, size_t elems);
int ensure_small(pk_ctx* ctx, size_t bytes);
int ensure_stage(pk_ctx* ctx, size_t bytes);
cudaError_t init_kernel_attributes();
// device-pointer forms of the C-ABI helpers (pkwhir.cu): scalars, points and results stay on the device
int commit_batch_dev(pk_ctx* ctx, const void* const* coeffs, int batch, int log_n, int log_inv_rate, pk_commitment** out,
                     void* root_mont_dev);
int dev_eval_univariate(pk_ctx* ctx, const void* const* polys, int k, size_t n, const void* z_dev, void* out_dev);
int dev_eval_eq(pk_ctx* ctx, const void* point_dev, int mode, int n, const void* scale_dev, void* out);
int dev_mle_eval_prefix(pk_ctx* ctx, const void* const* evals, int k, int log_n, size_t n_prefix, const void* point_dev,
                        void* out_dev);
int dev_eval_eq_roots(pk_ctx* ctx, const uint64_t* exps_dev, const uint32_t* count_dev, size_t kmax, int log_d, int n,
                      const void* scalars_dev, void* out);
// skyscraper/core/src/pow.rs:61-82 threshold for `bits` of difficulty (incl. PROVER_BIAS), canonical limbs
void pow_threshold(double bits, uint64_t out[4]);
inline fr_arg to_arg(const uint64_t x[4]) {
    fr_arg a;
    for (int i = 0; i < 4; i++) {
        a.v[2 * i] = (uint32_t)x[i];
        a.v[2 * i + 1] = (uint32_t)(x[i] >> 32);
    }
    return a;
}
}  // namespace pk

// CUDA's current device is per host thread: every entry point binds the calling thread to the ctx's device
#define PK_BIND(ctx)                                   \
    do {                                               \
        if (ctx) cudaSetDevice((ctx)->device);         \
    } while (0)
#define PK_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return pk::set_err(ctx, e__ == cudaErrorMemoryAllocation ? PK_ERR_OOM : PK_ERR_CUDA,    \
                               "%s failed: %s", #call, cudaGetErrorString(e__));                    \
    } while (0)
#define PK_CHECK(ctx, cond, ...)                                                 \
    do {                                                                         \
        if (!(cond)) return pk::set_err(ctx, PK_ERR_INVALID_ARG, __VA_ARGS__);   \
    } while (0)
#define PK_TRY(expr)              \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != PK_OK) return rc__; \
    } while (0)
