// provekit_b200/csrc/pkwhir.cu — the C-ABI of include/pkwhir.h over the sm_100a kernels.
// No CPU fallback: every compute entry point launches CUDA kernels on ctx->stream or fails.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "host/fr_host.h"
#include "pk_internal.h"

namespace pk {

int set_err(pk_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

static int grow(pk_ctx* ctx, void** p, size_t* cur, size_t want, size_t unit) {
    if (*cur >= want) return PK_OK;
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cur = 0;
    PK_CUDA(ctx, cudaMalloc(p, want * unit));
    *cur = want;
    return PK_OK;
}
int ensure_scratch(pk_ctx* ctx, size_t elems) { return grow(ctx, &ctx->d_scratch, &ctx->scratch_elems, elems, 32); }
int ensure_tables(pk_ctx* ctx, size_t elems) { return grow(ctx, &ctx->d_tables, &ctx->tables_elems, elems, 32); }
int ensure_small(pk_ctx* ctx, size_t bytes) { return grow(ctx, &ctx->d_small, &ctx->small_bytes, bytes, 1); }
int ensure_stage(pk_ctx* ctx, size_t bytes) {
    if (ctx->h_stage_bytes >= bytes) return PK_OK;
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stage_bytes = 0;
    PK_CUDA(ctx, cudaMallocHost(&ctx->h_stage, bytes));
    ctx->h_stage_bytes = bytes;
    return PK_OK;
}

// twiddle table for the largest M seen so far; smaller domains use it strided (omega_{M/2} = omega_M^2)
int ensure_twiddles(pk_ctx* ctx, int log_m) {
    if (log_m < 1) log_m = 1;
    if (ctx->twiddle_log_m >= log_m) return PK_OK;
    if (log_m > 28) return set_err(ctx, PK_ERR_INVALID_ARG, "domain 2^%d exceeds the field's 2-adicity", log_m);
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_twiddles) cudaFree(ctx->d_twiddles);
    if (ctx->d_twiddles_can) cudaFree(ctx->d_twiddles_can);
    ctx->d_twiddles = ctx->d_twiddles_can = nullptr;
    ctx->twiddle_log_m = 0;
    PK_CUDA(ctx, cudaMalloc(&ctx->d_twiddles, ((size_t)32 << (log_m - 1))));
    PK_CUDA(ctx, cudaMalloc(&ctx->d_twiddles_can, ((size_t)32 << (log_m - 1))));
    uint32_t pow2[28][8];
    pkh::Fr g = pkh::root_of_unity(log_m);
    for (int b = 0; b < 28; b++) {
        std::memcpy(pow2[b], g.l, 32);
        g = pkh::sqr(g);
    }
    ctx->launches += launch_twiddle_table(ctx->stream, ctx->d_twiddles, log_m, &pow2[0][0]);
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->d_twiddles_can, ctx->d_twiddles, (size_t)32 << (log_m - 1), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->launches += launch_to_mont(ctx->stream, ctx->d_twiddles_can, (size_t)1 << (log_m - 1), false);
    PK_CUDA(ctx, cudaGetLastError());
    ctx->twiddle_log_m = log_m;
    return PK_OK;
}
// stage split of a 2^L-point column NTT into passes of at most NTT8_MAX_S stage bits (balanced: 17 -> 9 + 8) and the
// twiddle slice table of every pass (built once per L, kept for the life of the ctx)
static int ensure_ntt_plan(pk_ctx* ctx, int L, const pk_ctx::NttPlan** out) {
    for (const pk_ctx::NttPlan& p : ctx->ntt_plans)
        if (p.L == L) {
            *out = &p;
            return PK_OK;
        }
    PK_TRY(ensure_twiddles(ctx, L));
    pk_ctx::NttPlan plan;
    plan.L = L;
    plan.npass = (L + NTT8_MAX_S - 1) / NTT8_MAX_S;
    if (plan.npass > 4) return set_err(ctx, PK_ERR_INVALID_ARG, "rs_encode: column length 2^%d unsupported", L);
    int done = 0;
    for (int p = 0; p < plan.npass; p++) {
        const int S = (L - done + (plan.npass - p) - 1) / (plan.npass - p);
        plan.S[p] = S;
        plan.l[p] = L - done - S;
        PK_CUDA(ctx, cudaMalloc(&plan.tw[p], (size_t)32 << (plan.l[p] + S)));
        ctx->launches += launch_ntt_tile_twiddles(ctx->stream, plan.tw[p], ctx->d_twiddles, ctx->twiddle_log_m, L, plan.l[p], S);
        done += S;
    }
    PK_CUDA(ctx, cudaGetLastError());
    ctx->ntt_plans.reserve(16);  // pointers handed out stay valid
    ctx->ntt_plans.push_back(plan);
    *out = &ctx->ntt_plans.back();
    return PK_OK;
}

static int fetch_result(pk_ctx* ctx, uint64_t* out, int n_elems) {
    PK_CUDA(ctx, cudaGetLastError());
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, (size_t)n_elems * 32, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(out, ctx->h_result, (size_t)n_elems * 32);
    return PK_OK;
}

}  // namespace pk

using namespace pk;

extern "C" {

const char* pk_version(void) { return "pkwhir 0.1 (sm_100a)"; }

// Host threads that wait for the device (every challenge round trip does) either spin or sleep.  Spinning is the CUDA
// default and the fastest for one proof; with many proofs in flight on many GPUs of one box the spinning threads can
// outnumber the cores.  Must be called before the device's primary context exists (before any other CUDA call on it).
int pk_set_blocking_sync(int device, int on) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return PK_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return PK_ERR_CUDA;
    cudaError_t e = cudaSetDeviceFlags(on ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return PK_ERR_CUDA;  // the primary context is already active: too late for this process
    }
    return PK_OK;
}
int pk_ctx_create(int device, pk_ctx** out) {
    if (!out) return PK_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return PK_ERR_NO_DEVICE;
    pk_ctx* ctx = new pk_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&ctx->d_partials, REDUCE_AREA_BYTES) != cudaSuccess ||
        cudaMemset(ctx->d_partials, 0, REDUCE_AREA_BYTES) != cudaSuccess ||
        cudaMalloc(&ctx->d_result, 64 * 32) != cudaSuccess || cudaMallocHost((void**)&ctx->h_result, 64 * 32) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_best, 8) != cudaSuccess || init_kernel_attributes() != cudaSuccess) {
        pk_ctx_destroy(ctx);
        return PK_ERR_CUDA;
    }
    {   // private pool; freed blocks stay cached in it: allocation becomes a pointer bump
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&ctx->pool, &props) != cudaSuccess) {
            cudaGetLastError();
            ctx->pool = nullptr;
            cudaDeviceGetDefaultMemPool(&ctx->pool, device);  // fall back to the shared pool (never destroyed by us)
            ctx->shared_pool = true;
        }
        if (ctx->pool) {
            uint64_t thr = ~0ULL;
            cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    *out = ctx;
    return PK_OK;
}
void pk_ctx_destroy(pk_ctx* ctx) {
    PK_BIND(ctx);
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_partials);
    cudaFree(ctx->d_result);
    if (ctx->d_shard_status) cudaFree(ctx->d_shard_status);
    if (ctx->d_shard_gather) cudaFree(ctx->d_shard_gather);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    cudaFree(ctx->d_best);
    cudaFree(ctx->d_twiddles);
    cudaFree(ctx->d_twiddles_can);
    for (pk_ctx::NttPlan& p : ctx->ntt_plans)
        for (int i = 0; i < 4; i++) cudaFree(p.tw[i]);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->d_tables);
    cudaFree(ctx->d_small);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->pool && !ctx->shared_pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
const char* pk_last_error(const pk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
uint64_t pk_launch_count(const pk_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* pk_ctx_stream(pk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int pk_ctx_sync(pk_ctx* ctx) {
    PK_BIND(ctx);
    if (!ctx) return PK_ERR_INVALID_ARG;
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

// ---- buffers -----------------------------------------------------------------------------------
int pk_buf_alloc(pk_ctx* ctx, size_t n, pk_buf** out) {
    PK_BIND(ctx);
    if (!ctx || !out) return PK_ERR_INVALID_ARG;
    *out = nullptr;
    pk_buf* b = new pk_buf();
    b->n = n;
    cudaError_t e = ctx_malloc(ctx, &b->d, n ? n * 32 : 32);
    if (e != cudaSuccess) {
        delete b;
        return set_err(ctx, PK_ERR_OOM, "cudaMalloc(%zu elems): %s", n, cudaGetErrorString(e));
    }
    *out = b;
    return PK_OK;
}
void pk_buf_free(pk_ctx* ctx, pk_buf* b) {
    PK_BIND(ctx);
    if (!b) return;
    if (b->plain) {
        if (ctx && ctx->stream) cudaStreamSynchronize(ctx->stream);
        cudaFree(b->d);
        delete b;
        return;
    }
    if (ctx && ctx->stream)
        cudaFreeAsync(b->d, ctx->stream);  // stream-ordered: no host synchronisation
    else
        cudaFree(b->d);
    delete b;
}
size_t pk_buf_len(const pk_buf* b) { return b ? b->n : 0; }
void* pk_buf_device_ptr(pk_buf* b) { return b ? b->d : nullptr; }
int pk_buf_upload(pk_ctx* ctx, pk_buf* dst, size_t off, const uint64_t* host, size_t n) {
    PK_BIND(ctx);
    PK_CHECK(ctx, dst && host && off + n <= dst->n, "pk_buf_upload: range out of bounds");
    PK_CUDA(ctx, cudaMemcpyAsync((char*)dst->d + off * 32, host, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}
int pk_buf_download(pk_ctx* ctx, const pk_buf* src, size_t off, uint64_t* host, size_t n) {
    PK_BIND(ctx);
    PK_CHECK(ctx, src && host && off + n <= src->n, "pk_buf_download: range out of bounds");
    PK_CUDA(ctx, cudaMemcpyAsync(host, (const char*)src->d + off * 32, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}
int pk_buf_copy(pk_ctx* ctx, pk_buf* dst, size_t doff, const pk_buf* src, size_t soff, size_t n) {
    PK_BIND(ctx);
    PK_CHECK(ctx, dst && src && doff + n <= dst->n && soff + n <= src->n, "pk_buf_copy: range out of bounds");
    PK_CUDA(ctx, cudaMemcpyAsync((char*)dst->d + doff * 32, (const char*)src->d + soff * 32, n * 32, cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    return PK_OK;
}
int pk_buf_zero(pk_ctx* ctx, pk_buf* dst, size_t off, size_t n) {
    PK_BIND(ctx);
    PK_CHECK(ctx, dst && off + n <= dst->n, "pk_buf_zero: range out of bounds");
    PK_CUDA(ctx, cudaMemsetAsync((char*)dst->d + off * 32, 0, n * 32, ctx->stream));
    return PK_OK;
}

// ---- mask generator (device twin of F::rand(&mut thread_rng()), zk_utils.rs:13-22) ---------------------
int pk_rng_fill(pk_ctx* ctx, pk_buf* dst, size_t off, size_t n, const uint8_t seed[32], uint32_t stream) {
    PK_BIND(ctx);
    PK_CHECK(ctx, dst && seed && off + n <= dst->n, "pk_rng_fill: range out of bounds");
    uint32_t key[8];
    for (int i = 0; i < 8; i++)
        key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
                 ((uint32_t)seed[4 * i + 3] << 24);
    ctx->launches += pk::launch_rng_fill(ctx->stream, (char*)dst->d + off * 32, n, key, stream);
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}

// ---- multi-GPU plumbing: IPC-exportable buffers --------------------------------------------------
int pk_buf_alloc_shared(pk_ctx* ctx, size_t n, pk_buf** out) {
    PK_BIND(ctx);
    if (!ctx || !out) return PK_ERR_INVALID_ARG;
    *out = nullptr;
    pk_buf* b = new pk_buf();
    b->n = n;
    b->plain = true;  // cudaMalloc (not the stream-ordered pool): legacy CUDA IPC can export it
    cudaError_t e = cudaMalloc(&b->d, n ? n * 32 : 32);
    if (e != cudaSuccess) {
        delete b;
        return set_err(ctx, PK_ERR_OOM, "cudaMalloc(%zu elems): %s", n, cudaGetErrorString(e));
    }
    *out = b;
    return PK_OK;
}
int pk_ipc_export(pk_ctx* ctx, const pk_buf* buf, uint8_t handle_out[64]) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && buf && handle_out && buf->plain, "ipc_export: needs a buffer from pk_buf_alloc_shared");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    PK_CUDA(ctx, cudaIpcGetMemHandle(&h, buf->d));
    std::memcpy(handle_out, &h, 64);
    return PK_OK;
}
int pk_ipc_open(pk_ctx* ctx, const uint8_t handle[64], void** dptr_out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && handle && dptr_out, "ipc_open: null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    PK_CUDA(ctx, cudaIpcOpenMemHandle(dptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return PK_OK;
}
int pk_ipc_close(pk_ctx* ctx, void* dptr) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && dptr, "ipc_close: null argument");
    PK_CUDA(ctx, cudaIpcCloseMemHandle(dptr));
    return PK_OK;
}

// ---- Skyscraper ----------------------------------------------------------------------------------
int pk_skyscraper_compress_many(pk_ctx* ctx, const uint8_t* messages, uint8_t* hashes, size_t n) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && (n == 0 || (messages && hashes)), "compress_many: null buffer");
    if (n == 0) return PK_OK;
    // bounded device staging so arbitrarily large host batches stream through
    const size_t CH = (size_t)1 << 22;
    PK_TRY(ensure_scratch(ctx, 3 * (n < CH ? n : CH)));
    for (size_t off = 0; off < n; off += CH) {
        size_t m = n - off < CH ? n - off : CH;
        char* d_msgs = (char*)ctx->d_scratch;
        char* d_out = d_msgs + m * 64;
        PK_CUDA(ctx, cudaMemcpyAsync(d_msgs, messages + off * 64, m * 64, cudaMemcpyHostToDevice, ctx->stream));
        ctx->launches += launch_compress_many(ctx->stream, d_msgs, d_out, m);
        PK_CUDA(ctx, cudaGetLastError());
        PK_CUDA(ctx, cudaMemcpyAsync(hashes + off * 32, d_out, m * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}
int pk_skyscraper_compress_many_dev(pk_ctx* ctx, const pk_buf* messages, pk_buf* hashes, size_t n) {
    PK_BIND(ctx);
    PK_CHECK(ctx, messages && hashes && messages->n >= 2 * n && hashes->n >= n, "compress_many_dev: buffer too small");
    ctx->launches += launch_compress_many(ctx->stream, messages->d, hashes->d, n);
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}

// skyscraper/core/src/pow.rs:61-82
static void f64_to_u256(double f, uint64_t out[4]) {
    uint64_t bits;
    std::memcpy(&bits, &f, 8);
    bool sign = bits >> 63;
    int exp_bits = (int)((bits >> 52) & 0x7ff);
    uint64_t frac = bits & ((1ULL << 52) - 1);
    int exp = exp_bits == 0 ? -1022 : exp_bits - 1023;
    uint64_t sig = exp_bits == 0 ? frac : frac + (1ULL << 52);
    std::memset(out, 0, 32);
    if (sign) return;
    if (exp > 256) {
        std::memset(out, 0xff, 32);
        return;
    }
    int shift = exp - 52;
    if (shift < 0) {
        out[0] = (uint64_t)std::round(f);
    } else {
        int limb = shift / 64, sh = shift % 64;
        out[limb] = sig << sh;
        if (sh != 0 && limb < 3) out[limb + 1] = sig >> (64 - sh);
    }
}
}  // extern "C"
void pk::pow_threshold(double bits, uint64_t out[4]) {
    double modulus = (double)pkh::P[3] * std::ldexp(1.0, 192);  // pow.rs:19
    f64_to_u256(std::exp2(-(bits + 0.01)) * modulus, out);      // PROVER_BIAS, pow.rs:6,37
}
extern "C" {
int pk_pow_solve(pk_ctx* ctx, const uint64_t challenge[4], double bits, uint64_t* nonce) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && challenge && nonce, "pow_solve: null argument");
    // provekit/common/src/skyscraper/pow.rs:16: assert!((0.0..60.0).contains(&bits))
    PK_CHECK(ctx, bits >= 0.0 && bits < 60.0, "bits must be smaller than 60");
    if (bits == 0.0) {  // pow.rs:34-36
        *nonce = 0;
        return PK_OK;
    }
    uint64_t thr[4];
    pk::pow_threshold(bits, thr);
    unsigned long long init = ~0ULL;
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->d_best, &init, 8, cudaMemcpyHostToDevice, ctx->stream));
    // one launch covers ~8x the expected nonce (miss probability e^-8); blocks above the first hit exit immediately and
    // nonces in flight are abandoned between round pairs (k_pow_scan), so the cost tracks the winning nonce, not the
    // chunk size.  Small difficulties (the blinding WHIR's 4..11 bits) get small chunks: a 2^18 floor made every one of
    // them hash 2^18 nonces, all resident at once, before the early exit could act.
    uint64_t chunk = (uint64_t)1 << 12;
    double want = std::exp2(bits + 3.0);
    while ((double)chunk < want && chunk < ((uint64_t)1 << 30)) chunk <<= 1;
    for (uint64_t base = 0;; base += chunk) {
        {
            ProfScope ps(ctx, PROF_POW);
            ctx->launches += launch_pow_scan(ctx->stream, to_arg(challenge), to_arg(thr), base, chunk, ctx->d_best);
        }
        PK_CUDA(ctx, cudaGetLastError());
        PK_CUDA(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_best, 8, cudaMemcpyDeviceToHost, ctx->stream));
        PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->h_result[0] != ~0ULL) {
            *nonce = ctx->h_result[0];
            return PK_OK;
        }
        if (base + chunk < base) return set_err(ctx, PK_ERR_INTERNAL, "pow_solve: nonce space exhausted");
    }
}

// ---- wavelet -----------------------------------------------------------------------------------
static int wavelet(pk_ctx* ctx, pk_buf* buf, int log_n, bool inverse) {
    PK_CHECK(ctx, buf && log_n >= 0 && log_n < 40 && ((size_t)1 << log_n) <= buf->n, "wavelet: 2^%d exceeds buffer", log_n);
    {
        ProfScope ps(ctx, PROF_WAVELET);
        ctx->launches += launch_wavelet(ctx->stream, buf->d, log_n, inverse);
    }
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}
int pk_evals_to_coeffs(pk_ctx* ctx, pk_buf* buf, int log_n) {
    PK_BIND(ctx); return wavelet(ctx, buf, log_n, true); }
int pk_coeffs_to_evals(pk_ctx* ctx, pk_buf* buf, int log_n) {
    PK_BIND(ctx); return wavelet(ctx, buf, log_n, false); }

// ---- commit ------------------------------------------------------------------------------------
// canonical: emit the codeword as canonical integers (the Merkle leaf hash's input form) instead of Montgomery elements
static int rs_encode_raw(pk_ctx* ctx, const void* coeffs, int log_n, int log_inv_rate, int fold, void* leaves,
                         size_t leaf_stride, size_t col_offset, bool canonical = false) {
    PK_CHECK(ctx, fold == 4, "only FoldingFactor::Constant(4) is supported (r1cs-compiler/src/whir_r1cs.rs:44)");
    PK_CHECK(ctx, log_n >= fold && log_inv_rate >= 0 && log_n + log_inv_rate <= 28, "rs_encode: bad sizes");
    const int L = log_n - fold, logM = L + log_inv_rate;
    PK_TRY(ensure_twiddles(ctx, logM));
    PK_TRY(ensure_scratch(ctx, (size_t)1 << (log_n + log_inv_rate)));
    // Kernel choice (measured on B200, profiles/r02_ntt_kernel_choice.jsonl): the TMA-staged radix-8 kernel runs the columns of
    // the witness commitment (L = 17, 18: two passes of full 2^8..2^9-point tiles; within 2 % of the radix-2 kernel's time at
    // 58 % of its DRAM traffic).  Shorter columns (32-128-thread CTAs) and a third pass (L >= 19) are 5-15 % slower than the
    // register-staged radix-2 kernel, which keeps them.  PK_NTT_KERNEL=r8 | radix2 forces one kernel (tests, A/B runs).
    const char* force = std::getenv("PK_NTT_KERNEL");
    const bool legacy_env = std::getenv("PK_NTT_LEGACY") && std::getenv("PK_NTT_LEGACY")[0] == '1';
    bool use_r8 = L >= 17 && L <= 2 * NTT8_MAX_S;
    if (force && force[0] == 'r' && force[1] == '8') use_r8 = L >= NTT8_MIN_L;
    if ((force && force[0] == 'r' && force[1] == 'a') || legacy_env) use_r8 = false;
    if (!use_r8) {  // tiny and very long transforms: the register-staged radix-2 kernel
        ProfScope ps(ctx, PROF_NTT);
        ctx->launches += launch_rs_encode(ctx->stream, coeffs, log_n, log_inv_rate, fold, leaves, leaf_stride, col_offset,
                                          ctx->d_scratch, ctx->d_twiddles, ctx->twiddle_log_m, canonical ? ctx->d_twiddles_can : nullptr);
        PK_CUDA(ctx, cudaGetLastError());
        return PK_OK;
    }
    const pk_ctx::NttPlan* plan;
    PK_TRY(ensure_ntt_plan(ctx, L, &plan));
    // two scratch halves ping-pong between passes (a pass may not overwrite the layout it is still reading)
    const size_t cw = (size_t)1 << (log_n + log_inv_rate);
    if (plan->npass > 2) PK_TRY(ensure_scratch(ctx, 2 * cw));
    ProfScope ps(ctx, PROF_NTT);
    for (int p = 0; p < plan->npass; p++) {
        NttR8 P = {};
        P.first = p == 0;
        P.last = p == plan->npass - 1;
        char* ping = (char*)ctx->d_scratch + (size_t)(p & 1) * cw * 32;
        char* pong = (char*)ctx->d_scratch + (size_t)((p + 1) & 1) * cw * 32;
        P.in = (const fr*)(P.first ? coeffs : (const void*)pong);
        P.out = (fr*)(P.last ? leaves : (void*)ping);
        P.tw = (const fr*)plan->tw[p];
        P.Wtwist = (const fr*)(canonical ? ctx->d_twiddles_can : ctx->d_twiddles);
        P.tbl_shift = ctx->twiddle_log_m - (logM < 1 ? 1 : logM);
        P.L = L;
        P.l = plan->l[p];
        P.S = plan->S[p];
        P.logE = log_inv_rate;
        P.logM = logM < 1 ? 1 : logM;
        P.canonical = canonical ? 1 : 0;
        P.l_next = P.last ? 0 : plan->l[p + 1];
        P.S_next = P.last ? 0 : plan->S[p + 1];
        P.leaf_stride = leaf_stride;
        P.col_offset = col_offset;
        ctx->launches += launch_ntt_r8_pass(ctx->stream, P);
    }
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}
int pk_rs_encode(pk_ctx* ctx, const pk_buf* coeffs, int log_n, int log_inv_rate, int fold, pk_buf* leaves,
                 size_t leaf_stride, size_t col_offset) {
    PK_BIND(ctx);
    PK_CHECK(ctx, coeffs && leaves && log_n >= 0 && log_n < 40 && coeffs->n >= ((size_t)1 << log_n), "rs_encode: coeffs too small");
    size_t rows = (size_t)1 << (log_n + log_inv_rate - fold);
    PK_CHECK(ctx, col_offset + ((size_t)1 << fold) <= leaf_stride && leaves->n >= rows * leaf_stride, "rs_encode: leaves too small");
    return rs_encode_raw(ctx, coeffs->d, log_n, log_inv_rate, fold, leaves->d, leaf_stride, col_offset);
}
int pk_rs_encode_sharded(pk_ctx* ctx, const pk_buf* coeffs, int log_n, int log_inv_rate, int fold, int col_first, int n_cols,
                         void* const* peer_leaves, int n_peers, size_t leaf_stride, size_t col_offset) {
    PK_BIND(ctx);
    PK_CHECK(ctx, coeffs && peer_leaves && fold == 4, "rs_encode_sharded: bad arguments");
    PK_CHECK(ctx, log_n >= fold && log_inv_rate >= 0 && log_n + log_inv_rate <= 28 && coeffs->n >= ((size_t)1 << log_n),
             "rs_encode_sharded: bad sizes");
    int nc_log = 0;
    while ((1 << nc_log) < n_cols) nc_log++;
    PK_CHECK(ctx, (1 << nc_log) == n_cols && n_cols >= 1 && n_cols <= 16 && col_first % n_cols == 0 && col_first + n_cols <= 16,
             "rs_encode_sharded: columns must be an aligned power-of-two block of the 16");
    int logM = log_n - fold + log_inv_rate;
    PK_CHECK(ctx, n_peers >= 1 && n_peers <= 8 && (n_peers & (n_peers - 1)) == 0 && (1 << logM) >= n_peers,
             "rs_encode_sharded: peers must be a power of two <= 8");
    for (int i = 0; i < n_peers; i++) PK_CHECK(ctx, peer_leaves[i], "rs_encode_sharded: null peer pointer");
    PK_CHECK(ctx, col_offset + 16 <= leaf_stride, "rs_encode_sharded: leaf_stride too small");
    PK_TRY(ensure_twiddles(ctx, logM));
    PK_TRY(ensure_scratch(ctx, ((size_t)n_cols << logM)));
    {
        ProfScope ps(ctx, PROF_NTT);
        ctx->launches += launch_rs_encode_cols(ctx->stream, coeffs->d, log_n, log_inv_rate, fold, col_first, nc_log, peer_leaves,
                                               n_peers, leaf_stride, col_offset, ctx->d_scratch, ctx->d_twiddles, ctx->twiddle_log_m);
    }
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}
// top of a sharded tree: `n` canonical sub-tree roots (n = power of two <= 8, rank order) -> canonical root
int pk_merkle_combine_roots(pk_ctx* ctx, const uint64_t* roots, int n, uint64_t root_out[4]) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && roots && root_out && n >= 1 && n <= 8 && (n & (n - 1)) == 0, "combine_roots: bad arguments");
    uint64_t cur[8 * 4], nxt[4 * 4];
    std::memcpy(cur, roots, (size_t)n * 32);
    for (int m = n; m > 1; m >>= 1) {
        PK_TRY(pk_skyscraper_compress_many(ctx, (const uint8_t*)cur, (uint8_t*)nxt, (size_t)m / 2));
        std::memcpy(cur, nxt, (size_t)(m / 2) * 32);
    }
    std::memcpy(root_out, cur, 32);
    return PK_OK;
}
int pk_merkle_build(pk_ctx* ctx, const pk_buf* leaves, size_t L, size_t w, pk_buf* nodes) {
    PK_BIND(ctx);
    PK_CHECK(ctx, leaves && nodes, "merkle_build: null buffer");
    if (w == 0) return set_err(ctx, PK_ERR_EMPTY_INPUT, "IncorrectInputLength(0)");
    PK_CHECK(ctx, L >= 2 && (L & (L - 1)) == 0, "merkle_build: leaf count must be a power of two >= 2");
    PK_CHECK(ctx, leaves->n >= L * w && nodes->n >= 2 * L, "merkle_build: buffer too small");
    {
        ProfScope ps(ctx, PROF_MERKLE_LEAVES);
        ctx->launches += launch_merkle_leaves(ctx->stream, leaves->d, L, w, nodes->d, false);
    }
    {
        ProfScope ps(ctx, PROF_MERKLE_UPPER);
        ctx->launches += launch_merkle_upper(ctx->stream, L, nodes->d);
    }
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}
}  // extern "C"
namespace pk {
// commit_batch without a host round trip: the root is left on the device, in Montgomery form, at root_mont_dev
int commit_batch_dev(pk_ctx* ctx, const void* const* coeffs, int batch, int log_n, int log_inv_rate, pk_commitment** out,
                     void* root_mont_dev) {
    const int fold = 4;
    pk_commitment* c = new pk_commitment();
    c->w = ((size_t)batch) << fold;
    c->L = (size_t)1 << (log_n + log_inv_rate - fold);
    c->depth = log_n + log_inv_rate - fold;
    if (ctx_malloc(ctx, &c->leaves, c->L * c->w * 32) != cudaSuccess || ctx_malloc(ctx, &c->nodes, 2 * c->L * 32) != cudaSuccess) {
        pk_commit_free(ctx, c);
        return set_err(ctx, PK_ERR_OOM, "commit_batch: out of device memory");
    }
    // the commitment keeps its codeword as canonical integers: the leaf hash consumes that form directly
    c->canonical_leaves = !(std::getenv("PK_LEAVES_MONTGOMERY") && std::getenv("PK_LEAVES_MONTGOMERY")[0] == '1');
    for (int b = 0; b < batch; b++) {
        int rc = rs_encode_raw(ctx, coeffs[b], log_n, log_inv_rate, fold, c->leaves, c->w, (size_t)b << fold, c->canonical_leaves);
        if (rc != PK_OK) {
            pk_commit_free(ctx, c);
            return rc;
        }
    }
    {
        ProfScope ps(ctx, PROF_MERKLE_LEAVES);
        ctx->launches += launch_merkle_leaves(ctx->stream, c->leaves, c->L, c->w, c->nodes, c->canonical_leaves);
    }
    {
        ProfScope ps(ctx, PROF_MERKLE_UPPER);
        ctx->launches += launch_merkle_upper(ctx->stream, c->L, c->nodes, root_mont_dev);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pk_commit_free(ctx, c);
        return set_err(ctx, PK_ERR_CUDA, "commit_batch: %s", cudaGetErrorString(e));
    }
    *out = c;
    return PK_OK;
}
}  // namespace pk
extern "C" {
int pk_commit_batch(pk_ctx* ctx, const pk_buf* const* coeffs, int batch, int log_n, int log_inv_rate, int fold,
                    pk_commitment** out, uint64_t root_out[4]) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && coeffs && out && root_out && batch >= 1 && batch <= 8, "commit_batch: bad arguments");
    PK_CHECK(ctx, fold == 4 && log_n >= fold && log_n + log_inv_rate - fold >= 1, "commit_batch: bad sizes");
    const void* ptrs[8];
    for (int b = 0; b < batch; b++) {
        PK_CHECK(ctx, coeffs[b] && coeffs[b]->n >= ((size_t)1 << log_n), "commit_batch: poly %d too small", b);
        ptrs[b] = coeffs[b]->d;
    }
    pk_commitment* c = nullptr;
    PK_TRY(commit_batch_dev(ctx, ptrs, batch, log_n, log_inv_rate, &c, ctx->d_result));
    int rc = fetch_result(ctx, root_out, 1);
    if (rc != PK_OK) {
        pk_commit_free(ctx, c);
        return rc;
    }
    *out = c;
    return PK_OK;
}
void pk_commit_free(pk_ctx* ctx, pk_commitment* c) {
    PK_BIND(ctx);
    if (!c) return;
    if (!c->owns) {
        delete c;
        return;
    }
    if (ctx && ctx->stream) {
        if (c->leaves) cudaFreeAsync(c->leaves, ctx->stream);
        if (c->nodes) cudaFreeAsync(c->nodes, ctx->stream);
    } else {
        cudaFree(c->leaves);
        cudaFree(c->nodes);
    }
    delete c;
}
size_t pk_commit_num_leaves(const pk_commitment* c) { return c ? c->L : 0; }
size_t pk_commit_leaf_width(const pk_commitment* c) { return c ? c->w : 0; }

// A shard's leaf block and sub-tree (pk_rs_encode_sharded + pk_merkle_build) as a commitment, so that pk_commit_open* runs on
// it with LOCAL row indexes; non-owning.
int pk_commit_wrap(pk_ctx* ctx, const pk_buf* leaves, const pk_buf* nodes, size_t num_leaves, size_t leaf_width, pk_commitment** out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && leaves && nodes && out, "commit_wrap: null argument");
    PK_CHECK(ctx, num_leaves >= 1 && (num_leaves & (num_leaves - 1)) == 0 && leaf_width >= 1, "commit_wrap: leaf count must be a power of two");
    PK_CHECK(ctx, leaves->n >= num_leaves * leaf_width && nodes->n >= 2 * num_leaves, "commit_wrap: buffers too small");
    pk_commitment* c = new pk_commitment();
    c->leaves = leaves->d;
    c->nodes = nodes->d;
    c->L = num_leaves;
    c->w = leaf_width;
    c->depth = 0;
    while (((size_t)1 << c->depth) < num_leaves) c->depth++;
    c->owns = false;
    *out = c;
    return PK_OK;
}
// STIR answers and the UNCOMPRESSED authentication paths: paths_out[q][level], level 0 = the sibling leaf digest, level
// depth-1 = the sibling below the root (canonical digests).  The building block of pk_commit_open; a sharded opening gathers
// these per owner rank, appends the levels above the sub-trees and compresses once (pk_multipath_build).
int pk_commit_open_paths(pk_ctx* ctx, const pk_commitment* c, const uint64_t* sorted_idx, size_t n_idx, uint64_t* leaves_out,
                         uint64_t* paths_out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && c && (n_idx == 0 || (sorted_idx && leaves_out && (paths_out || c->depth == 0))), "commit_open: null argument");
    if (n_idx == 0) return PK_OK;
    for (size_t i = 0; i < n_idx; i++) {
        PK_CHECK(ctx, sorted_idx[i] < c->L, "commit_open: index %llu out of range", (unsigned long long)sorted_idx[i]);
        PK_CHECK(ctx, i == 0 || sorted_idx[i] > sorted_idx[i - 1], "commit_open: indexes must be strictly increasing");
    }
    const int depth = c->depth;
    size_t n_rows = n_idx * c->w, n_path = n_idx * (size_t)depth;
    PK_TRY(ensure_small(ctx, n_idx * 8));
    PK_TRY(ensure_tables(ctx, n_rows + n_path));
    PK_TRY(ensure_stage(ctx, (n_rows + n_path) * 32));
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->d_small, sorted_idx, n_idx * 8, cudaMemcpyHostToDevice, ctx->stream));
    char* d_rows = (char*)ctx->d_tables;
    char* d_path = d_rows + n_rows * 32;
    ctx->launches += launch_gather_rows(ctx->stream, c->leaves, c->w, (const uint64_t*)ctx->d_small, n_idx, d_rows);
    if (c->canonical_leaves) ctx->launches += launch_to_mont(ctx->stream, d_rows, n_rows, true);  // the ABI returns field elements
    if (depth > 0) ctx->launches += launch_gather_paths(ctx->stream, c->nodes, c->L, (const uint64_t*)ctx->d_small, n_idx, depth, d_path);
    PK_CUDA(ctx, cudaGetLastError());
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage, d_rows, (n_rows + n_path) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(leaves_out, ctx->h_stage, n_rows * 32);
    if (n_path) std::memcpy(paths_out, (char*)ctx->h_stage + n_rows * 32, n_path * 32);
    return PK_OK;
}
// ark MultiPath from uncompressed paths (paths[q][level], level 0 = sibling leaf digest): sibling leaf digest + auth path
// root->leaf (excluding the leaf level), prefix-compressed against the previous path
// (recursive-verifier/app/circuit/mt.go:36-50, utilities.go:71-82).  Host-side bookkeeping, no device work.
int pk_multipath_build(pk_ctx* ctx, const uint64_t* paths, size_t n_idx, int depth, uint64_t* sibling_out, uint64_t* prefix_len_out,
                       uint64_t* suffix_out, uint64_t* suffix_len_out, size_t suffix_cap) {
    PK_CHECK(ctx, depth >= 1 && depth < 64, "multipath_build: a tree with a single leaf has no path");
    PK_CHECK(ctx, n_idx == 0 || (paths && sibling_out && prefix_len_out && suffix_out && suffix_len_out), "multipath_build: null argument");
    size_t used = 0;
    const int plen = depth - 1;
    for (size_t q = 0; q < n_idx; q++) {
        const uint64_t* cur = paths + q * depth * 4;
        std::memcpy(sibling_out + 4 * q, cur, 32);
        int k = 0;
        if (q > 0) {
            const uint64_t* prev = paths + (q - 1) * depth * 4;
            // root->leaf element j is gathered level depth-1-j
            while (k < plen && std::memcmp(prev + (size_t)(depth - 1 - k) * 4, cur + (size_t)(depth - 1 - k) * 4, 32) == 0) k++;
        }
        prefix_len_out[q] = (uint64_t)k;
        suffix_len_out[q] = (uint64_t)(plen - k);
        PK_CHECK(ctx, used + (size_t)(plen - k) <= suffix_cap, "multipath_build: suffix_out too small");
        for (int j = k; j < plen; j++) std::memcpy(suffix_out + 4 * (used++), cur + (size_t)(depth - 1 - j) * 4, 32);
    }
    return PK_OK;
}
int pk_commit_open(pk_ctx* ctx, const pk_commitment* c, const uint64_t* sorted_idx, size_t n_idx, uint64_t* leaves_out,
                   uint64_t* sibling_out, uint64_t* prefix_len_out, uint64_t* suffix_out, uint64_t* suffix_len_out,
                   size_t suffix_cap) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && c && (n_idx == 0 || (sorted_idx && leaves_out && sibling_out && prefix_len_out && suffix_out && suffix_len_out)),
             "commit_open: null argument");
    if (n_idx == 0) return PK_OK;
    std::vector<uint64_t> paths(n_idx * (size_t)c->depth * 4);
    PK_TRY(pk_commit_open_paths(ctx, c, sorted_idx, n_idx, leaves_out, paths.data()));
    return pk_multipath_build(ctx, paths.data(), n_idx, c->depth, sibling_out, prefix_len_out, suffix_out, suffix_len_out, suffix_cap);
}

// ---- univariate / multilinear helpers ------------------------------------------------------------
// The dev_* forms take every scalar, point and count as a DEVICE pointer and leave their result on the device: the
// protocol flow (host/flow.cpp) chains them without a host round trip.  The C-ABI entry points stage their host
// arguments into ctx->d_small and call the same code.
}  // extern "C"
namespace pk {
static int split_bits(int n) { return n / 2; }  // low-table bits

// stages `bytes` of host data into ctx->d_small at `off` (the caller sized d_small); pageable source: ordered by the stream
static int stage_small(pk_ctx* ctx, size_t off, const void* host, size_t bytes) {
    if (bytes) PK_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_small + off, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PK_OK;
}
// (hi, lo) tensor tables of ONE point with nv variables into ctx->d_tables
static int point_tables(pk_ctx* ctx, const void* point_dev, int mode, int nv, bool eq_mode, const void* scale_dev, char** t_hi,
                        char** t_lo, int* lo_bits) {
    int lo = split_bits(nv), hi = nv - lo;
    PK_TRY(ensure_tables(ctx, ((size_t)1 << lo) + ((size_t)1 << hi)));
    *t_hi = (char*)ctx->d_tables;
    *t_lo = *t_hi + ((size_t)32 << hi);
    *lo_bits = lo;
    TensorSrc src = {};
    src.mode = mode;
    ctx->launches += launch_tensor_tables(ctx->stream, point_dev, 1, nv, hi, lo, scale_dev, eq_mode, *t_hi, *t_lo, src);
    return PK_OK;
}
int dev_eval_univariate(pk_ctx* ctx, const void* const* polys, int k, size_t n, const void* z_dev, void* out_dev) {
    int nv = 0;
    while (((size_t)1 << nv) < n) nv++;
    char *t_hi, *t_lo;
    int lo;
    // point (z^(2^(nv-1)), ..., z^2, z): power tables instead of eq tables
    ProfScope ps(ctx, PROF_OTHER);
    PK_TRY(point_tables(ctx, z_dev, TENSOR_PTS_UNIVARIATE, nv, false, nullptr, &t_hi, &t_lo, &lo));
    int rc = launch_multi_tensor_dot(ctx->stream, polys, k, n, t_hi, t_lo, lo, ctx->d_partials, out_dev);
    if (rc < 0) return set_err(ctx, PK_ERR_INVALID_ARG, "eval_univariate: unsupported batch size %d", k);
    ctx->launches += rc;
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}
// out[x] += scale * eq(point, x) over n variables; mode: explicit point (n elements) or univariate z (one element)
int dev_eval_eq(pk_ctx* ctx, const void* point_dev, int mode, int n, const void* scale_dev, void* out) {
    char *t_hi, *t_lo;
    int lo;
    ProfScope ps(ctx, PROF_OTHER);
    PK_TRY(point_tables(ctx, point_dev, mode, n, true, scale_dev, &t_hi, &t_lo, &lo));
    ctx->launches += launch_tensor_accumulate(ctx->stream, out, n, t_hi, t_lo, 1, lo);
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}
int dev_mle_eval_prefix(pk_ctx* ctx, const void* const* evals, int k, int log_n, size_t n_prefix, const void* point_dev,
                        void* out_dev) {
    char *t_hi, *t_lo;
    int lo;
    ProfScope ps(ctx, PROF_OTHER);
    PK_TRY(point_tables(ctx, point_dev, TENSOR_PTS_EXPLICIT, log_n, true, nullptr, &t_hi, &t_lo, &lo));
    int rc = launch_multi_tensor_dot(ctx->stream, evals, k, n_prefix, t_hi, t_lo, lo, ctx->d_partials, out_dev);
    if (rc < 0) return set_err(ctx, PK_ERR_INVALID_ARG, "mle_eval: unsupported batch size %d", k);
    ctx->launches += rc;
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}
// out[x] += sum_{k < count} scalars[k] * eq(pow(omega_D^(exps[k])), x); kmax bounds the batch (grid sizes, method choice),
// count_dev (device, may be null = kmax) is the actual size.  See kernels.cu "eq weights of many UNIVARIATE points".
int dev_eval_eq_roots(pk_ctx* ctx, const uint64_t* exps_dev, const uint32_t* count_dev, size_t kmax, int log_d, int n,
                      const void* scalars_dev, void* out) {
    if (kmax == 0) return PK_OK;
    const size_t N = (size_t)1 << n, D = (size_t)1 << log_d;
    PK_TRY(ensure_twiddles(ctx, log_d));
    // the transform costs ~ (D/2)(log D - 4) + 15 * 2^n multiplications, the direct method k * 2^n
    const bool use_dft = n >= 12 && n <= log_d && log_d >= 8 && (double)kmax * (double)N > 2.0 * (0.5 * D * (log_d - 4) + 16.0 * N);
    if (!use_dft) {
        int lo = split_bits(n), hi = n - lo;
        PK_TRY(ensure_tables(ctx, kmax * (((size_t)1 << lo) + ((size_t)1 << hi))));
        char* t_hi = (char*)ctx->d_tables;
        char* t_lo = t_hi + kmax * ((size_t)32 << hi);
        TensorSrc src = {};
        src.mode = TENSOR_PTS_ROOTS;
        src.exps = exps_dev;
        src.count = count_dev;
        src.W = ctx->d_twiddles;
        src.tbl_shift = ctx->twiddle_log_m - log_d;
        src.log_d = log_d;
        ProfScope ps(ctx, PROF_OTHER);
        ctx->launches += launch_tensor_tables(ctx->stream, nullptr, kmax, n, hi, lo, scalars_dev, true, t_hi, t_lo, src);
        ctx->launches += launch_tensor_accumulate(ctx->stream, out, n, t_hi, t_lo, kmax, lo);
        PK_CUDA(ctx, cudaGetLastError());
        return PK_OK;
    }
    void *sparse = nullptr, *lv = nullptr;
    if (ctx_malloc(ctx, &sparse, D * 32) != cudaSuccess || ctx_malloc(ctx, &lv, D * 32) != cudaSuccess) {
        if (sparse) cudaFreeAsync(sparse, ctx->stream);
        return set_err(ctx, PK_ERR_OOM, "eval_eq_roots: out of device memory");
    }
    int rc = PK_OK;
    do {
        if (cudaMemsetAsync(sparse, 0, D * 32, ctx->stream) != cudaSuccess) {
            rc = set_err(ctx, PK_ERR_CUDA, "eval_eq_roots: staging failed");
            break;
        }
        {
            ProfScope ps(ctx, PROF_OTHER);
            ctx->launches += launch_scatter_add(ctx->stream, sparse, exps_dev, scalars_dev, kmax, count_dev);
        }
        if ((rc = rs_encode_raw(ctx, sparse, log_d, 0, 4, lv, 16, 0)) != PK_OK) break;
        // u overwrites the (consumed) sparse vector, then M^T in place, then out += u
        ProfScope ps(ctx, PROF_OTHER);
        ctx->launches += launch_dft16_combine(ctx->stream, lv, sparse, N, log_d, ctx->d_twiddles, ctx->twiddle_log_m);
        ctx->launches += launch_wavelet_mode(ctx->stream, sparse, n, 2);
        ctx->launches += launch_add_inplace(ctx->stream, out, sparse, N);
        if (cudaGetLastError() != cudaSuccess) rc = set_err(ctx, PK_ERR_CUDA, "eval_eq_roots: launch failed");
    } while (0);
    cudaFreeAsync(sparse, ctx->stream);
    cudaFreeAsync(lv, ctx->stream);
    return rc;
}
}  // namespace pk
extern "C" {

int pk_eval_univariate_batch(pk_ctx* ctx, const pk_buf* const* coeffs, int k, size_t n, const uint64_t z[4], uint64_t* out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, coeffs && z && out && k >= 1 && k <= 3 && n >= 1 && (n & (n - 1)) == 0, "eval_univariate: bad arguments");
    const void* ptrs[3] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < k; j++) {
        PK_CHECK(ctx, coeffs[j] && coeffs[j]->n >= n, "eval_univariate: polynomial %d too small", j);
        ptrs[j] = coeffs[j]->d;
    }
    PK_TRY(ensure_small(ctx, 64));
    PK_TRY(stage_small(ctx, 0, z, 32));
    PK_TRY(dev_eval_univariate(ctx, ptrs, k, n, ctx->d_small, ctx->d_result));
    return fetch_result(ctx, out, k);  // also orders the pageable upload before we return
}
int pk_eval_univariate(pk_ctx* ctx, const pk_buf* coeffs, size_t n, const uint64_t z[4], uint64_t out[4]) {
    PK_BIND(ctx);
    return pk_eval_univariate_batch(ctx, &coeffs, 1, n, z, out);
}
int pk_axpy(pk_ctx* ctx, pk_buf* y, const pk_buf* x, const uint64_t a[4], size_t n) {
    PK_BIND(ctx);
    PK_CHECK(ctx, y && x && a && y->n >= n && x->n >= n, "axpy: buffer too small");
    ProfScope ps(ctx, PROF_OTHER);
    ctx->launches += launch_axpy(ctx->stream, y->d, x->d, to_arg(a), nullptr, n);
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}
int pk_dot(pk_ctx* ctx, const pk_buf* a, const pk_buf* b, size_t n, uint64_t out[4]) {
    PK_BIND(ctx);
    PK_CHECK(ctx, a && b && out && a->n >= n && b->n >= n && n >= 1, "dot: buffer too small");
    ctx->launches += launch_dot(ctx->stream, a->d, b->d, n, ctx->d_partials, ctx->d_result);
    return fetch_result(ctx, out, 1);
}
int pk_multi_dot(pk_ctx* ctx, const pk_buf* const* a, int na, const pk_buf* const* b, int nb, size_t n, uint64_t* out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, a && b && out && n >= 1 && ((na == 3 && nb == 2) || (na == 1 && nb == 2)), "multi_dot: unsupported shape");
    const void *pa[3] = {nullptr, nullptr, nullptr}, *pb[2] = {nullptr, nullptr};
    for (int i = 0; i < na; i++) {
        PK_CHECK(ctx, a[i] && a[i]->n >= n, "multi_dot: a[%d] too small", i);
        pa[i] = a[i]->d;
    }
    for (int i = 0; i < nb; i++) {
        PK_CHECK(ctx, b[i] && b[i]->n >= n, "multi_dot: b[%d] too small", i);
        pb[i] = b[i]->d;
    }
    ctx->launches += launch_multi_dot(ctx->stream, pa, na, pb, nb, n, ctx->d_partials, ctx->d_result);
    return fetch_result(ctx, out, na * nb);
}
int pk_eval_eq_batch(pk_ctx* ctx, const uint64_t* points, size_t k, int n, const uint64_t* scalars, pk_buf* out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, points && scalars && out && n >= 0 && n < 40 && out->n >= ((size_t)1 << n), "eval_eq: bad arguments");
    if (k == 0) return PK_OK;
    int lo = split_bits(n), hi = n - lo;
    size_t pts_bytes = k * (size_t)(n > 0 ? n : 1) * 32, sc_bytes = k * 32;
    PK_TRY(ensure_small(ctx, pts_bytes + sc_bytes));
    PK_TRY(ensure_tables(ctx, k * (((size_t)1 << lo) + ((size_t)1 << hi))));
    char* d_pts = (char*)ctx->d_small;
    char* d_sc = d_pts + pts_bytes;
    if (n > 0) PK_TRY(stage_small(ctx, 0, points, k * (size_t)n * 32));
    PK_TRY(stage_small(ctx, pts_bytes, scalars, sc_bytes));
    char* t_hi = (char*)ctx->d_tables;
    char* t_lo = t_hi + k * ((size_t)32 << hi);
    {
        ProfScope ps(ctx, PROF_OTHER);
        ctx->launches += launch_tensor_tables(ctx->stream, d_pts, k, n, hi, lo, d_sc, true, t_hi, t_lo);
        ctx->launches += launch_tensor_accumulate(ctx->stream, out->d, n, t_hi, t_lo, k, lo);
    }
    PK_CUDA(ctx, cudaGetLastError());
    // the host arrays may be reused by the caller right away
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}
int pk_eval_eq_roots_batch(pk_ctx* ctx, const uint64_t* exps, size_t k, int log_d, int n, const uint64_t* scalars, pk_buf* out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, exps && scalars && out && n >= 0 && n < 40 && log_d >= 0 && log_d <= 28 && out->n >= ((size_t)1 << n),
             "eval_eq_roots: bad arguments");
    if (k == 0) return PK_OK;
    for (size_t i = 0; i < k; i++) PK_CHECK(ctx, exps[i] < ((uint64_t)1 << log_d), "eval_eq_roots: exponent %zu outside the domain", i);
    PK_TRY(ensure_small(ctx, k * 8 + k * 32));
    // 32-byte elements first: they are read with 128-bit loads
    PK_TRY(stage_small(ctx, 0, scalars, k * 32));
    PK_TRY(stage_small(ctx, k * 32, exps, k * 8));
    PK_TRY(dev_eval_eq_roots(ctx, (const uint64_t*)((char*)ctx->d_small + k * 32), nullptr, k, log_d, n, ctx->d_small, out->d));
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host arrays may be reused by the caller right away
    return PK_OK;
}
int pk_eval_eq(pk_ctx* ctx, const uint64_t* point, int n, const uint64_t scalar[4], pk_buf* out) {
    PK_BIND(ctx);
    return pk_eval_eq_batch(ctx, point, 1, n, scalar, out);
}
int pk_mle_eval_batch(pk_ctx* ctx, const pk_buf* const* evals, int k, int log_n, const uint64_t* point, uint64_t* out) {
    return pk_mle_eval_batch_prefix(ctx, evals, k, log_n, (size_t)1 << (log_n >= 0 && log_n < 40 ? log_n : 0), point, out);
}
int pk_mle_eval_batch_prefix(pk_ctx* ctx, const pk_buf* const* evals, int k, int log_n, size_t n_prefix, const uint64_t* point,
                             uint64_t* out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, evals && point && out && k >= 1 && k <= 3 && log_n >= 0 && log_n < 40, "mle_eval: bad arguments");
    PK_CHECK(ctx, n_prefix >= 1 && n_prefix <= ((size_t)1 << log_n), "mle_eval: prefix must be within 2^log_n");
    const void* ptrs[3] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < k; j++) {
        PK_CHECK(ctx, evals[j] && evals[j]->n >= ((size_t)1 << log_n), "mle_eval: array %d too small", j);
        ptrs[j] = evals[j]->d;
    }
    PK_TRY(ensure_small(ctx, (size_t)(log_n + 1) * 32));
    PK_TRY(stage_small(ctx, 0, point, (size_t)log_n * 32));
    PK_TRY(dev_mle_eval_prefix(ctx, ptrs, k, log_n, n_prefix, ctx->d_small, ctx->d_result));
    return fetch_result(ctx, out, k);
}
int pk_mle_eval(pk_ctx* ctx, const pk_buf* evals, int log_n, const uint64_t* point, uint64_t out[4]) {
    PK_BIND(ctx);
    return pk_mle_eval_batch(ctx, &evals, 1, log_n, point, out);
}
int pk_fold_coeffs(pk_ctx* ctx, const pk_buf* coeffs, int log_n, const uint64_t* r, int k, pk_buf* out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, coeffs && r && out && k >= 0 && k <= 4 && log_n >= k && log_n < 40, "fold_coeffs: bad arguments");
    PK_CHECK(ctx, coeffs->n >= ((size_t)1 << log_n) && out->n >= ((size_t)1 << (log_n - k)), "fold_coeffs: buffer too small");
    PK_TRY(ensure_small(ctx, 4 * 32));
    if (k > 0) PK_TRY(stage_small(ctx, 0, r, (size_t)k * 32));
    {
        ProfScope ps(ctx, PROF_OTHER);
        ctx->launches += launch_fold_coeffs(ctx->stream, coeffs->d, log_n, ctx->d_small, 1, k, out->d);
    }
    PK_CUDA(ctx, cudaGetLastError());
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

// ---- sumchecks -----------------------------------------------------------------------------------
int pk_zk_sumcheck_round(pk_ctx* ctx, pk_buf* a, pk_buf* b, pk_buf* c, pk_buf* eq, int log_n, const uint64_t* fold,
                         uint64_t out3[12]) {
    PK_BIND(ctx);
    PK_CHECK(ctx, a && b && c && eq && out3, "zk_sumcheck_round: null argument");
    // sumcheck.rs:21-24: power of two, >= 2, equal lengths; :27: >= 4 when folding
    PK_CHECK(ctx, log_n >= (fold ? 2 : 1) && log_n < 40, "zk_sumcheck_round: size must be >= %d", fold ? 4 : 2);
    size_t n = (size_t)1 << log_n;
    PK_CHECK(ctx, a->n >= n && b->n >= n && c->n >= n && eq->n >= n, "zk_sumcheck_round: arrays shorter than 2^log_n");
    fr_arg f = {};
    if (fold) f = to_arg(fold);
    {
        ProfScope ps(ctx, PROF_ZK_SUMCHECK);
        ctx->launches += launch_zk_sumcheck_round(ctx->stream, a->d, b->d, c->d, eq->d, log_n, fold != nullptr, f, nullptr,
                                                  ctx->d_partials, ctx->d_result);
    }
    return fetch_result(ctx, out3, 3);
}
int pk_whir_sumcheck_round(pk_ctx* ctx, const pk_buf* p_in, const pk_buf* w_in, pk_buf* p_out, pk_buf* w_out, int log_n,
                           const uint64_t* fold, uint64_t out3[12]) {
    PK_BIND(ctx);
    PK_CHECK(ctx, p_in && w_in && out3, "whir_sumcheck_round: null argument");
    PK_CHECK(ctx, log_n >= (fold ? 2 : 1) && log_n < 40, "whir_sumcheck_round: size must be >= %d", fold ? 4 : 2);
    size_t n = (size_t)1 << log_n;
    PK_CHECK(ctx, p_in->n >= n && w_in->n >= n, "whir_sumcheck_round: inputs shorter than 2^log_n");
    fr_arg f = {};
    if (fold) {
        PK_CHECK(ctx, p_out && w_out && p_out->n >= n / 2 && w_out->n >= n / 2, "whir_sumcheck_round: outputs too small");
        PK_CHECK(ctx, p_out->d != p_in->d && w_out->d != w_in->d, "whir_sumcheck_round: folding needs distinct output buffers");
        f = to_arg(fold);
    }
    {
        ProfScope ps(ctx, PROF_WHIR_SUMCHECK);
        ctx->launches += launch_whir_sumcheck_round(ctx->stream, p_in->d, w_in->d, fold ? p_out->d : nullptr,
                                                    fold ? w_out->d : nullptr, log_n, fold != nullptr, f, nullptr, ctx->d_partials,
                                                    ctx->d_result);
    }
    return fetch_result(ctx, out3, 3);
}

// ---- sharded sumchecks: the round's partial sums are exchanged over peer memory and summed on the device --------------
size_t pk_shard_mailbox_elems(void) { return SHARD_MAILBOX_BYTES / 32; }
int pk_shard_group_set(pk_ctx* ctx, int rank, int world, void* const* mailboxes) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && mailboxes, "shard_group_set: null argument");
    PK_CHECK(ctx, world >= 1 && world <= SHARD_MAX_WORLD && (world & (world - 1)) == 0 && rank >= 0 && rank < world,
             "shard_group_set: world must be a power of two <= %d and 0 <= rank < world", SHARD_MAX_WORLD);
    for (int i = 0; i < world; i++) PK_CHECK(ctx, mailboxes[i] != nullptr, "shard_group_set: mailbox %d is null", i);
    if (!ctx->d_shard_status) PK_CUDA(ctx, cudaMalloc((void**)&ctx->d_shard_status, 4));
    if (!ctx->d_shard_gather) PK_CUDA(ctx, cudaMalloc(&ctx->d_shard_gather, (size_t)3 * SHARD_MAX_WORLD * 32));
    PK_CUDA(ctx, cudaMemsetAsync(ctx->d_shard_status, 0, 4, ctx->stream));  // sticky failure flag of the exchange kernels
    ctx->shard = {};
    for (int i = 0; i < world; i++) ctx->shard.mbox[i] = (uint8_t*)mailboxes[i];
    ctx->shard.rank = rank;
    ctx->shard.world = world;
    // Stale flags: a mailbox that served an earlier group may still hold a flag equal to a sequence number the new group
    // will use.  Each rank therefore wipes ITS OWN mailbox here (the caller's barrier between group_set and the first
    // round orders the wipe before any peer's first store), and the sequence number starts at 0 for every group so that all
    // ranks of a new group agree on it whatever their history.
    PK_CUDA(ctx, cudaMemsetAsync(mailboxes[rank], 0, SHARD_MAILBOX_BYTES, ctx->stream));
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->shard_seq = 0;
    return PK_OK;
}
int pk_shard_group_clear(pk_ctx* ctx) {
    PK_BIND(ctx);
    if (!ctx) return PK_ERR_INVALID_ARG;
    ctx->shard = {};
    ctx->shard_seq = 0;
    return PK_OK;
}
// result[0..3) on the device -> global sums -> host
static int exchange_and_fetch(pk_ctx* ctx, uint64_t out3[12]) {
    PK_CHECK(ctx, ctx->shard.world >= 1, "sharded round without pk_shard_group_set");
    ctx->launches += launch_shard_exchange(ctx->stream, ctx->d_result, ctx->shard, ++ctx->shard_seq, ctx->d_shard_status);
    uint32_t status = 0;
    PK_CUDA(ctx, cudaGetLastError());
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, 96, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->h_result + 16, ctx->d_shard_status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(out3, ctx->h_result, 96);
    std::memcpy(&status, ctx->h_result + 16, 4);
    if (status != 0) {
        // the ranks' sequence numbers are no longer in step: the group stays unusable until pk_shard_group_set resets it
        uint32_t seq = ctx->shard_seq;
        ctx->shard = {};
        ctx->shard_seq = 0;
        return set_err(ctx, PK_ERR_CUDA, "sharded round %u: a peer did not publish its partial sums in time; group disabled", seq);
    }
    return PK_OK;
}
// Barrier ON THE STREAM across the group (no host synchronisation): what is enqueued behind it starts only after every rank
// has reached its own pk_shard_barrier.  The sharded commitment puts it between the last NTT pass (which stores rows into the
// peers' leaf blocks) and the leaf hashing.  A rank that never arrives makes the NEXT synchronising call of the group fail.
int pk_shard_barrier(pk_ctx* ctx) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && ctx->shard.world >= 1, "shard_barrier without pk_shard_group_set");
    ctx->launches += launch_shard_gather(ctx->stream, ctx->d_result, ctx->shard, ++ctx->shard_seq, ctx->d_shard_status, ctx->d_shard_gather);
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}
// All-gather of ONE field element per rank over the peer mailboxes: out[r] = rank r's src[off] (4 limbs each, rank order),
// identical on every rank.  The sharded commitment gathers the sub-tree roots with it (32 bytes per rank) instead of a host
// collective.  Synchronises the stream; also a barrier like pk_shard_barrier.
int pk_shard_allgather(pk_ctx* ctx, const pk_buf* src, size_t off, uint64_t* out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && src && out && off < src->n, "shard_allgather: bad argument");
    PK_CHECK(ctx, ctx->shard.world >= 1, "shard_allgather without pk_shard_group_set");
    const int world = ctx->shard.world;
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->d_result, (const char*)src->d + off * 32, 32, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->launches += launch_shard_gather(ctx->stream, ctx->d_result, ctx->shard, ++ctx->shard_seq, ctx->d_shard_status, ctx->d_shard_gather);
    PK_CUDA(ctx, cudaGetLastError());
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_shard_gather, (size_t)3 * world * 32, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(ctx, cudaMemcpyAsync(ctx->h_result + 12 * SHARD_MAX_WORLD, ctx->d_shard_status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t status = 0;
    std::memcpy(&status, ctx->h_result + 12 * SHARD_MAX_WORLD, 4);
    if (status != 0) {
        uint32_t seq = ctx->shard_seq;
        ctx->shard = {};
        ctx->shard_seq = 0;
        return set_err(ctx, PK_ERR_CUDA, "shard_allgather (exchange %u): a peer did not arrive in time; group disabled", seq);
    }
    for (int r = 0; r < world; r++) std::memcpy(out + 4 * r, ctx->h_result + 12 * r, 32);
    return PK_OK;
}
int pk_zk_sumcheck_round_sharded(pk_ctx* ctx, pk_buf* a, pk_buf* b, pk_buf* c, pk_buf* eq, int log_n, const uint64_t* fold,
                                 uint64_t out3[12]) {
    PK_BIND(ctx);
    PK_CHECK(ctx, a && b && c && eq && out3, "zk_sumcheck_round_sharded: null argument");
    PK_CHECK(ctx, log_n >= (fold ? 2 : 1) && log_n < 40, "zk_sumcheck_round_sharded: local size must be >= %d", fold ? 4 : 2);
    size_t n = (size_t)1 << log_n;
    PK_CHECK(ctx, a->n >= n && b->n >= n && c->n >= n && eq->n >= n, "zk_sumcheck_round_sharded: arrays shorter than 2^log_n");
    fr_arg f = {};
    if (fold) f = to_arg(fold);
    ProfScope ps(ctx, PROF_ZK_SUMCHECK);
    ctx->launches += launch_zk_sumcheck_round(ctx->stream, a->d, b->d, c->d, eq->d, log_n, fold != nullptr, f, nullptr,
                                              ctx->d_partials, ctx->d_result);
    return exchange_and_fetch(ctx, out3);
}
int pk_whir_sumcheck_round_sharded(pk_ctx* ctx, const pk_buf* p_in, const pk_buf* w_in, pk_buf* p_out, pk_buf* w_out, int log_n,
                                   const uint64_t* fold, uint64_t out3[12]) {
    PK_BIND(ctx);
    PK_CHECK(ctx, p_in && w_in && out3, "whir_sumcheck_round_sharded: null argument");
    PK_CHECK(ctx, log_n >= (fold ? 2 : 1) && log_n < 40, "whir_sumcheck_round_sharded: local size must be >= %d", fold ? 4 : 2);
    size_t n = (size_t)1 << log_n;
    PK_CHECK(ctx, p_in->n >= n && w_in->n >= n, "whir_sumcheck_round_sharded: inputs shorter than 2^log_n");
    fr_arg f = {};
    if (fold) {
        PK_CHECK(ctx, p_out && w_out && p_out->n >= n / 2 && w_out->n >= n / 2, "whir_sumcheck_round_sharded: outputs too small");
        PK_CHECK(ctx, p_out->d != p_in->d && w_out->d != w_in->d, "whir_sumcheck_round_sharded: folding needs distinct output buffers");
        f = to_arg(fold);
    }
    ProfScope ps(ctx, PROF_WHIR_SUMCHECK);
    ctx->launches += launch_whir_sumcheck_round(ctx->stream, p_in->d, w_in->d, fold ? p_out->d : nullptr, fold ? w_out->d : nullptr,
                                                log_n, fold != nullptr, f, nullptr, ctx->d_partials, ctx->d_result);
    return exchange_and_fetch(ctx, out3);
}

// ---- per-kernel-class device timing (CUDA events on the ctx stream) --------------------------------
int pk_profile_begin(pk_ctx* ctx) {
    PK_BIND(ctx);
    if (!ctx) return PK_ERR_INVALID_ARG;
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->profiling = true;
    ctx->ev_spans.clear();
    ctx->ev_used = 0;
    return PK_OK;
}
int pk_profile_end(pk_ctx* ctx, double ms_by_class[8], uint64_t launches_by_class[8], double max_ms_by_class[8]) {
    PK_BIND(ctx);
    if (!ctx || !ms_by_class || !launches_by_class || !max_ms_by_class) return PK_ERR_INVALID_ARG;
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 8; i++) {
        ms_by_class[i] = 0;
        launches_by_class[i] = 0;
        max_ms_by_class[i] = 0;
    }
    for (const pk_ctx::EvSpan& s : ctx->ev_spans) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev_pool[s.a], ctx->ev_pool[s.b]);
        ms_by_class[s.cls] += ms;
        launches_by_class[s.cls] += 1;
        if (ms > max_ms_by_class[s.cls]) max_ms_by_class[s.cls] = ms;
    }
    ctx->profiling = false;
    ctx->ev_spans.clear();
    ctx->ev_used = 0;
    return PK_OK;
}

// ---- measurement helper (not part of the reference surface) ----------------------------------------
static int modmul_bench(pk_ctx* ctx, size_t n_threads, int iters, float* ms_out, bool square);
int pk_modmul_bench(pk_ctx* ctx, size_t n_threads, int iters, float* ms_out) {
    PK_BIND(ctx);
    return modmul_bench(ctx, n_threads, iters, ms_out, false);
}
int pk_modsqr_bench(pk_ctx* ctx, size_t n_threads, int iters, float* ms_out) {
    PK_BIND(ctx);
    return modmul_bench(ctx, n_threads, iters, ms_out, true);
}
static int modmul_bench(pk_ctx* ctx, size_t n_threads, int iters, float* ms_out, bool square) {
    PK_CHECK(ctx, ctx && ms_out && n_threads % 256 == 0 && n_threads > 0, "modmul_bench: n_threads must be a multiple of 256");
    PK_TRY(ensure_scratch(ctx, 2 * n_threads));
    PK_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0x11, 2 * n_threads * 32, ctx->stream));
    ctx->launches += launch_to_mont(ctx->stream, ctx->d_scratch, 2 * n_threads, true);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch_modmul_bench(ctx->stream, ctx->d_scratch, n_threads, 8, square);  // warm-up
    cudaEventRecord(e0, ctx->stream);
    ctx->launches += launch_modmul_bench(ctx->stream, ctx->d_scratch, n_threads, iters, square);
    cudaEventRecord(e1, ctx->stream);
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return PK_OK;
}

}  // extern "C"
