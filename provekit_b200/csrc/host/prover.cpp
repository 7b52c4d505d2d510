// provekit_b200/csrc/host/prover.cpp — host driver of the whole hot path: pk_prove.
//
// Mirrors `WhirR1CSProver::prove` (provekit/prover/src/whir_r1cs.rs:42-100) with
//   batch_commit_to_polynomial      :182-209   -> Prover::batch_commit
//   run_zk_sumcheck_prover          :228-369   -> Prover::zk_sumcheck
//   create_combined_statement_...   :382-412   -> weights + dot products on device
//   run_zk_whir_pcs_prover          :414-437   -> Prover::whir_prove  ([whir] Prover::prove; message order
//                                                 restated from recursive-verifier/app/circuit/whir.go:51-220)
// Everything data-parallel runs on the GPU through the C-ABI entry points of this library; the host owns
// the transcript (sponge, challenges, hint serialisation) exactly like the Rust host would.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../pk_internal.h"
#include "fr_host.h"
#include "transcript.hpp"

using pkh::Fr;

namespace {

constexpr int FOLD = 4;  // FoldingFactor::Constant(4), provekit/r1cs-compiler/src/whir_r1cs.rs:44

struct RoundCfg {
    double pow_bits;
    int num_queries, ood_samples, log_inv_rate, num_variables, domain_log;
};
// [whir] WhirConfig::new for ProveKit's fixed parameter choice (r1cs-compiler/src/whir_r1cs.rs:38-52);
// reproduces both WhirConfigs stored in the reference fixture poseidon-1000.nps.
struct WhirCfg {
    int num_variables = 0, batch_size = 2, n_rounds = 0, starting_domain_log = 0;
    std::vector<RoundCfg> rounds;
    int final_queries = 0, final_sumcheck_rounds = 0, final_domain_log = 0;
    double final_pow_bits = 0;
    WhirCfg() {}
    WhirCfg(int nv, int batch) : num_variables(nv), batch_size(batch) {
        int max_pow = nv + 1 - 3;  // default_max_pow(num_variables, 1)
        int sec = 128 - max_pow;
        if (sec < 0) sec = 0;
        final_sumcheck_rounds = nv % FOLD;
        n_rounds = (nv - final_sumcheck_rounds) / FOLD - 1;
        starting_domain_log = nv + 1;
        int rate = 1, v = nv - FOLD, dl = starting_domain_log;
        for (int r = 0; r < n_rounds; r++) {
            int q = (sec + rate - 1) / rate;
            double pb = 128.0 - (double)q * rate;
            rounds.push_back({pb > 0 ? pb : 0.0, q, 1, rate, v, dl});
            v -= FOLD;
            rate += FOLD - 1;
            dl -= 1;
        }
        final_queries = (sec + rate - 1) / rate;
        double fb = 128.0 - (double)final_queries * rate;
        final_pow_bits = fb > 0 ? fb : 0.0;
        final_domain_log = dl;
    }
};

int ceil_div(int a, int b) { return (a + b - 1) / b; }
int next_pow2_log(uint64_t n) {  // provekit/common/src/utils/mod.rs:71-79
    int a = 0;
    uint64_t p = 1;
    while (p < n) {
        p <<= 1;
        a++;
    }
    return a;
}

// provekit/common/src/whir_r1cs.rs:28-39 + utils/sumcheck.rs:123-141 + [whir] domainsep (labels unpinned)
void ds_pow(pkh::DomSep& d, double bits) {
    if (bits > 0) d.squeeze(ceil_div(32, 15), "pow-queries").absorb(8, "pow-nonce");
}
void ds_commit(pkh::DomSep& d, const WhirCfg& c) {
    d.absorb(1, "merkle_digest").squeeze(1, "ood_query").absorb(c.batch_size, "ood_ans");
    if (c.batch_size > 1) d.squeeze(1, "batching_randomness");
}
void ds_sumcheck(pkh::DomSep& d, int rounds) {
    for (int i = 0; i < rounds; i++) d.absorb(3, "sumcheck_poly").squeeze(1, "folding_randomness");
}
void ds_whir(pkh::DomSep& d, const WhirCfg& c) {
    d.squeeze(1, "initial_combination_randomness");
    ds_sumcheck(d, FOLD);
    for (const RoundCfg& r : c.rounds) {
        int nb = ceil_div(r.domain_log - FOLD, 8);
        d.absorb(1, "merkle_digest").squeeze(1, "ood_query").absorb(1, "ood_ans");
        ds_pow(d, r.pow_bits);
        d.squeeze(ceil_div(r.num_queries * nb, 15), "stir_queries").hint("stir_answers").hint("merkle_proof");
        d.squeeze(1, "combination_randomness");
        ds_sumcheck(d, FOLD);
    }
    int nb = ceil_div(c.final_domain_log - FOLD, 8);
    d.absorb((size_t)1 << c.final_sumcheck_rounds, "final_coeffs");
    ds_pow(d, c.final_pow_bits);
    d.squeeze(ceil_div(c.final_queries * nb, 15), "final_queries").hint("stir_answers").hint("merkle_proof");
    ds_sumcheck(d, c.final_sumcheck_rounds);
    d.hint("deferred_weight_evaluations");
}
std::string build_domsep(const WhirCfg& cw, const WhirCfg& ch, int m0) {
    pkh::DomSep d("\xF0\x9F\x8C\xAA\xEF\xB8\x8F");  // "🌪️"
    ds_commit(d, cw);
    d.squeeze(m0, "rand");
    ds_commit(d, ch);
    d.absorb(1, "Sum of G over boolean hypercube").squeeze(1, "Rho");
    for (int i = 0; i < m0; i++) d.absorb(4, "Sumcheck Polynomials").squeeze(1, "Sumcheck Random");
    d.absorb(2, "Polynomial sums");
    ds_whir(d, ch);
    d.hint("claimed_evaluations");
    ds_whir(d, cw);
    return d.str();
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct DevCsr {
    uint64_t* row_start = nullptr;
    uint32_t *col = nullptr, *val = nullptr;
    size_t rows = 0, nnz = 0;
    // rows longer than pk::SPMV_LONG_ROW, split into chunks (see kernels.cuh)
    uint64_t *chunk_start = nullptr, *chunk_end = nullptr;
    uint32_t *long_row = nullptr, *long_first = nullptr, *long_cnt = nullptr;
    void* chunk_partials = nullptr;
    size_t n_chunks = 0, n_long = 0;
};
// out = M * x on the device (short rows: thread per row; long rows: chunked)
int spmv(pk_ctx* ctx, const DevCsr& M, const void* interned, const void* x, void* out, size_t rows) {
    ctx->launches += pk::launch_spmv(ctx->stream, M.row_start, M.col, M.val, interned, x, out, rows, M.nnz);
    ctx->launches += pk::launch_spmv_long(ctx->stream, M.col, M.val, interned, x, out, M.chunk_start, M.chunk_end, M.n_chunks,
                                          M.long_row, M.long_first, M.long_cnt, M.n_long, M.chunk_partials);
    PK_CUDA(ctx, cudaGetLastError());
    return PK_OK;
}

}  // namespace

// device + pinned work areas of the protocol flow, one set per prover (one proof at a time per prover)
constexpr size_t NARG_CAP_WORDS = ((size_t)2 << 20) / 4;  // proof string capacity (m = 23 needs ~0.5 MiB)
constexpr size_t HINT_WORDS = 72 * 1024;                  // lens + "stir_answers" + "merkle_proof" of one round
constexpr size_t BOARD_ELEMS = 2048;                      // challenges, round messages and other scalars of one proof
constexpr size_t PIN_SCALAR_BYTES = 32 * 1024;
struct FlowMem {
    char* dev = nullptr;
    void* ts = nullptr;             // pk::DevTs + proof string
    char* board = nullptr;
    uint32_t* hint = nullptr;       // [0,4): lengths; payload area behind
    uint32_t* stir_bytes = nullptr; // 512 words
    uint64_t* stir_idx = nullptr;   // OPEN_MAX_QUERIES
    uint32_t* n_idx = nullptr;
    pk::PowCtrl* pow = nullptr;
    uint8_t* pin = nullptr;         // pinned host memory: [scalar staging | hint mirror | DevTs + proof-string mirror]
    uint8_t *pin_hint = nullptr, *pin_narg = nullptr;
};

struct pk_prover {
    pk_ctx* ctx = nullptr;
    uint64_t num_constraints = 0, num_witnesses = 0, num_interned = 0;
    int m = 0, m0 = 0, mh = 0;
    WhirCfg cw, ch;
    std::string domsep;
    pk::fr_arg iv_canonical = {}, half = {};  // sponge IV (Keccak tag of the domain separator, mod p); 1/2 in Montgomery form
    bool host_transcript = false;             // run the in-tree sponge on the host instead of on the device
    void* d_interned = nullptr;
    DevCsr A, B, At, Bt, Ct;  // rows of A,B for M*z; transposes (CSC) of A,B,C for eq^T*M
    double timings[9] = {0};
    // staged inputs (pk_prover_upload_inputs): [z || mask_w], g_w, [blind || mask_h], g_h in evaluation form
    pk_buf *masked_w = nullptr, *g_w = nullptr, *masked_h = nullptr, *g_h = nullptr;
    bool staged = false;
    FlowMem mem;
    bool pending = false;       // a proof was enqueued (pk_prove_*_enqueue) and not collected yet
    size_t pending_bound = 0;
    double pending_t0 = 0;
    uint64_t host_syncs = 0;  // cudaStreamSynchronize calls of the last proof (pk_prover_host_syncs)
    ~pk_prover();  // frees every device allocation, also after a partially failed pk_prover_create
};

namespace {

struct NvtxRange {  // the reference's tracing spans (SpanStats tree, tooling/cli) as NVTX ranges: an nsys timeline lines up
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// device buffer with RAII over the C-ABI
struct Buf {
    pk_ctx* ctx;
    pk_buf* b = nullptr;
    Buf(pk_ctx* c, size_t n) : ctx(c) {
        if (pk_buf_alloc(c, n, &b) != PK_OK) b = nullptr;
    }
    ~Buf() {
        if (b) pk_buf_free(ctx, b);
    }
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    operator pk_buf*() const { return b; }
    void* d() const { return b->d; }
};
using BufP = std::unique_ptr<Buf>;

inline char* el(void* base, long i) { return (char*)base + 32 * i; }  // element i of a device array of field elements

struct Commitment {
    pk_ctx* ctx;
    pk_commitment* c = nullptr;
    int n = 0, batch = 1, domain_log = 0;
    BufP poly;                                   // batched coefficients (2^n)
    char *root = nullptr, *ood = nullptr, *ans = nullptr, *batching = nullptr;  // board slots
    explicit Commitment(pk_ctx* x) : ctx(x) {}
    ~Commitment() {
        if (c) pk_commit_free(ctx, c);
    }
};

// The protocol flow of WhirR1CSProver::prove.  Every prover message is produced on the device and stays there; what
// differs between the two transcript modes is only how a message reaches the sponge and how a challenge comes back:
//   device transcript (fs == nullptr): one-warp kernels on the pk::DevTs in HBM; the host only enqueues, one sync at the end
//   host transcript   (fs != nullptr): D2H of the message, the caller's (or the in-tree) sponge, H2D of the challenge
class Flow {
   public:
    Flow(pk_prover* p, pkh::Transcript* fs) : P(p), ctx(p->ctx), fs(fs), dev(fs == nullptr), M(p->mem) {}
    // enqueue_only (device transcript): return as soon as everything is enqueued; the proof string's D2H is in flight
    int run(bool enqueue_only = false);
    size_t bound_words = 0;  // upper bound of the proof string written so far (device mode: size of the final D2H)

   private:
    pk_prover* P;
    pk_ctx* ctx;
    pkh::Transcript* fs;
    const bool dev;
    FlowMem& M;
    size_t board_used = 0, pin_cursor = 0;
    double* T() { return P->timings; }
    cudaStream_t st() const { return ctx->stream; }

    char* board(size_t n) {
        if (board_used + n > BOARD_ELEMS) return nullptr;
        char* p = M.board + 32 * board_used;
        board_used += n;
        return p;
    }
    int sync() {
        PK_CUDA(ctx, cudaStreamSynchronize(st()));
        P->host_syncs++;
        pin_cursor = 0;
        return PK_OK;
    }
    int pin_take(size_t bytes, uint8_t** out) {
        bytes = (bytes + 31) & ~(size_t)31;
        if (bytes > PIN_SCALAR_BYTES) return pk::set_err(ctx, PK_ERR_INTERNAL, "flow: staging request too large");
        if (pin_cursor + bytes > PIN_SCALAR_BYTES) PK_TRY(sync());  // earlier async uploads must land before reuse
        *out = M.pin + pin_cursor;
        pin_cursor += bytes;
        return PK_OK;
    }
    int exchange(const void* d_absorb, int na, void* d_squeeze, int ns);
    int challenge_bytes(uint32_t* d_words, int n_bytes);
    int hint_static(const uint32_t* d_payload, uint32_t words);
    int hint_open(uint32_t h1max, uint32_t h2max);
    int pow_prove(double bits);
    int batch_commit(int m, pk_buf* masked_evals, pk_buf* g_evals, Commitment* cm, bool timed);
    int whir_prove(const WhirCfg& cfg, Commitment* cm, pk_buf* const* weights, int n_weights, size_t weights_len);
};

// add_scalars(absorb) then challenge_scalars(squeeze); all values are device-resident Montgomery-form field elements
int Flow::exchange(const void* d_absorb, int na, void* d_squeeze, int ns) {
    bound_words += 8 * (size_t)na;
    if (dev) {
        pk::ProfScope ps(ctx, pk::PROF_OTHER);
        ctx->launches += pk::launch_ts_exchange(st(), M.ts, d_absorb, na, false, d_squeeze, ns);
        PK_CUDA(ctx, cudaGetLastError());
        return PK_OK;
    }
    if (na) {
        uint8_t* h;
        PK_TRY(pin_take(32 * (size_t)na, &h));
        PK_CUDA(ctx, cudaMemcpyAsync(h, d_absorb, 32 * (size_t)na, cudaMemcpyDeviceToHost, st()));
        PK_TRY(sync());
        std::vector<Fr> v(na);
        std::memcpy((void*)v.data(), h, 32 * (size_t)na);
        fs->add_scalars(v.data(), na);
    }
    if (ns) {
        std::vector<Fr> c(ns);
        fs->challenge_scalars(c.data(), ns);
        uint8_t* h;
        PK_TRY(pin_take(32 * (size_t)ns, &h));
        std::memcpy(h, c.data(), 32 * (size_t)ns);
        PK_CUDA(ctx, cudaMemcpyAsync(d_squeeze, h, 32 * (size_t)ns, cudaMemcpyHostToDevice, st()));
    }
    return PK_OK;
}
int Flow::challenge_bytes(uint32_t* d_words, int n_bytes) {
    if (dev) {
        pk::ProfScope ps(ctx, pk::PROF_OTHER);
        ctx->launches += pk::launch_ts_challenge_bytes(st(), M.ts, d_words, n_bytes);
        PK_CUDA(ctx, cudaGetLastError());
        return PK_OK;
    }
    const size_t padded = ((size_t)n_bytes + 3) & ~(size_t)3;
    uint8_t* h;
    PK_TRY(pin_take(padded, &h));
    std::memset(h, 0, padded);
    fs->challenge_bytes(h, n_bytes);
    PK_CUDA(ctx, cudaMemcpyAsync(d_words, h, padded, cudaMemcpyHostToDevice, st()));
    return PK_OK;
}
// ProverState::hint of a payload whose size is known on the host
int Flow::hint_static(const uint32_t* d_payload, uint32_t words) {
    bound_words += 1 + (size_t)words;
    if (dev) {
        pk::ProfScope ps(ctx, pk::PROF_OTHER);
        ctx->launches += pk::launch_ts_hint(st(), M.ts, d_payload, words, nullptr);
        PK_CUDA(ctx, cudaGetLastError());
        return PK_OK;
    }
    PK_CUDA(ctx, cudaMemcpyAsync(M.pin_hint, d_payload, 4 * (size_t)words, cudaMemcpyDeviceToHost, st()));
    PK_TRY(sync());
    fs->hint(std::vector<uint8_t>(M.pin_hint, M.pin_hint + 4 * (size_t)words));
    return PK_OK;
}
// the two hints of one opening (k_open_hints): lengths at M.hint[0], M.hint[1], payloads at M.hint + 4 and + 4 + h1max
int Flow::hint_open(uint32_t h1max, uint32_t h2max) {
    bound_words += 2 + (size_t)h1max + h2max;
    if (dev) {
        pk::ProfScope ps(ctx, pk::PROF_OTHER);
        ctx->launches += pk::launch_ts_hint(st(), M.ts, M.hint + 4, 0, M.hint);
        ctx->launches += pk::launch_ts_hint(st(), M.ts, M.hint + 4 + h1max, 0, M.hint + 1);
        PK_CUDA(ctx, cudaGetLastError());
        return PK_OK;
    }
    const size_t words = 4 + (size_t)h1max + h2max;
    PK_CUDA(ctx, cudaMemcpyAsync(M.pin_hint, M.hint, 4 * words, cudaMemcpyDeviceToHost, st()));
    PK_TRY(sync());
    const uint32_t* w = reinterpret_cast<const uint32_t*>(M.pin_hint);
    if (w[0] > h1max || w[1] > h2max) return pk::set_err(ctx, PK_ERR_INTERNAL, "flow: hint larger than its bound");
    const uint8_t* p1 = M.pin_hint + 16;
    const uint8_t* p2 = p1 + 4 * (size_t)h1max;
    fs->hint(std::vector<uint8_t>(p1, p1 + 4 * (size_t)w[0]));
    fs->hint(std::vector<uint8_t>(p2, p2 + 4 * (size_t)w[1]));
    return PK_OK;
}
// [whir] challenge_pow::<SkyscraperPoW>: 32 challenge bytes, grind, 8-byte big-endian nonce (provekit/common/src/
// skyscraper/pow.rs:14-30 -> skyscraper/core/src/pow.rs:33-41)
int Flow::pow_prove(double bits) {
    if (bits <= 0) return PK_OK;
    NvtxRange nv("pow");
    double t0 = now_s();
    uint64_t thr[4];
    pk::pow_threshold(bits, thr);
    // batches above the first hit exit at once and nonces in flight are abandoned between round pairs, so the cost tracks
    // the winning nonce; small difficulties get small grids (every resident thread hashes at least one nonce)
    int blocks = 1;
    const double want = std::exp2(bits + 3.0) / 128.0;
    while (blocks < 148 * 8 && (double)blocks < want) blocks <<= 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    bound_words += 2;
    if (dev) {
        pk::ProfScope ps(ctx, pk::PROF_POW);
        ctx->launches += pk::launch_pow_begin(st(), M.ts, M.pow);
        ctx->launches += pk::launch_pow_grind(st(), M.pow, pk::to_arg(thr), blocks);
        ctx->launches += pk::launch_pow_end(st(), M.ts, M.pow);
        PK_CUDA(ctx, cudaGetLastError());
        return PK_OK;
    }
    uint8_t* h;
    PK_TRY(pin_take(sizeof(pk::PowCtrl), &h));
    pk::PowCtrl* c = reinterpret_cast<pk::PowCtrl*>(h);
    fs->challenge_bytes(reinterpret_cast<uint8_t*>(c->challenge), 32);
    c->ticket = 0;
    c->best = ~0ull;
    PK_CUDA(ctx, cudaMemcpyAsync(M.pow, c, sizeof(pk::PowCtrl), cudaMemcpyHostToDevice, st()));
    {
        pk::ProfScope ps(ctx, pk::PROF_POW);
        ctx->launches += pk::launch_pow_grind(st(), M.pow, pk::to_arg(thr), blocks);
    }
    PK_CUDA(ctx, cudaGetLastError());
    uint8_t* hb;
    PK_TRY(pin_take(8, &hb));
    PK_CUDA(ctx, cudaMemcpyAsync(hb, &M.pow->best, 8, cudaMemcpyDeviceToHost, st()));
    PK_TRY(sync());
    uint64_t nonce;
    std::memcpy(&nonce, hb, 8);
    uint8_t nb[8];
    for (int i = 0; i < 8; i++) nb[i] = (uint8_t)(nonce >> (56 - 8 * i));  // big-endian nonce
    fs->add_bytes(nb, 8);
    T()[4] += now_s() - t0;
    return PK_OK;
}

// batch_commit_to_polynomial (whir_r1cs.rs:182-209) on the staged evaluation vectors [f || mask] and g, then [whir]
// CommitmentWriter::commit_batch: root -> OOD point -> OOD answers -> batching randomness -> batched polynomial
int Flow::batch_commit(int m, pk_buf* masked_evals, pk_buf* g_evals, Commitment* cm, bool timed) {
    NvtxRange nv("batch_commit_to_polynomial");
    const size_t N = (size_t)1 << m;
    BufP mc(new Buf(ctx, N)), gc(new Buf(ctx, N));
    if (!mc->b || !gc->b) return pk::set_err(ctx, PK_ERR_OOM, "batch_commit: out of device memory");
    PK_TRY(pk_buf_copy(ctx, *mc, 0, masked_evals, 0, N));
    PK_TRY(pk_buf_copy(ctx, *gc, 0, g_evals, 0, N));
    PK_TRY(pk_evals_to_coeffs(ctx, *mc, m));
    PK_TRY(pk_evals_to_coeffs(ctx, *gc, m));
    const void* polys[2] = {mc->d(), gc->d()};
    cm->root = board(1);
    cm->ood = board(1);
    cm->ans = board(2);
    cm->batching = board(1);
    if (!cm->batching) return pk::set_err(ctx, PK_ERR_INTERNAL, "flow: scalar board exhausted");
    double t0 = now_s();
    PK_TRY(pk::commit_batch_dev(ctx, polys, 2, m, 1, &cm->c, cm->root));
    cm->n = m;
    cm->batch = 2;
    cm->domain_log = m + 1;
    PK_TRY(exchange(cm->root, 1, cm->ood, 1));
    if (timed && !dev) T()[0] += now_s() - t0;  // NTT + Merkle of the big commitment (split by CUDA events in bench.py)
    PK_TRY(pk::dev_eval_univariate(ctx, polys, 2, N, cm->ood, cm->ans));
    PK_TRY(exchange(cm->ans, 2, cm->batching, 1));
    {   // batched polynomial p0 + b*p1
        pk::ProfScope ps(ctx, pk::PROF_OTHER);
        ctx->launches += pk::launch_axpy(st(), mc->d(), gc->d(), pk::fr_arg(), cm->batching, N);
    }
    PK_CUDA(ctx, cudaGetLastError());
    cm->poly = std::move(mc);
    return PK_OK;
}

// [whir] Prover::prove (message order restated from recursive-verifier/app/circuit/whir.go:51-220)
// weights_len: the linear weights vanish beyond their first weights_len elements (zero-extended R1CS rows)
int Flow::whir_prove(const WhirCfg& cfg, Commitment* cm, pk_buf* const* weights, int n_weights, size_t weights_len) {
    NvtxRange nv("run_zk_whir_pcs_prover");
    const int n = cfg.num_variables;
    const size_t N = (size_t)1 << n;
    if (weights_len == 0 || weights_len > N) weights_len = N;
    char* gamma = board(1);
    char* gpow = board(3);
    char* h3 = board(3);
    char* RR = board(n);  // folding randomness, NEWEST FIRST: challenge number t lives at RR[n-1-t]
    char* dv = board(3);
    if (!dv) return pk::set_err(ctx, PK_ERR_INTERNAL, "flow: scalar board exhausted");
    PK_TRY(exchange(nullptr, 0, gamma, 1));
    ctx->launches += pk::launch_powers(st(), gpow, gamma, n_weights, 1, nullptr);
    BufP Pb[2] = {BufP(new Buf(ctx, N)), BufP(new Buf(ctx, N / 2))};
    BufP Wb[2] = {BufP(new Buf(ctx, N)), BufP(new Buf(ctx, N / 2))};
    if (!Pb[0]->b || !Pb[1]->b || !Wb[0]->b || !Wb[1]->b) return pk::set_err(ctx, PK_ERR_OOM, "whir_prove: out of device memory");
    PK_TRY(pk_buf_zero(ctx, *Wb[0], 0, N));
    // W = eq(pow(ood_point), .) + sum_j gamma^(j+1) w_j : the OOD constraint goes first
    PK_TRY(pk::dev_eval_eq(ctx, cm->ood, pk::TENSOR_PTS_UNIVARIATE, n, nullptr, Wb[0]->d()));
    {
        pk::ProfScope ps(ctx, pk::PROF_OTHER);
        for (int j = 0; j < n_weights; j++)
            ctx->launches += pk::launch_axpy(st(), Wb[0]->d(), weights[j]->d, pk::fr_arg(), el(gpow, j), weights_len);
    }
    PK_TRY(pk_buf_copy(ctx, *Pb[0], 0, *cm->poly, 0, N));
    PK_TRY(pk_coeffs_to_evals(ctx, *Pb[0], n));

    int which = 0, cur_log = n, t = 0;  // t: folding challenges drawn so far
    bool pending = false;               // the fold by challenge t-1 has not been applied to P, W yet
    // one sumcheck round = (fold by the previous challenge, fused) + [h(0), h(1), h(2)] -> transcript -> next challenge
    auto sumcheck_rounds = [&](int rounds) -> int {
        double t0 = now_s();
        for (int i = 0; i < rounds; i++) {
            const int w = which;
            {
                pk::ProfScope ps(ctx, pk::PROF_WHIR_SUMCHECK);
                if (pending) {
                    ctx->launches += pk::launch_whir_sumcheck_round(st(), Pb[w]->d(), Wb[w]->d(), Pb[1 - w]->d(), Wb[1 - w]->d(), cur_log,
                                                                  true, pk::fr_arg(), el(RR, n - t), ctx->d_partials, h3);
                    which = 1 - w;
                    cur_log--;
                } else {
                    ctx->launches += pk::launch_whir_sumcheck_round(st(), Pb[w]->d(), Wb[w]->d(), nullptr, nullptr, cur_log, false,
                                                                  pk::fr_arg(), nullptr, ctx->d_partials, h3);
                }
            }
            PK_CUDA(ctx, cudaGetLastError());
            PK_TRY(exchange(h3, 3, el(RR, n - 1 - t), 1));
            t++;
            pending = true;
        }
        if (!dev) T()[3] += now_s() - t0;
        return PK_OK;
    };
    PK_TRY(sumcheck_rounds(FOLD));

    BufP coeffs;                       // folded coefficient list of the current round (null: still cm->poly)
    const void* cur_coeffs = cm->poly->d();
    int nv_ = n, domain_log = cm->domain_log;
    pk_commitment* prev = cm->c;
    std::unique_ptr<Commitment> prev_owned;
    for (int ri = 0; ri <= cfg.n_rounds; ri++) {
        const bool is_final = ri == cfg.n_rounds;
        const int nvp = nv_ - FOLD;
        BufP folded(new Buf(ctx, (size_t)1 << nvp));
        if (!folded->b) return pk::set_err(ctx, PK_ERR_OOM, "whir_prove: out of device memory");
        {   // CoefficientList::fold by the four challenges of this block (r[j] binds bit j): RR is newest-first, stride -1
            pk::ProfScope ps(ctx, pk::PROF_OTHER);
            ctx->launches += pk::launch_fold_coeffs(st(), cur_coeffs, nv_, el(RR, n - 1 - FOLD * ri), -1, FOLD, folded->d());
        }
        PK_CUDA(ctx, cudaGetLastError());
        std::unique_ptr<Commitment> next;
        char* ood_pt = nullptr;
        double pow_bits;
        int nq;
        if (!is_final) {
            const RoundCfg& rc = cfg.rounds[ri];
            const int new_dl = domain_log - 1;
            next.reset(new Commitment(ctx));
            next->root = board(1);
            ood_pt = board(1);
            char* ood_ans = board(1);
            if (!ood_ans) return pk::set_err(ctx, PK_ERR_INTERNAL, "flow: scalar board exhausted");
            const void* polys[1] = {folded->d()};
            PK_TRY(pk::commit_batch_dev(ctx, polys, 1, nvp, new_dl - nvp, &next->c, next->root));
            next->n = nvp;
            next->domain_log = new_dl;
            PK_TRY(exchange(next->root, 1, ood_pt, 1));
            PK_TRY(pk::dev_eval_univariate(ctx, polys, 1, (size_t)1 << nvp, ood_pt, ood_ans));
            PK_TRY(exchange(ood_ans, 1, nullptr, 0));
            pow_bits = rc.pow_bits;
            nq = rc.num_queries;
        } else {
            PK_TRY(exchange(folded->d(), 1 << nvp, nullptr, 0));  // the final coefficients go out in the clear
            pow_bits = cfg.final_pow_bits;
            nq = cfg.final_queries;
        }
        PK_TRY(pow_prove(pow_bits));
        double t_open = now_s();
        {   // STIR queries into the previous oracle, answers and Merkle multi-proof as hints
            NvtxRange nvo("stir_open");
            const int folded_log = domain_log - FOLD, nb = ceil_div(folded_log, 8);
            if (nq > pk::OPEN_MAX_QUERIES || nq * nb > 2048)
                return pk::set_err(ctx, PK_ERR_INTERNAL, "whir_prove: %d queries exceed the device work areas", nq);
            PK_TRY(challenge_bytes(M.stir_bytes, nq * nb));
            const uint32_t w = (uint32_t)prev->w, depth = (uint32_t)prev->depth, q = (uint32_t)nq;
            const uint32_t h1max = 2 + q * (2 + 8 * w);
            const uint32_t h2max = (2 + 8 * q) + (2 + 2 * q) + (2 + 2 * q + 8 * q * (depth > 1 ? depth - 1 : 0)) + (2 + 2 * q);
            if (4 + (size_t)h1max + h2max > HINT_WORDS) return pk::set_err(ctx, PK_ERR_INTERNAL, "whir_prove: hints exceed the work area");
            {
                pk::ProfScope ps(ctx, pk::PROF_OTHER);
                ctx->launches += pk::launch_stir_indices(st(), M.stir_bytes, nq, nb, folded_log, M.stir_idx, M.n_idx);
                ctx->launches += pk::launch_open_hints(st(), prev->leaves, prev->canonical_leaves, (int)w, prev->nodes, prev->L, (int)depth,
                                                     M.stir_idx, M.n_idx, M.hint + 4, M.hint + 4 + h1max, M.hint);
            }
            PK_CUDA(ctx, cudaGetLastError());
            PK_TRY(hint_open(h1max, h2max));
        }
        if (!dev) T()[5] += now_s() - t_open;
        if (!is_final) {
            char* gam = board(1);
            char* sc = board((size_t)nq + 1);
            if (!sc) return pk::set_err(ctx, PK_ERR_INTERNAL, "flow: scalar board exhausted");
            PK_TRY(exchange(nullptr, 0, gam, 1));
            // combination scalars 1, gam, gam^2, ..: the OOD point first, then the (de-duplicated) STIR points
            ctx->launches += pk::launch_powers(st(), sc, gam, nq + 1, 0, M.n_idx);
            // the fold by the last challenge must land before new equality weights are added
            if (pending) {
                pk::ProfScope ps(ctx, pk::PROF_WHIR_SUMCHECK);
                ctx->launches += pk::launch_whir_sumcheck_round(st(), Pb[which]->d(), Wb[which]->d(), Pb[1 - which]->d(), Wb[1 - which]->d(),
                                                              cur_log, true, pk::fr_arg(), el(RR, n - t), ctx->d_partials, ctx->d_result);
                which = 1 - which;
                cur_log--;
                pending = false;
            }
            // new equality constraints: the OOD point, then the STIR points z_q = omega_D^idx[q] on the folded domain
            // D = 2^(domain_log - 4) (recursive-verifier/app/circuit/whir.go:139-142); large batches take the sparse-DFT form
            PK_TRY(pk::dev_eval_eq(ctx, ood_pt, pk::TENSOR_PTS_UNIVARIATE, nvp, sc, Wb[which]->d()));
            PK_TRY(pk::dev_eval_eq_roots(ctx, M.stir_idx, M.n_idx, (size_t)nq, domain_log - FOLD, nvp, el(sc, 1), Wb[which]->d()));
            PK_TRY(sumcheck_rounds(FOLD));
        } else {
            PK_TRY(sumcheck_rounds(cfg.final_sumcheck_rounds));
        }
        coeffs = std::move(folded);
        cur_coeffs = coeffs->d();
        nv_ = nvp;
        if (!is_final) {
            prev_owned = std::move(next);
            prev = prev_owned->c;
            domain_log -= 1;
        }
    }
    // deferred weight evaluations at the reversed folding randomness (RR is stored reversed already)
    const void* wp[3] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < n_weights; j++) wp[j] = weights[j]->d;
    PK_TRY(pk::dev_mle_eval_prefix(ctx, wp, n_weights, n, weights_len, RR, dv));
    ctx->launches += pk::launch_hint_scalars(st(), M.hint + 4, dv, 1, n_weights, 1, 0);
    PK_TRY(hint_static(M.hint + 4, 2 + 8 * (uint32_t)n_weights));
    return PK_OK;
}

int Flow::run(bool enqueue_only) {
    NvtxRange nv("prove");
    const double t_start = now_s();
    const int m = P->m, m0 = P->m0, mh = P->mh;
    const size_t N = (size_t)1 << m, N0 = (size_t)1 << m0, Nh = (size_t)1 << mh;
    const size_t nw = P->num_witnesses, nc = P->num_constraints;
    P->host_syncs = 0;
    if (dev) {
        ctx->launches += pk::launch_ts_init(st(), M.ts, P->iv_canonical, (uint32_t)NARG_CAP_WORDS);
        PK_CUDA(ctx, cudaGetLastError());
    }
    // [f || mask] and g in evaluation form were staged by pk_prover_upload_inputs
    pk_buf *masked_w = P->masked_w, *g_w = P->g_w, *masked_h = P->masked_h, *g_h = P->g_h;
    Commitment cmw(ctx);
    PK_TRY(batch_commit(m, masked_w, g_w, &cmw, true));

    // ---- zk-sumcheck (run_zk_sumcheck_prover, whir_r1cs.rs:228-369) ----
    char* r = board(m0);
    char* alpha = board(m0);
    pk::ZkGlue G = {};
    G.blind = (const pk::fr*)masked_h->d;  // the 4 m_0 blinding coefficients open the staged [blind || mask_h] vector
    G.suffix = (pk::fr*)board(m0);
    G.state = (pk::fr*)board(4);
    G.alpha = (pk::fr*)alpha;
    G.h3 = (pk::fr*)board(3);
    G.cf = (pk::fr*)board(4);
    G.m0 = m0;
    char* fg = board(2);
    char* fg6 = board(6);
    if (!fg6) return pk::set_err(ctx, PK_ERR_INTERNAL, "flow: scalar board exhausted");
    Buf a(ctx, N0), b(ctx, N0), c(ctx, N0), eq(ctx, N0);
    if (!a.b || !b.b || !c.b || !eq.b) return pk::set_err(ctx, PK_ERR_OOM, "prove: out of device memory");
    Commitment cmh(ctx);
    {
        NvtxRange nvz("run_zk_sumcheck_prover");
        PK_TRY(exchange(nullptr, 0, r, m0));
        double t0 = now_s();
        {
            NvtxRange nvb("calculate_witness_bounds");
            PK_TRY(pk_buf_zero(ctx, a, 0, N0));
            PK_TRY(pk_buf_zero(ctx, b, 0, N0));
            PK_TRY(pk_buf_zero(ctx, c, 0, N0));
            // sumcheck.rs:181-193: a = A z, b = B z, c = a o b; z = first nw of masked_w
            pk::ProfScope ps(ctx, pk::PROF_OTHER);
            PK_TRY(spmv(ctx, P->A, P->d_interned, masked_w->d, a.d(), nc));
            PK_TRY(spmv(ctx, P->B, P->d_interned, masked_w->d, b.d(), nc));
            ctx->launches += pk::launch_mul(st(), a.d(), b.d(), c.d(), nc);
        }
        {
            NvtxRange nve("calculate_evaluations_over_boolean_hypercube_for_eq");
            PK_TRY(pk_buf_zero(ctx, eq, 0, N0));
            PK_TRY(pk::dev_eval_eq(ctx, r, pk::TENSOR_PTS_EXPLICIT, m0, nullptr, eq.d()));
        }
        if (!dev) T()[6] += now_s() - t0;

        PK_TRY(batch_commit(mh, masked_h, g_h, &cmh, false));
        {
            pk::ProfScope ps(ctx, pk::PROF_OTHER);
            ctx->launches += pk::launch_zk_init(st(), G, P->half);
        }
        PK_TRY(exchange(el(G.state, 3), 1, el(G.state, 0), 1));  // sum of g over the hypercube -> rho
        int cur = m0;
        t0 = now_s();
        for (int idx = 0; idx < m0; idx++) {
            {
                pk::ProfScope ps(ctx, pk::PROF_ZK_SUMCHECK);
                ctx->launches += pk::launch_zk_sumcheck_round(st(), a.d(), b.d(), c.d(), eq.d(), cur, idx > 0, pk::fr_arg(),
                                                            idx > 0 ? el(alpha, idx - 1) : nullptr, ctx->d_partials, G.h3);
            }
            if (idx) cur--;
            {   // blinded cubic of the round: 4 coefficients out, alpha_idx back (whir_r1cs.rs:300-345)
                pk::ProfScope ps(ctx, pk::PROF_OTHER);
                ctx->launches += pk::launch_zk_glue(st(), G, idx, P->half, dev ? M.ts : nullptr);
            }
            PK_CUDA(ctx, cudaGetLastError());
            if (dev)
                bound_words += 32;
            else
                PK_TRY(exchange(G.cf, 4, el(alpha, idx), 1));
        }
        if (!dev) T()[2] += now_s() - t0;

        // statement over the blinding commitment: weight = expand_powers(alpha), zero-extended (whir_r1cs.rs:347-377)
        Buf dw(ctx, Nh);
        if (!dw.b) return pk::set_err(ctx, PK_ERR_OOM, "prove: out of device memory");
        {
            pk::ProfScope ps(ctx, pk::PROF_OTHER);
            ctx->launches += pk::launch_expand_powers(st(), dw.d(), alpha, m0, Nh);
            const void* wa[1] = {dw.d()};
            const void* fb[2] = {masked_h->d, g_h->d};
            ctx->launches += pk::launch_multi_dot(st(), wa, 1, fb, 2, Nh, ctx->d_partials, fg);
        }
        PK_CUDA(ctx, cudaGetLastError());
        PK_TRY(exchange(fg, 2, nullptr, 0));
        pk_buf* ws[1] = {dw};
        PK_TRY(whir_prove(P->ch, &cmh, ws, 1, 0));
    }

    // ---- weights from the R1CS instance: eq(alpha)^T * {A,B,C}, zero-extended to 2^m ----
    double t0 = now_s();
    BufP wts[3] = {BufP(new Buf(ctx, N)), BufP(new Buf(ctx, N)), BufP(new Buf(ctx, N))};
    {
        NvtxRange nvr("calculate_external_row_of_r1cs_matrices");
        Buf eq_alpha(ctx, N0);
        if (!eq_alpha.b || !wts[0]->b || !wts[1]->b || !wts[2]->b) return pk::set_err(ctx, PK_ERR_OOM, "prove: out of device memory");
        PK_TRY(pk_buf_zero(ctx, eq_alpha, 0, N0));
        PK_TRY(pk::dev_eval_eq(ctx, alpha, pk::TENSOR_PTS_EXPLICIT, m0, nullptr, eq_alpha.d()));
        const DevCsr* T3[3] = {&P->At, &P->Bt, &P->Ct};
        for (int j = 0; j < 3; j++) {
            PK_TRY(pk_buf_zero(ctx, *wts[j], 0, N));
            pk::ProfScope ps(ctx, pk::PROF_OTHER);
            PK_TRY(spmv(ctx, *T3[j], P->d_interned, eq_alpha.d(), wts[j]->d(), nw));
        }
        // the weights vanish beyond the witness: num_witnesses terms suffice for <w_j, f> and <w_j, g>
        pk::ProfScope ps(ctx, pk::PROF_OTHER);
        const void* wa[3] = {wts[0]->d(), wts[1]->d(), wts[2]->d()};
        const void* fb[2] = {masked_w->d, g_w->d};
        ctx->launches += pk::launch_multi_dot(st(), wa, 3, fb, 2, nw, ctx->d_partials, fg6);
        // hint claimed_evaluations: (Vec<F>, Vec<F>) = (f_sums, g_sums); fg6 holds them interleaved
        ctx->launches += pk::launch_hint_scalars(st(), M.hint + 4, fg6, 2, 3, 2, 1);
    }
    PK_CUDA(ctx, cudaGetLastError());
    if (!dev) T()[6] += now_s() - t0;
    PK_TRY(hint_static(M.hint + 4, 2 * (2 + 8 * 3)));
    pk_buf* ws[3] = {*wts[0], *wts[1], *wts[2]};
    PK_TRY(whir_prove(P->cw, &cmw, ws, 3, nw));
    if (dev) {
        // the one host round trip of the device-transcript flow: header + proof string
        const size_t bytes = pk::DEVTS_HEADER_BYTES + 4 * std::min(bound_words, NARG_CAP_WORDS);
        PK_CUDA(ctx, cudaMemcpyAsync(M.pin_narg, M.ts, bytes, cudaMemcpyDeviceToHost, st()));
        if (enqueue_only) return PK_OK;
    }
    PK_TRY(sync());
    if (fs && !fs->ok()) return pk::set_err(ctx, PK_ERR_INVALID_ARG, "prove: a transcript callback reported failure");
    T()[8] += now_s() - t_start;
    T()[7] = T()[8] - (T()[0] + T()[1] + T()[2] + T()[3] + T()[4] + T()[5] + T()[6]);
    return PK_OK;
}

int upload_csr_arrays(pk_ctx* ctx, const std::vector<uint64_t>& rs, const std::vector<uint32_t>& col,
                      const std::vector<uint32_t>& val, DevCsr* out) {
    out->rows = rs.size();
    out->nnz = col.size();
    PK_CUDA(ctx, cudaMalloc((void**)&out->row_start, (rs.size() + 1) * 8));
    PK_CUDA(ctx, cudaMalloc((void**)&out->col, (col.size() + 1) * 4));
    PK_CUDA(ctx, cudaMalloc((void**)&out->val, (val.size() + 1) * 4));
    PK_CUDA(ctx, cudaMemcpy(out->row_start, rs.data(), rs.size() * 8, cudaMemcpyHostToDevice));
    PK_CUDA(ctx, cudaMemcpy(out->col, col.data(), col.size() * 4, cudaMemcpyHostToDevice));
    PK_CUDA(ctx, cudaMemcpy(out->val, val.data(), val.size() * 4, cudaMemcpyHostToDevice));
    // long rows -> chunks
    std::vector<uint64_t> cs, ce;
    std::vector<uint32_t> lrow, lfirst, lcnt;
    for (size_t r = 0; r < rs.size(); r++) {
        uint64_t s = rs[r], e = r + 1 < rs.size() ? rs[r + 1] : col.size();
        if (e - s <= (uint64_t)pk::SPMV_LONG_ROW) continue;
        lrow.push_back((uint32_t)r);
        lfirst.push_back((uint32_t)cs.size());
        uint32_t n = 0;
        for (uint64_t k = s; k < e; k += pk::SPMV_CHUNK, n++) {
            cs.push_back(k);
            ce.push_back(k + pk::SPMV_CHUNK < e ? k + pk::SPMV_CHUNK : e);
        }
        lcnt.push_back(n);
    }
    out->n_chunks = cs.size();
    out->n_long = lrow.size();
    if (out->n_long) {
        PK_CUDA(ctx, cudaMalloc((void**)&out->chunk_start, cs.size() * 8));
        PK_CUDA(ctx, cudaMalloc((void**)&out->chunk_end, cs.size() * 8));
        PK_CUDA(ctx, cudaMalloc((void**)&out->long_row, lrow.size() * 4));
        PK_CUDA(ctx, cudaMalloc((void**)&out->long_first, lrow.size() * 4));
        PK_CUDA(ctx, cudaMalloc((void**)&out->long_cnt, lrow.size() * 4));
        PK_CUDA(ctx, cudaMalloc(&out->chunk_partials, cs.size() * 32));
        PK_CUDA(ctx, cudaMemcpy(out->chunk_start, cs.data(), cs.size() * 8, cudaMemcpyHostToDevice));
        PK_CUDA(ctx, cudaMemcpy(out->chunk_end, ce.data(), ce.size() * 8, cudaMemcpyHostToDevice));
        PK_CUDA(ctx, cudaMemcpy(out->long_row, lrow.data(), lrow.size() * 4, cudaMemcpyHostToDevice));
        PK_CUDA(ctx, cudaMemcpy(out->long_first, lfirst.data(), lfirst.size() * 4, cudaMemcpyHostToDevice));
        PK_CUDA(ctx, cudaMemcpy(out->long_cnt, lcnt.data(), lcnt.size() * 4, cudaMemcpyHostToDevice));
    }
    return PK_OK;
}
int check_csr(pk_ctx* ctx, const pk_csr& m, uint64_t rows, uint64_t cols, uint64_t n_interned) {
    PK_CHECK(ctx, m.num_rows == rows && m.num_cols == cols, "R1CS matrix shape mismatch");
    PK_CHECK(ctx, (rows == 0 || m.row_start) && (m.nnz == 0 || (m.col && m.val)), "R1CS matrix has null arrays");
    for (uint64_t r = 0; r < rows; r++) {
        uint64_t s = m.row_start[r], e = r + 1 < rows ? m.row_start[r + 1] : m.nnz;
        PK_CHECK(ctx, s <= e && e <= m.nnz, "R1CS row offsets not monotone");
    }
    for (uint64_t k = 0; k < m.nnz; k++) PK_CHECK(ctx, m.col[k] < cols && m.val[k] < n_interned, "R1CS entry out of range");
    return PK_OK;
}
int upload_csr(pk_ctx* ctx, const pk_csr& m, DevCsr* rows_out, DevCsr* transposed_out) {
    if (rows_out) {
        std::vector<uint64_t> rs(m.row_start, m.row_start + m.num_rows);
        std::vector<uint32_t> col(m.col, m.col + m.nnz), val(m.val, m.val + m.nnz);
        PK_TRY(upload_csr_arrays(ctx, rs, col, val, rows_out));
    }
    // CSC: for eq^T * M  (Mul<HydratedSparseMatrix> for &[FieldElement], sparse_matrix.rs:168-184)
    std::vector<uint64_t> cs(m.num_cols + 1, 0);
    for (uint64_t k = 0; k < m.nnz; k++) cs[m.col[k] + 1]++;
    for (uint64_t c = 0; c < m.num_cols; c++) cs[c + 1] += cs[c];
    std::vector<uint32_t> rowidx(m.nnz), val(m.nnz);
    std::vector<uint64_t> fill(cs.begin(), cs.end() - 1);
    for (uint64_t r = 0; r < m.num_rows; r++) {
        uint64_t s = m.row_start[r], e = r + 1 < m.num_rows ? m.row_start[r + 1] : m.nnz;
        for (uint64_t k = s; k < e; k++) {
            uint64_t pos = fill[m.col[k]]++;
            rowidx[pos] = (uint32_t)r;
            val[pos] = m.val[k];
        }
    }
    cs.pop_back();
    return upload_csr_arrays(ctx, cs, rowidx, val, transposed_out);
}
void free_csr(DevCsr& c) {
    if (!c.row_start && !c.col && !c.val && !c.chunk_start) return;
    cudaFree(c.row_start);
    cudaFree(c.col);
    cudaFree(c.val);
    cudaFree(c.chunk_start);
    cudaFree(c.chunk_end);
    cudaFree(c.long_row);
    cudaFree(c.long_first);
    cudaFree(c.long_cnt);
    cudaFree(c.chunk_partials);
}

}  // namespace

pk_prover::~pk_prover() {
    if (!ctx) return;
    PK_BIND(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(d_interned);
    cudaFree(mem.dev);
    if (mem.pin) cudaFreeHost(mem.pin);
    pk_buf_free(ctx, masked_w);
    pk_buf_free(ctx, g_w);
    pk_buf_free(ctx, masked_h);
    pk_buf_free(ctx, g_h);
    free_csr(A);
    free_csr(B);
    free_csr(At);
    free_csr(Bt);
    free_csr(Ct);
}

extern "C" {

int pk_prover_create(pk_ctx* ctx, const pk_r1cs* r1cs, pk_prover** out) {
    PK_BIND(ctx);
    PK_CHECK(ctx, ctx && r1cs && out, "prover_create: null argument");
    *out = nullptr;
    PK_CHECK(ctx, r1cs->num_constraints >= 2 && r1cs->num_witnesses >= 1 && r1cs->interned, "prover_create: empty R1CS");
    PK_TRY(check_csr(ctx, r1cs->a, r1cs->num_constraints, r1cs->num_witnesses, r1cs->num_interned));
    PK_TRY(check_csr(ctx, r1cs->b, r1cs->num_constraints, r1cs->num_witnesses, r1cs->num_interned));
    PK_TRY(check_csr(ctx, r1cs->c, r1cs->num_constraints, r1cs->num_witnesses, r1cs->num_interned));
    std::unique_ptr<pk_prover> p(new pk_prover());
    p->ctx = ctx;
    p->num_constraints = r1cs->num_constraints;
    p->num_witnesses = r1cs->num_witnesses;
    p->num_interned = r1cs->num_interned;
    // scheme shapes: provekit/r1cs-compiler/src/whir_r1cs.rs:15-36
    p->m = next_pow2_log(r1cs->num_witnesses) + 1;
    p->m0 = next_pow2_log(r1cs->num_constraints);
    p->mh = next_pow2_log(4 * (uint64_t)p->m0) + 1;
    PK_CHECK(ctx, p->m >= FOLD + 1 && p->m <= 27 && p->m0 >= 1 && p->mh >= FOLD + 1, "prover_create: unsupported R1CS size");
    p->cw = WhirCfg(p->m, 2);
    p->ch = WhirCfg(p->mh, 2);
    p->domsep = build_domsep(p->cw, p->ch, p->m0);
    PK_CUDA(ctx, cudaMalloc(&p->d_interned, (size_t)r1cs->num_interned * 32 + 32));
    PK_CUDA(ctx, cudaMemcpy(p->d_interned, r1cs->interned, (size_t)r1cs->num_interned * 32, cudaMemcpyHostToDevice));
    PK_TRY(upload_csr(ctx, r1cs->a, &p->A, &p->At));
    PK_TRY(upload_csr(ctx, r1cs->b, &p->B, &p->Bt));
    PK_TRY(upload_csr(ctx, r1cs->c, nullptr, &p->Ct));
    {   // sponge IV: Keccak tag of the domain separator as a field element (sponge.rs:47-52); 1/2 (utils/mod.rs:23-25)
        uint8_t tag[32];
        uint64_t c[4];
        pkh::domsep_tag(p->domsep, tag);
        std::memcpy(c, tag, 32);
        pkh::to_canonical(pkh::from_canonical(c), c);
        p->iv_canonical = pk::to_arg(c);
        p->half = pk::to_arg(pkh::half().l);
    }
    {   // work areas of the protocol flow: one device allocation, one pinned allocation
        FlowMem& M = p->mem;
        const size_t ts_bytes = pk::DEVTS_HEADER_BYTES + 4 * NARG_CAP_WORDS, board_bytes = 32 * BOARD_ELEMS, hint_bytes = 4 * HINT_WORDS;
        const size_t stir_bytes = 2048, idx_bytes = 8 * pk::OPEN_MAX_QUERIES, misc_bytes = 256;
        PK_CUDA(ctx, cudaMalloc((void**)&M.dev, ts_bytes + board_bytes + hint_bytes + stir_bytes + idx_bytes + misc_bytes));
        char* q = M.dev;
        M.ts = q;
        q += ts_bytes;
        M.board = q;
        q += board_bytes;
        M.hint = (uint32_t*)q;
        q += hint_bytes;
        M.stir_bytes = (uint32_t*)q;
        q += stir_bytes;
        M.stir_idx = (uint64_t*)q;
        q += idx_bytes;
        M.n_idx = (uint32_t*)q;
        M.pow = (pk::PowCtrl*)(q + 64);
        PK_CUDA(ctx, cudaMemset(M.dev, 0, ts_bytes + board_bytes + hint_bytes + stir_bytes + idx_bytes + misc_bytes));
        PK_CUDA(ctx, cudaMallocHost((void**)&M.pin, PIN_SCALAR_BYTES + hint_bytes + ts_bytes));
        M.pin_hint = M.pin + PIN_SCALAR_BYTES;
        M.pin_narg = M.pin_hint + hint_bytes;
    }
    p->host_transcript = std::getenv("PK_HOST_TRANSCRIPT") && std::getenv("PK_HOST_TRANSCRIPT")[0] == '1';
    *out = p.release();
    return PK_OK;
}
int pk_prover_set_host_transcript(pk_prover* p, int on) {
    if (!p) return PK_ERR_INVALID_ARG;
    p->host_transcript = on != 0;
    return PK_OK;
}
uint64_t pk_prover_host_syncs(const pk_prover* p) { return p ? p->host_syncs : 0; }
void pk_prover_destroy(pk_prover* p) { delete p; }
// seam: `impl Mul<&[FieldElement]> for HydratedSparseMatrix` and its transposed twin (provekit/common/src/sparse_matrix.rs:
// 148-184) on the device-resident R1CS: out = M x (x: num_witnesses, out: num_constraints) or out = x^T M (x:
// num_constraints, out: num_witnesses).  C x is not available: the prover never forms it (c = a o b, sumcheck.rs:181-193).
int pk_prover_matvec(pk_prover* p, int which, int transposed, const pk_buf* x, pk_buf* out) {
    if (!p) return PK_ERR_INVALID_ARG;
    pk_ctx* ctx = p->ctx;
    PK_BIND(ctx);
    PK_CHECK(ctx, x && out && which >= 0 && which <= 2, "matvec: bad arguments");
    PK_CHECK(ctx, transposed || which < 2, "matvec: C x is not formed on this path (c = a o b)");
    const DevCsr* M = transposed ? (which == 0 ? &p->At : which == 1 ? &p->Bt : &p->Ct) : (which == 0 ? &p->A : &p->B);
    const size_t n_in = transposed ? p->num_constraints : p->num_witnesses, n_out = transposed ? p->num_witnesses : p->num_constraints;
    PK_CHECK(ctx, x->n >= n_in && out->n >= n_out, "matvec: buffer too small");
    PK_TRY(pk_buf_zero(ctx, out, 0, n_out));
    return spmv(ctx, *M, p->d_interned, x->d, out->d, n_out);
}
void pk_prover_shapes(const pk_prover* p, int* m, int* m0, int* mh) {
    if (m) *m = p ? p->m : 0;
    if (m0) *m0 = p ? p->m0 : 0;
    if (mh) *mh = p ? p->mh : 0;
}
}  // extern "C"
namespace {
// H2D staging of one proof's inputs: witness (zero-padded) || mask, g, blinding cubics || mask, g.  do_sync: return only
// when the host arrays may be reused (the stand-alone entry points); inside pk_prove the proof's own final
// synchronisation covers it
// H2D in pieces of PK_H2D_CHUNK_MB (0 / unset: one copy): experiment knob, see DESIGN.md section 6
static cudaError_t h2d(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    static const size_t chunk = [] {
        const char* e = std::getenv("PK_H2D_CHUNK_MB");
        return e ? (size_t)std::atoi(e) << 20 : (size_t)0;
    }();
    if (!chunk || bytes <= chunk) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    for (size_t off = 0; off < bytes; off += chunk) {
        cudaError_t e = cudaMemcpyAsync((char*)dst + off, (const char*)src + off, std::min(chunk, bytes - off), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
int upload_inputs(pk_prover* p, const uint64_t* witness, const pk_rand* rnd, bool do_sync) {
    if (!p) return PK_ERR_INVALID_ARG;
    pk_ctx* ctx = p->ctx;
    PK_BIND(ctx);
    PK_CHECK(ctx, witness && rnd, "upload_inputs: null argument");
    PK_CHECK(ctx, !p->pending, "upload_inputs: an enqueued proof still reads the staged inputs (pk_prove_collect first)");
    PK_CHECK(ctx, rnd->mask_w && rnd->g_w && rnd->blind && rnd->mask_h && rnd->g_h, "upload_inputs: null randomness");
    const size_t half = (size_t)1 << (p->m - 1), N = (size_t)1 << p->m;
    const size_t halfh = (size_t)1 << (p->mh - 1), Nh = (size_t)1 << p->mh;
    if (!p->masked_w) {
        PK_TRY(pk_buf_alloc(ctx, N, &p->masked_w));
        PK_TRY(pk_buf_alloc(ctx, N, &p->g_w));
        PK_TRY(pk_buf_alloc(ctx, Nh, &p->masked_h));
        PK_TRY(pk_buf_alloc(ctx, Nh, &p->g_h));
    }
    double t0 = now_s();
    cudaStream_t st = ctx->stream;
    // create_masked_polynomial (zk_utils.rs:3-11): [pad_to_power_of_two(witness) || mask]
    PK_CUDA(ctx, cudaMemsetAsync((char*)p->masked_w->d + p->num_witnesses * 32, 0, (half - p->num_witnesses) * 32, st));
    PK_CUDA(ctx, h2d(p->masked_w->d, witness, p->num_witnesses * 32, st));
    PK_CUDA(ctx, h2d((char*)p->masked_w->d + half * 32, rnd->mask_w, half * 32, st));
    PK_CUDA(ctx, h2d(p->g_w->d, rnd->g_w, N * 32, st));
    PK_CUDA(ctx, cudaMemsetAsync(p->masked_h->d, 0, halfh * 32, st));
    PK_CUDA(ctx, cudaMemcpyAsync(p->masked_h->d, rnd->blind, 4 * (size_t)p->m0 * 32, cudaMemcpyHostToDevice, st));
    PK_CUDA(ctx, cudaMemcpyAsync((char*)p->masked_h->d + halfh * 32, rnd->mask_h, halfh * 32, cudaMemcpyHostToDevice, st));
    PK_CUDA(ctx, cudaMemcpyAsync(p->g_h->d, rnd->g_h, Nh * 32, cudaMemcpyHostToDevice, st));
    if (do_sync) PK_CUDA(ctx, cudaStreamSynchronize(st));
    std::memset(p->timings, 0, sizeof p->timings);
    p->timings[1] = now_s() - t0;  // H2D staging time
    p->timings[8] = p->timings[1];
    p->staged = true;
    return PK_OK;
}
// same staging with the masks drawn ON THE DEVICE from `seed` (pk_rng_fill streams 0..4 = mask_w, g_w, blind, mask_h,
// g_h): only the witness crosses PCIe.  The reference draws them from thread_rng inside prove (whir_r1cs.rs:211-225).
int upload_inputs_seeded(pk_prover* p, const uint64_t* witness, const uint8_t seed[32], bool do_sync) {
    if (!p) return PK_ERR_INVALID_ARG;
    pk_ctx* ctx = p->ctx;
    PK_BIND(ctx);
    PK_CHECK(ctx, witness && seed, "upload_inputs_seeded: null argument");
    PK_CHECK(ctx, !p->pending, "upload_inputs: an enqueued proof still reads the staged inputs (pk_prove_collect first)");
    const size_t half = (size_t)1 << (p->m - 1), N = (size_t)1 << p->m;
    const size_t halfh = (size_t)1 << (p->mh - 1), Nh = (size_t)1 << p->mh;
    const size_t nb = 4 * (size_t)p->m0;
    if (!p->masked_w) {
        PK_TRY(pk_buf_alloc(ctx, N, &p->masked_w));
        PK_TRY(pk_buf_alloc(ctx, N, &p->g_w));
        PK_TRY(pk_buf_alloc(ctx, Nh, &p->masked_h));
        PK_TRY(pk_buf_alloc(ctx, Nh, &p->g_h));
    }
    double t0 = now_s();
    cudaStream_t st = ctx->stream;
    PK_CUDA(ctx, cudaMemcpyAsync(p->masked_w->d, witness, p->num_witnesses * 32, cudaMemcpyHostToDevice, st));
    PK_CUDA(ctx, cudaMemsetAsync((char*)p->masked_w->d + p->num_witnesses * 32, 0, (half - p->num_witnesses) * 32, st));
    PK_TRY(pk_rng_fill(ctx, p->masked_w, half, half, seed, 0));
    PK_TRY(pk_rng_fill(ctx, p->g_w, 0, N, seed, 1));
    PK_CUDA(ctx, cudaMemsetAsync(p->masked_h->d, 0, halfh * 32, st));
    PK_TRY(pk_rng_fill(ctx, p->masked_h, 0, nb, seed, 2));
    PK_TRY(pk_rng_fill(ctx, p->masked_h, halfh, halfh, seed, 3));
    PK_TRY(pk_rng_fill(ctx, p->g_h, 0, Nh, seed, 4));
    if (do_sync) PK_CUDA(ctx, cudaStreamSynchronize(st));
    std::memset(p->timings, 0, sizeof p->timings);
    p->timings[1] = now_s() - t0;
    p->timings[8] = p->timings[1];
    p->staged = true;
    return PK_OK;
}
}  // namespace
extern "C" {
int pk_prover_upload_inputs(pk_prover* p, const uint64_t* witness, const pk_rand* rnd) { return upload_inputs(p, witness, rnd, true); }
int pk_prover_upload_inputs_seeded(pk_prover* p, const uint64_t* witness, const uint8_t seed[32]) {
    return upload_inputs_seeded(p, witness, seed, true);
}
}  // extern "C"
namespace {
// the input uploads are asynchronous: on a failed proof make sure they no longer read the caller's arrays
int drain_on_error(pk_prover* p, int rc) {
    if (rc != PK_OK && p && p->ctx && p->ctx->stream) cudaStreamSynchronize(p->ctx->stream);
    return rc;
}
// the finished proof string of a device-transcript run: header check + copy out of the pinned mirror
int take_device_proof(pk_prover* p, size_t bound_words, uint8_t** out, size_t* out_len) {
    pk_ctx* ctx = p->ctx;
    uint32_t hdr[24];
    std::memcpy(hdr, p->mem.pin_narg, sizeof hdr);
    const uint32_t words = hdr[18], err = hdr[20];  // pk::DevTs: narg_words, error
    if (err != 0 || words > bound_words)
        return pk::set_err(ctx, PK_ERR_INTERNAL, "prove: device transcript reported error %u (%u words, bound %zu)", err, words, bound_words);
    const size_t len = 4 * (size_t)words;
    uint8_t* buf = (uint8_t*)std::malloc(len ? len : 1);
    if (!buf) return pk::set_err(ctx, PK_ERR_OOM, "prove: host allocation failed");
    std::memcpy(buf, p->mem.pin_narg + pk::DEVTS_HEADER_BYTES, len);
    *out = buf;
    *out_len = len;
    return PK_OK;
}
int enqueue_staged(pk_prover* p) {
    pk_ctx* ctx = p->ctx;
    PK_CHECK(ctx, !p->host_transcript, "prove_enqueue: needs the device transcript (pk_prover_set_host_transcript(p, 0))");
    PK_CHECK(ctx, p->staged, "prove_enqueue: inputs not staged");
    PK_CHECK(ctx, !p->pending, "prove_enqueue: the previous proof of this prover has not been collected");
    std::memset(p->timings, 0, sizeof p->timings);
    p->pending_t0 = now_s();
    Flow fl(p, nullptr);
    PK_TRY(fl.run(true));
    p->pending = true;
    p->pending_bound = fl.bound_words;
    return PK_OK;
}
}  // namespace
extern "C" {
// ---- asynchronous form of the device-transcript prover: ONE host thread keeps many proofs in flight ----
int pk_prove_staged_enqueue(pk_prover* p) {
    if (!p) return PK_ERR_INVALID_ARG;
    PK_BIND(p->ctx);
    return enqueue_staged(p);
}
int pk_prove_seeded_enqueue(pk_prover* p, const uint64_t* witness, const uint8_t seed[32]) {
    if (!p) return PK_ERR_INVALID_ARG;
    PK_BIND(p->ctx);
    PK_CHECK(p->ctx, !p->pending, "prove_enqueue: the previous proof of this prover has not been collected");
    PK_TRY(upload_inputs_seeded(p, witness, seed, false));
    return drain_on_error(p, enqueue_staged(p));
}
int pk_prove_enqueue(pk_prover* p, const uint64_t* witness, const pk_rand* rnd) {
    if (!p) return PK_ERR_INVALID_ARG;
    PK_BIND(p->ctx);
    PK_CHECK(p->ctx, !p->pending, "prove_enqueue: the previous proof of this prover has not been collected");
    PK_TRY(upload_inputs(p, witness, rnd, false));
    return drain_on_error(p, enqueue_staged(p));
}
int pk_prove_collect(pk_prover* p, uint8_t** out, size_t* out_len) {
    if (!p) return PK_ERR_INVALID_ARG;
    pk_ctx* ctx = p->ctx;
    PK_BIND(ctx);
    PK_CHECK(ctx, out && out_len, "prove_collect: null argument");
    PK_CHECK(ctx, p->pending, "prove_collect: nothing enqueued");
    p->pending = false;
    PK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    p->host_syncs = 1;
    p->timings[8] = now_s() - p->pending_t0;
    return take_device_proof(p, p->pending_bound, out, out_len);
}
int pk_prove_staged(pk_prover* p, uint8_t** out, size_t* out_len) {
    if (!p) return PK_ERR_INVALID_ARG;
    pk_ctx* ctx = p->ctx;
    PK_BIND(ctx);
    PK_CHECK(ctx, out && out_len, "prove: null argument");
    PK_CHECK(ctx, p->staged, "prove_staged: call pk_prover_upload_inputs first");
    PK_CHECK(ctx, !p->pending, "prove: an enqueued proof of this prover has not been collected");
    double keep1 = p->timings[1];
    std::memset(p->timings, 0, sizeof p->timings);
    p->timings[1] = keep1;
    p->timings[8] = keep1;
    const uint8_t* src;
    size_t len;
    std::unique_ptr<pkh::ProverState> host_fs;
    if (p->host_transcript) {  // the in-tree sponge on the host: every challenge is a host round trip
        host_fs.reset(new pkh::ProverState(p->domsep));
        Flow fl(p, host_fs.get());
        PK_TRY(fl.run());
        src = host_fs->narg().data();
        len = host_fs->narg().size();
    } else {  // default: the transcript lives on the device, the proof string arrives with the final synchronisation
        Flow fl(p, nullptr);
        PK_TRY(fl.run());
        return take_device_proof(p, fl.bound_words, out, out_len);
    }
    uint8_t* buf = (uint8_t*)std::malloc(len ? len : 1);
    if (!buf) return pk::set_err(ctx, PK_ERR_OOM, "prove: host allocation failed");
    std::memcpy(buf, src, len);
    *out = buf;
    *out_len = len;
    return PK_OK;
}
int pk_prove_staged_with_transcript(pk_prover* p, const pk_transcript_vtbl* vt, void* user) {
    if (!p) return PK_ERR_INVALID_ARG;
    pk_ctx* ctx = p->ctx;
    PK_BIND(ctx);
    PK_CHECK(ctx, vt && vt->add_scalars && vt->challenge_scalars && vt->add_bytes && vt->challenge_bytes && vt->hint,
             "prove_with_transcript: incomplete transcript vtable");
    PK_CHECK(ctx, p->staged, "prove_staged: call pk_prover_upload_inputs first");
    double keep1 = p->timings[1];
    std::memset(p->timings, 0, sizeof p->timings);
    p->timings[1] = keep1;
    p->timings[8] = keep1;
    pkh::CallbackTranscript fs(vt, user);
    Flow fl(p, &fs);
    return fl.run();
}
int pk_prove_with_transcript(pk_prover* p, const uint64_t* witness, const pk_rand* rnd, const pk_transcript_vtbl* vt, void* user) {
    PK_TRY(upload_inputs(p, witness, rnd, false));
    int rc = pk_prove_staged_with_transcript(p, vt, user);
    if (rc != PK_OK && p && p->ctx && p->ctx->stream) cudaStreamSynchronize(p->ctx->stream);
    return rc;
}
int pk_prove(pk_prover* p, const uint64_t* witness, const pk_rand* rnd, uint8_t** out, size_t* out_len) {
    PK_TRY(upload_inputs(p, witness, rnd, false));
    return drain_on_error(p, pk_prove_staged(p, out, out_len));
}
int pk_prove_seeded(pk_prover* p, const uint64_t* witness, const uint8_t seed[32], uint8_t** out, size_t* out_len) {
    PK_TRY(upload_inputs_seeded(p, witness, seed, false));
    return drain_on_error(p, pk_prove_staged(p, out, out_len));
}
void pk_free(void* p) { std::free(p); }
void pk_prover_timings(const pk_prover* p, double out[9]) {
    for (int i = 0; i < 9; i++) out[i] = p ? p->timings[i] : 0.0;
}

}  // extern "C"
