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

Fr eval_cubic(const Fr c[4], const Fr& x) {  // sumcheck.rs:174-176
    return pkh::add(c[0], pkh::mul(x, pkh::add(c[1], pkh::mul(x, pkh::add(c[2], pkh::mul(x, c[3]))))));
}
void expand_from_univariate(Fr z, int n, Fr* out) {
    for (int i = 0; i < n; i++) {
        out[n - 1 - i] = z;
        z = pkh::sqr(z);
    }
}
// compute_blinding_coefficients_for_round, whir_r1cs.rs:103-171
void blinding_coeffs_for_round(const Fr* g, int n, int compute_for, const Fr* alphas, Fr out[4]) {
    bool all_fixed = false;
    if (compute_for == n) {
        all_fixed = true;
        compute_for = n - 1;
    }
    Fr prefix = pkh::ZERO, suffix = pkh::ZERO;
    for (int i = 0; i < compute_for; i++) prefix = pkh::add(prefix, eval_cubic(g + 4 * i, alphas[i]));
    for (int i = compute_for + 1; i < n; i++)
        suffix = pkh::add(suffix, pkh::add(eval_cubic(g + 4 * i, pkh::ZERO), eval_cubic(g + 4 * i, pkh::ONE)));
    Fr pm = pkh::ONE;
    for (int i = 0; i < n - 1 - compute_for; i++) pm = pkh::dbl(pm);
    Fr sm = pkh::mul(pm, pkh::half());
    Fr cst = pkh::add(pkh::mul(pm, prefix), pkh::mul(sm, suffix));
    const Fr* c = g + 4 * compute_for;
    Fr r[4] = {pkh::add(pkh::mul(pm, c[0]), cst), pkh::mul(pm, c[1]), pkh::mul(pm, c[2]), pkh::mul(pm, c[3])};
    if (all_fixed) {
        out[0] = eval_cubic(r, alphas[compute_for]);
        out[1] = out[2] = out[3] = pkh::ZERO;
    } else {
        std::memcpy(out, r, sizeof r);
    }
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

struct pk_prover {
    pk_ctx* ctx = nullptr;
    uint64_t num_constraints = 0, num_witnesses = 0, num_interned = 0;
    int m = 0, m0 = 0, mh = 0;
    WhirCfg cw, ch;
    std::string domsep;
    void* d_interned = nullptr;
    DevCsr A, B, At, Bt, Ct;  // rows of A,B for M*z; transposes (CSC) of A,B,C for eq^T*M
    double timings[9] = {0};
    // staged inputs (pk_prover_upload_inputs): [z || mask_w], g_w, [blind || mask_h], g_h in evaluation form
    pk_buf *masked_w = nullptr, *g_w = nullptr, *masked_h = nullptr, *g_h = nullptr;
    std::vector<Fr> blind;
    bool staged = false;
    ~pk_prover();  // frees every device allocation, also after a partially failed pk_prover_create
};

namespace {

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
};
using BufP = std::unique_ptr<Buf>;

struct Commitment {
    pk_ctx* ctx;
    pk_commitment* c = nullptr;
    int n = 0, batch = 1, domain_log = 0;
    BufP poly;  // batched coefficients (2^n)
    Fr ood_point, ood_answer, batching;
    explicit Commitment(pk_ctx* x) : ctx(x) {}
    ~Commitment() {
        if (c) pk_commit_free(ctx, c);
    }
};

class Prover {
   public:
    Prover(pk_prover* p, pkh::Transcript& fs) : P(p), ctx(p->ctx), fs(fs) {}
    int run();

   private:
    pk_prover* P;
    pk_ctx* ctx;
    pkh::Transcript& fs;
    double* T() { return P->timings; }

    int batch_commit(int m, pk_buf* masked_evals, pk_buf* g_evals, Commitment* cm, bool timed);
    int whir_prove(const WhirCfg& cfg, Commitment* cm, pk_buf* const* weights, const Fr* sums, int n_weights, size_t weights_len);
    int whir_sumcheck_rounds(BufP (&Pb)[2], BufP (&Wb)[2], int* which, int* cur_log, int rounds, Fr* rs, bool* pending,
                             Fr* pending_r);
    int pow_prove(double bits);
    size_t stir_queries(int domain_log, int num_queries, std::vector<uint64_t>* idx);
};

int Prover::pow_prove(double bits) {
    if (bits <= 0) return PK_OK;
    double t0 = now_s();
    uint8_t ch[32], nb[8];
    uint64_t c[4], nonce = 0;
    fs.challenge_bytes(ch, 32);
    std::memcpy(c, ch, 32);
    PK_TRY(pk_pow_solve(ctx, c, bits, &nonce));
    for (int i = 0; i < 8; i++) nb[i] = (uint8_t)(nonce >> (56 - 8 * i));  // big-endian nonce
    fs.add_bytes(nb, 8);
    T()[4] += now_s() - t0;
    return PK_OK;
}

// [whir] get_challenge_stir_queries (recursive-verifier/app/circuit/whir_utilities.go:48-77) + sort/dedup
size_t Prover::stir_queries(int domain_log, int num_queries, std::vector<uint64_t>* idx) {
    int folded_log = domain_log - FOLD;
    int nb = ceil_div(folded_log, 8);
    std::vector<uint8_t> bytes((size_t)num_queries * nb);
    fs.challenge_bytes(bytes.data(), bytes.size());
    idx->resize(num_queries);
    for (int i = 0; i < num_queries; i++) {
        uint64_t v = 0;
        for (int j = 0; j < nb; j++) v = (v << 8) | bytes[(size_t)i * nb + j];
        (*idx)[i] = v & (((uint64_t)1 << folded_log) - 1);
    }
    std::sort(idx->begin(), idx->end());
    idx->erase(std::unique(idx->begin(), idx->end()), idx->end());
    return idx->size();
}

// [whir] CommitmentWriter::commit_batch on the coefficient forms of [f || mask] and g
int Prover::batch_commit(int m, pk_buf* masked_evals, pk_buf* g_evals, Commitment* cm, bool timed) {
    size_t N = (size_t)1 << m;
    BufP mc(new Buf(ctx, N)), gc(new Buf(ctx, N));
    if (!mc->b || !gc->b) return pk::set_err(ctx, PK_ERR_OOM, "batch_commit: out of device memory");
    PK_TRY(pk_buf_copy(ctx, *mc, 0, masked_evals, 0, N));
    PK_TRY(pk_buf_copy(ctx, *gc, 0, g_evals, 0, N));
    PK_TRY(pk_evals_to_coeffs(ctx, *mc, m));
    PK_TRY(pk_evals_to_coeffs(ctx, *gc, m));
    const pk_buf* polys[2] = {*mc, *gc};
    uint64_t root[4];
    double t0 = now_s();
    PK_TRY(pk_commit_batch(ctx, polys, 2, m, 1, FOLD, &cm->c, root));
    if (timed) T()[0] += now_s() - t0;  // NTT + Merkle of the big commitment (split by CUDA events in bench.py)
    cm->n = m;
    cm->batch = 2;
    cm->domain_log = m + 1;
    Fr rootf;
    std::memcpy(rootf.l, root, 32);
    fs.add_scalars(&rootf, 1);
    fs.challenge_scalars(&cm->ood_point, 1);
    Fr ans[2];
    PK_TRY(pk_eval_univariate_batch(ctx, polys, 2, N, cm->ood_point.l, ans[0].l));
    fs.add_scalars(ans, 2);
    fs.challenge_scalars(&cm->batching, 1);
    PK_TRY(pk_axpy(ctx, *mc, *gc, cm->batching.l, N));  // batched polynomial p0 + b*p1
    cm->ood_answer = pkh::add(ans[0], pkh::mul(cm->batching, ans[1]));
    cm->poly = std::move(mc);
    return PK_OK;
}

int Prover::whir_sumcheck_rounds(BufP (&Pb)[2], BufP (&Wb)[2], int* which, int* cur_log, int rounds, Fr* rs, bool* pending,
                                 Fr* pending_r) {
    double t0 = now_s();
    for (int i = 0; i < rounds; i++) {
        Fr h[3];
        int w = *which;
        if (*pending) {
            PK_TRY(pk_whir_sumcheck_round(ctx, *Pb[w], *Wb[w], *Pb[1 - w], *Wb[1 - w], *cur_log, pending_r->l, h[0].l));
            *which = 1 - w;
            (*cur_log)--;
        } else {
            PK_TRY(pk_whir_sumcheck_round(ctx, *Pb[w], *Wb[w], nullptr, nullptr, *cur_log, nullptr, h[0].l));
        }
        fs.add_scalars(h, 3);
        fs.challenge_scalars(&rs[i], 1);
        *pending = true;
        *pending_r = rs[i];
    }
    T()[3] += now_s() - t0;
    return PK_OK;
}

// [whir] Prover::prove
// weights_len: the linear weights vanish beyond their first weights_len elements (zero-extended R1CS rows)
int Prover::whir_prove(const WhirCfg& cfg, Commitment* cm, pk_buf* const* weights, const Fr* sums, int n_weights,
                       size_t weights_len) {
    const int n = cfg.num_variables;
    const size_t N = (size_t)1 << n;
    if (weights_len == 0 || weights_len > N) weights_len = N;
    Fr gamma, g = pkh::ONE;
    fs.challenge_scalars(&gamma, 1);
    BufP Pb[2] = {BufP(new Buf(ctx, N)), BufP(new Buf(ctx, N / 2))};
    BufP Wb[2] = {BufP(new Buf(ctx, N)), BufP(new Buf(ctx, N / 2))};
    if (!Pb[0]->b || !Pb[1]->b || !Wb[0]->b || !Wb[1]->b) return pk::set_err(ctx, PK_ERR_OOM, "whir_prove: out of device memory");
    PK_TRY(pk_buf_zero(ctx, *Wb[0], 0, N));
    std::vector<Fr> pt(n);
    expand_from_univariate(cm->ood_point, n, pt.data());
    PK_TRY(pk_eval_eq(ctx, pt[0].l, n, g.l, *Wb[0]));  // OOD constraint goes first
    for (int j = 0; j < n_weights; j++) {
        g = pkh::mul(g, gamma);
        PK_TRY(pk_axpy(ctx, *Wb[0], weights[j], g.l, weights_len));
    }
    (void)sums;  // the claimed sum only enters the verifier's checks; h(0), h(1), h(2) are computed directly
    PK_TRY(pk_buf_copy(ctx, *Pb[0], 0, *cm->poly, 0, N));
    PK_TRY(pk_coeffs_to_evals(ctx, *Pb[0], n));

    std::vector<Fr> all_r;
    int which = 0, cur_log = n;
    bool pending = false;
    Fr pending_r = pkh::ZERO, fold_r[FOLD];
    PK_TRY(whir_sumcheck_rounds(Pb, Wb, &which, &cur_log, FOLD, fold_r, &pending, &pending_r));
    all_r.insert(all_r.end(), fold_r, fold_r + FOLD);

    BufP coeffs;              // folded coefficient list of the current round (null: still cm->poly)
    pk_buf* cur_coeffs = *cm->poly;
    int nv = n, domain_log = cm->domain_log;
    pk_commitment* prev = cm->c;
    std::unique_ptr<Commitment> prev_owned;
    for (int ri = 0; ri <= cfg.n_rounds; ri++) {
        const bool is_final = ri == cfg.n_rounds;
        const int nvp = nv - FOLD;
        BufP folded(new Buf(ctx, (size_t)1 << nvp));
        if (!folded->b) return pk::set_err(ctx, PK_ERR_OOM, "whir_prove: out of device memory");
        PK_TRY(pk_fold_coeffs(ctx, cur_coeffs, nv, fold_r[0].l, FOLD, *folded));
        std::unique_ptr<Commitment> next;
        Fr ood_pt = pkh::ZERO, ood_ans = pkh::ZERO;
        double pow_bits;
        int nq;
        if (!is_final) {
            const RoundCfg& rc = cfg.rounds[ri];
            const int new_dl = domain_log - 1;
            next.reset(new Commitment(ctx));
            const pk_buf* polys[1] = {*folded};
            uint64_t root[4];
            PK_TRY(pk_commit_batch(ctx, polys, 1, nvp, new_dl - nvp, FOLD, &next->c, root));
            next->n = nvp;
            next->domain_log = new_dl;
            Fr rootf;
            std::memcpy(rootf.l, root, 32);
            fs.add_scalars(&rootf, 1);
            fs.challenge_scalars(&ood_pt, 1);
            PK_TRY(pk_eval_univariate(ctx, *folded, (size_t)1 << nvp, ood_pt.l, ood_ans.l));
            fs.add_scalars(&ood_ans, 1);
            pow_bits = rc.pow_bits;
            nq = rc.num_queries;
        } else {
            std::vector<Fr> fc((size_t)1 << nvp);
            PK_TRY(pk_buf_download(ctx, *folded, 0, fc[0].l, fc.size()));
            fs.add_scalars(fc.data(), fc.size());
            pow_bits = cfg.final_pow_bits;
            nq = cfg.final_queries;
        }
        PK_TRY(pow_prove(pow_bits));
        double t_open = now_s();
        std::vector<uint64_t> idx;
        const size_t nidx = stir_queries(domain_log, nq, &idx);
        const size_t w = pk_commit_leaf_width(prev);
        const int depth = domain_log - FOLD;
        std::vector<Fr> leaves(nidx * w), sib(nidx), suf(nidx * (size_t)(depth > 0 ? depth : 1));
        std::vector<uint64_t> pre(nidx), slen(nidx);
        PK_TRY(pk_commit_open(ctx, prev, idx.data(), nidx, leaves[0].l, sib[0].l, pre.data(), suf[0].l, slen.data(), suf.size()));
        {   // hint stir_answers: Vec<Vec<F>>
            std::vector<uint8_t> hb;
            hb.reserve(16 + nidx * (8 + w * 32));
            pkh::put_u64(hb, nidx);
            for (size_t q = 0; q < nidx; q++) {
                pkh::put_u64(hb, w);
                for (size_t k = 0; k < w; k++) pkh::put_fr(hb, leaves[q * w + k]);
            }
            fs.hint(hb);
        }
        {   // hint merkle_proof: ark MultiPath (digests already canonical)
            std::vector<uint8_t> hb;
            pkh::put_u64(hb, nidx);
            for (size_t q = 0; q < nidx; q++) pkh::put_canonical(hb, sib[q].l);
            pkh::put_u64(hb, nidx);
            for (size_t q = 0; q < nidx; q++) pkh::put_u64(hb, pre[q]);
            pkh::put_u64(hb, nidx);
            size_t pos = 0;
            for (size_t q = 0; q < nidx; q++) {
                pkh::put_u64(hb, slen[q]);
                for (uint64_t d = 0; d < slen[q]; d++) pkh::put_canonical(hb, suf[pos++].l);
            }
            pkh::put_u64(hb, nidx);
            for (size_t q = 0; q < nidx; q++) pkh::put_u64(hb, idx[q]);
            fs.hint(hb);
        }
        T()[5] += now_s() - t_open;
        if (!is_final) {
            Fr gam, gp = pkh::ONE;
            fs.challenge_scalars(&gam, 1);
            // the fold by the last challenge must land before new equality weights are added
            if (pending) {
                Fr dummy[3];
                PK_TRY(pk_whir_sumcheck_round(ctx, *Pb[which], *Wb[which], *Pb[1 - which], *Wb[1 - which], cur_log, pending_r.l,
                                              dummy[0].l));
                which = 1 - which;
                cur_log--;
                pending = false;
            }
            // new equality constraints, scaled by 1, gam, gam^2, ..: the OOD point, then the STIR points
            // z_q = (domain_gen^16)^idx[q] = omega_D^idx[q] on the folded domain D = 2^(domain_log - 4)
            // (recursive-verifier/app/circuit/whir.go:139-142).  The STIR batch goes through the sparse-DFT form.
            std::vector<Fr> pts((size_t)nvp), sc(nidx + 1);
            expand_from_univariate(ood_pt, nvp, pts.data());
            sc[0] = gp;
            for (size_t q = 0; q < nidx; q++) {
                gp = pkh::mul(gp, gam);
                sc[q + 1] = gp;
            }
            // (the folded STIR values only enter the verifier's claimed sum, not the prover's messages)
            PK_TRY(pk_eval_eq(ctx, pts[0].l, nvp, sc[0].l, *Wb[which]));
            PK_TRY(pk_eval_eq_roots_batch(ctx, idx.data(), nidx, domain_log - FOLD, nvp, sc[1].l, *Wb[which]));
            PK_TRY(whir_sumcheck_rounds(Pb, Wb, &which, &cur_log, FOLD, fold_r, &pending, &pending_r));
            all_r.insert(all_r.end(), fold_r, fold_r + FOLD);
        } else {
            Fr fr_r[FOLD];
            PK_TRY(whir_sumcheck_rounds(Pb, Wb, &which, &cur_log, cfg.final_sumcheck_rounds, fr_r, &pending, &pending_r));
            all_r.insert(all_r.end(), fr_r, fr_r + cfg.final_sumcheck_rounds);
        }
        coeffs = std::move(folded);
        cur_coeffs = *coeffs;
        nv = nvp;
        if (!is_final) {
            prev_owned = std::move(next);
            prev = prev_owned->c;
            domain_log -= 1;
        }
    }
    // deferred weight evaluations at the reversed folding randomness
    std::vector<Fr> R(n, pkh::ZERO);
    for (int i = 0; i < n && i < (int)all_r.size(); i++) R[i] = all_r[all_r.size() - 1 - i];
    std::vector<uint8_t> hb;
    pkh::put_u64(hb, (uint64_t)n_weights);
    Fr dv[3];
    PK_TRY(pk_mle_eval_batch_prefix(ctx, weights, n_weights, n, weights_len, R[0].l, dv[0].l));
    for (int j = 0; j < n_weights; j++) pkh::put_fr(hb, dv[j]);
    fs.hint(hb);
    return PK_OK;
}

int Prover::run() {
    const double t_start = now_s();
    const int m = P->m, m0 = P->m0, mh = P->mh;
    const size_t N = (size_t)1 << m, N0 = (size_t)1 << m0;
    const size_t nw = P->num_witnesses, nc = P->num_constraints;
    // [f || mask] and g in evaluation form were staged by pk_prover_upload_inputs
    pk_buf *masked_w = P->masked_w, *g_w = P->g_w, *masked_h = P->masked_h, *g_h = P->g_h;
    Commitment cmw(ctx);
    PK_TRY(batch_commit(m, masked_w, g_w, &cmw, true));

    // ---- zk-sumcheck (run_zk_sumcheck_prover) ----
    std::vector<Fr> r(m0), alpha(m0);
    fs.challenge_scalars(r.data(), m0);
    Buf a(ctx, N0), b(ctx, N0), c(ctx, N0), eq(ctx, N0);
    if (!a.b || !b.b || !c.b || !eq.b) return pk::set_err(ctx, PK_ERR_OOM, "prove: out of device memory");
    double t0 = now_s();
    PK_TRY(pk_buf_zero(ctx, a, 0, N0));
    PK_TRY(pk_buf_zero(ctx, b, 0, N0));
    PK_TRY(pk_buf_zero(ctx, c, 0, N0));
    PK_TRY(pk_buf_zero(ctx, eq, 0, N0));
    // calculate_witness_bounds (sumcheck.rs:181-193): a = A z, b = B z, c = a o b; z = first nw of masked_w
    PK_TRY(spmv(ctx, P->A, P->d_interned, masked_w->d, a.b->d, nc));
    PK_TRY(spmv(ctx, P->B, P->d_interned, masked_w->d, b.b->d, nc));
    ctx->launches += pk::launch_mul(ctx->stream, a.b->d, b.b->d, c.b->d, nc);
    PK_TRY(pk_eval_eq(ctx, r[0].l, m0, pkh::ONE.l, eq));
    T()[6] += now_s() - t0;

    const Fr* blind = P->blind.data();
    const size_t Nh = (size_t)1 << mh;
    Commitment cmh(ctx);
    PK_TRY(batch_commit(mh, masked_h, g_h, &cmh, false));

    Fr c0[4];
    blinding_coeffs_for_round(blind, m0, 0, nullptr, c0);
    Fr sum_g = pkh::add(eval_cubic(c0, pkh::ZERO), eval_cubic(c0, pkh::ONE));  // sum_over_hypercube
    fs.add_scalars(&sum_g, 1);
    Fr rho;
    fs.challenge_scalars(&rho, 1);
    Fr saved = pkh::mul(rho, sum_g);
    const Fr HALF = pkh::half();
    int cur = m0;
    t0 = now_s();
    for (int idx = 0; idx < m0; idx++) {
        Fr h3[3];
        PK_TRY(pk_zk_sumcheck_round(ctx, a, b, c, eq, cur, idx ? alpha[idx - 1].l : nullptr, h3[0].l));
        if (idx) cur--;
        Fr gp[4], cf[4];
        blinding_coeffs_for_round(blind, m0, idx, alpha.data(), gp);
        cf[0] = pkh::add(h3[0], pkh::mul(rho, gp[0]));
        Fr g_m1 = pkh::sub(pkh::add(pkh::sub(gp[0], gp[1]), gp[2]), gp[3]);
        Fr c_m1 = pkh::add(h3[1], pkh::mul(rho, g_m1));
        cf[2] = pkh::mul(HALF, pkh::sub(pkh::sub(pkh::sub(pkh::add(saved, c_m1), cf[0]), cf[0]), cf[0]));
        cf[3] = pkh::add(h3[2], pkh::mul(rho, gp[3]));
        cf[1] = pkh::sub(pkh::sub(pkh::sub(pkh::sub(saved, cf[0]), cf[0]), cf[3]), cf[2]);
        // whir_r1cs.rs:326-333: the round identity the reference asserts
        Fr chk = pkh::add(pkh::add(pkh::add(pkh::add(cf[0], cf[0]), cf[1]), cf[2]), cf[3]);
        if (chk != saved) return pk::set_err(ctx, PK_ERR_INTERNAL, "zk-sumcheck round %d identity failed", idx);
        fs.add_scalars(cf, 4);
        fs.challenge_scalars(&alpha[idx], 1);
        saved = eval_cubic(cf, alpha[idx]);
    }
    T()[2] += now_s() - t0;

    {   // statement over the blinding commitment: weight = expand_powers(alpha), zero-extended
        std::vector<Fr> wt(Nh, pkh::ZERO);
        for (int i = 0; i < m0; i++) {
            wt[4 * i] = pkh::ONE;
            wt[4 * i + 1] = alpha[i];
            wt[4 * i + 2] = pkh::sqr(alpha[i]);
            wt[4 * i + 3] = pkh::mul(wt[4 * i + 2], alpha[i]);
        }
        Buf dw(ctx, Nh);
        if (!dw.b) return pk::set_err(ctx, PK_ERR_OOM, "prove: out of device memory");
        PK_TRY(pk_buf_upload(ctx, dw, 0, wt[0].l, Nh));
        Fr fg[2];
        const pk_buf* wa[1] = {dw};
        const pk_buf* fb[2] = {masked_h, g_h};
        PK_TRY(pk_multi_dot(ctx, wa, 1, fb, 2, Nh, fg[0].l));
        Fr stmt = pkh::add(fg[0], pkh::mul(cmh.batching, fg[1]));
        fs.add_scalars(fg, 2);
        pk_buf* ws[1] = {dw};
        PK_TRY(whir_prove(P->ch, &cmh, ws, &stmt, 1, 0));
    }

    // ---- weights from the R1CS instance: eq(alpha)^T * {A,B,C}, zero-extended to 2^m ----
    t0 = now_s();
    Buf eq_alpha(ctx, N0);
    BufP wts[3] = {BufP(new Buf(ctx, N)), BufP(new Buf(ctx, N)), BufP(new Buf(ctx, N))};
    if (!eq_alpha.b || !wts[0]->b || !wts[1]->b || !wts[2]->b) return pk::set_err(ctx, PK_ERR_OOM, "prove: out of device memory");
    PK_TRY(pk_buf_zero(ctx, eq_alpha, 0, N0));
    PK_TRY(pk_eval_eq(ctx, alpha[0].l, m0, pkh::ONE.l, eq_alpha));
    const DevCsr* T3[3] = {&P->At, &P->Bt, &P->Ct};
    Fr f_sums[3], g_sums[3], stmts[3], fg6[6];
    for (int j = 0; j < 3; j++) {
        PK_TRY(pk_buf_zero(ctx, *wts[j], 0, N));
        PK_TRY(spmv(ctx, *T3[j], P->d_interned, eq_alpha.b->d, wts[j]->b->d, nw));
    }
    {   // the weights vanish beyond the witness: num_witnesses terms suffice for <w_j, f> and <w_j, g>
        const pk_buf* wa[3] = {*wts[0], *wts[1], *wts[2]};
        const pk_buf* fb[2] = {masked_w, g_w};
        PK_TRY(pk_multi_dot(ctx, wa, 3, fb, 2, nw, fg6[0].l));
    }
    for (int j = 0; j < 3; j++) {
        f_sums[j] = fg6[2 * j];
        g_sums[j] = fg6[2 * j + 1];
        stmts[j] = pkh::add(f_sums[j], pkh::mul(cmw.batching, g_sums[j]));
    }
    T()[6] += now_s() - t0;
    {   // hint claimed_evaluations: (Vec<F>, Vec<F>)
        std::vector<uint8_t> hb;
        pkh::put_u64(hb, 3);
        for (int j = 0; j < 3; j++) pkh::put_fr(hb, f_sums[j]);
        pkh::put_u64(hb, 3);
        for (int j = 0; j < 3; j++) pkh::put_fr(hb, g_sums[j]);
        fs.hint(hb);
    }
    pk_buf* ws[3] = {*wts[0], *wts[1], *wts[2]};
    PK_TRY(whir_prove(P->cw, &cmw, ws, stmts, 3, nw));
    PK_TRY(pk_ctx_sync(ctx));
    if (!fs.ok()) return pk::set_err(ctx, PK_ERR_INVALID_ARG, "prove: a transcript callback reported failure");
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
    *out = p.release();
    return PK_OK;
}
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
// H2D staging of one proof's inputs: witness (zero-padded) || mask, g, blinding cubics || mask, g
int pk_prover_upload_inputs(pk_prover* p, const uint64_t* witness, const pk_rand* rnd) {
    if (!p) return PK_ERR_INVALID_ARG;
    pk_ctx* ctx = p->ctx;
    PK_BIND(ctx);
    PK_CHECK(ctx, witness && rnd, "upload_inputs: null argument");
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
    PK_CUDA(ctx, cudaMemcpyAsync(p->masked_w->d, witness, p->num_witnesses * 32, cudaMemcpyHostToDevice, st));
    PK_CUDA(ctx, cudaMemcpyAsync((char*)p->masked_w->d + half * 32, rnd->mask_w, half * 32, cudaMemcpyHostToDevice, st));
    PK_CUDA(ctx, cudaMemcpyAsync(p->g_w->d, rnd->g_w, N * 32, cudaMemcpyHostToDevice, st));
    PK_CUDA(ctx, cudaMemsetAsync(p->masked_h->d, 0, halfh * 32, st));
    PK_CUDA(ctx, cudaMemcpyAsync(p->masked_h->d, rnd->blind, 4 * (size_t)p->m0 * 32, cudaMemcpyHostToDevice, st));
    PK_CUDA(ctx, cudaMemcpyAsync((char*)p->masked_h->d + halfh * 32, rnd->mask_h, halfh * 32, cudaMemcpyHostToDevice, st));
    PK_CUDA(ctx, cudaMemcpyAsync(p->g_h->d, rnd->g_h, Nh * 32, cudaMemcpyHostToDevice, st));
    p->blind.assign(reinterpret_cast<const Fr*>(rnd->blind), reinterpret_cast<const Fr*>(rnd->blind) + 4 * (size_t)p->m0);
    PK_CUDA(ctx, cudaStreamSynchronize(st));
    std::memset(p->timings, 0, sizeof p->timings);
    p->timings[1] = now_s() - t0;  // H2D staging time
    p->timings[8] = p->timings[1];
    p->staged = true;
    return PK_OK;
}
// same staging with the masks drawn ON THE DEVICE from `seed` (pk_rng_fill streams 0..4 = mask_w, g_w, blind, mask_h,
// g_h): only the witness crosses PCIe.  The reference draws them from thread_rng inside prove (whir_r1cs.rs:211-225).
int pk_prover_upload_inputs_seeded(pk_prover* p, const uint64_t* witness, const uint8_t seed[32]) {
    if (!p) return PK_ERR_INVALID_ARG;
    pk_ctx* ctx = p->ctx;
    PK_BIND(ctx);
    PK_CHECK(ctx, witness && seed, "upload_inputs_seeded: null argument");
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
    p->blind.resize(nb);
    PK_CUDA(ctx, cudaMemcpyAsync(p->blind.data(), p->masked_h->d, nb * 32, cudaMemcpyDeviceToHost, st));
    PK_CUDA(ctx, cudaStreamSynchronize(st));
    std::memset(p->timings, 0, sizeof p->timings);
    p->timings[1] = now_s() - t0;
    p->timings[8] = p->timings[1];
    p->staged = true;
    return PK_OK;
}
int pk_prove_staged(pk_prover* p, uint8_t** out, size_t* out_len) {
    if (!p) return PK_ERR_INVALID_ARG;
    pk_ctx* ctx = p->ctx;
    PK_BIND(ctx);
    PK_CHECK(ctx, out && out_len, "prove: null argument");
    PK_CHECK(ctx, p->staged, "prove_staged: call pk_prover_upload_inputs first");
    double keep1 = p->timings[1];
    std::memset(p->timings, 0, sizeof p->timings);
    p->timings[1] = keep1;
    p->timings[8] = keep1;
    pkh::ProverState fs(p->domsep);
    Prover pr(p, fs);
    PK_TRY(pr.run());
    std::vector<uint8_t>& narg = fs.narg();
    uint8_t* buf = (uint8_t*)std::malloc(narg.size() ? narg.size() : 1);
    if (!buf) return pk::set_err(ctx, PK_ERR_OOM, "prove: host allocation failed");
    std::memcpy(buf, narg.data(), narg.size());
    *out = buf;
    *out_len = narg.size();
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
    Prover pr(p, fs);
    return pr.run();
}
int pk_prove_with_transcript(pk_prover* p, const uint64_t* witness, const pk_rand* rnd, const pk_transcript_vtbl* vt, void* user) {
    PK_TRY(pk_prover_upload_inputs(p, witness, rnd));
    return pk_prove_staged_with_transcript(p, vt, user);
}
int pk_prove(pk_prover* p, const uint64_t* witness, const pk_rand* rnd, uint8_t** out, size_t* out_len) {
    PK_TRY(pk_prover_upload_inputs(p, witness, rnd));
    int rc = pk_prove_staged(p, out, out_len);
    return rc;
}
int pk_prove_seeded(pk_prover* p, const uint64_t* witness, const uint8_t seed[32], uint8_t** out, size_t* out_len) {
    PK_TRY(pk_prover_upload_inputs_seeded(p, witness, seed));
    return pk_prove_staged(p, out, out_len);
}
void pk_free(void* p) { std::free(p); }
void pk_prover_timings(const pk_prover* p, double out[9]) {
    for (int i = 0; i < 9; i++) out[i] = p ? p->timings[i] : 0.0;
}

}  // extern "C"
