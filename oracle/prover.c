/* oracle/prover.c — CPU restatement of WhirR1CSProver::prove and of the verifier.
 * TEST INFRASTRUCTURE ONLY — see oracle/pk_oracle.h.
 *
 * Prover: provekit/prover/src/whir_r1cs.rs (whole file) + [EXT] whir CommitmentWriter::commit_batch and
 * Prover::prove, whose message order / algebra is restated from the in-tree Go verifier
 * (recursive-verifier/app/circuit/whir.go:51-220, whir_utilities.go, mtUtilities.go, circuit.go:43-82)
 * and from the byte layout of the reference-produced proof.
 * Verifier: provekit/verifier/src/whir_r1cs.rs + the Go circuit (incl. its R1CS matrix-extension
 * check, matrix_evaluation.go:47-79, which the Rust verifier leaves as TODO).
 */
#include <math.h>
#include <stdio.h>
#include <time.h>

#include "transcript.h"

fr_t orc_compress_fr(fr_t l, fr_t r);

#define MAX_ROUNDS 8
#define FOLD 4

typedef struct {
    double pow_bits;
    int num_queries, ood_samples, log_inv_rate, num_variables, domain_log;
} round_cfg;
typedef struct {
    int num_variables, batch_size, starting_log_inv_rate, starting_domain_log, n_rounds, max_pow_bits;
    round_cfg rounds[MAX_ROUNDS];
    int final_queries, final_log_inv_rate, final_sumcheck_rounds, final_domain_log;
    double final_pow_bits;
} whir_cfg;

/* [EXT] WhirConfig::new for ProveKit's fixed choices (provekit/r1cs-compiler/src/whir_r1cs.rs:38-52:
 * ConjectureList, security 128, fold 4, rate 1/2, pow_bits = default_max_pow(n, 1) = n - 2);
 * pinned by both WhirConfigs inside poseidon-1000.nps (tests/test_fixture.py). */
static void whir_cfg_new(whir_cfg *c, int num_variables, int batch_size) {
    memset(c, 0, sizeof *c);
    c->num_variables = num_variables;
    c->batch_size = batch_size;
    c->starting_log_inv_rate = 1;
    c->max_pow_bits = num_variables + 1 - 3;
    int protocol_security = 128 - c->max_pow_bits;
    if (protocol_security < 0) protocol_security = 0;
    c->final_sumcheck_rounds = num_variables % FOLD;
    c->n_rounds = (num_variables - c->final_sumcheck_rounds) / FOLD - 1;
    c->starting_domain_log = num_variables + 1;
    int rate = 1, nv = num_variables - FOLD, dl = c->starting_domain_log;
    for (int r = 0; r < c->n_rounds; r++) {
        int q = (protocol_security + rate - 1) / rate;
        double pb = 128.0 - (double)q * rate;
        c->rounds[r] = (round_cfg){pb > 0 ? pb : 0, q, 1, rate, nv, dl};
        nv -= FOLD;
        rate += FOLD - 1;
        dl -= 1;
    }
    c->final_queries = (protocol_security + rate - 1) / rate;
    double fb = 128.0 - (double)c->final_queries * rate;
    c->final_pow_bits = fb > 0 ? fb : 0;
    c->final_log_inv_rate = rate;
    c->final_domain_log = dl;
}

/* ---- domain separator (provekit/common/src/whir_r1cs.rs:28-39, utils/sumcheck.rs:123-141, [EXT] whir
 * domainsep labels — parity unpinned) ---- */
static int ceil_div(int a, int b) { return (a + b - 1) / b; }
static void ds_pow(bytebuf *b, double bits) {
    if (bits > 0) {
        ds_op(b, 'S', ceil_div(32, 15), "pow-queries");
        ds_op(b, 'A', 8, "pow-nonce");
    }
}
static void ds_commit(bytebuf *b, const whir_cfg *c) {
    ds_op(b, 'A', 1, "merkle_digest");
    ds_op(b, 'S', 1, "ood_query");
    ds_op(b, 'A', (size_t)c->batch_size, "ood_ans");
    if (c->batch_size > 1) ds_op(b, 'S', 1, "batching_randomness");
}
static void ds_sumcheck(bytebuf *b, int rounds) {
    for (int i = 0; i < rounds; i++) {
        ds_op(b, 'A', 3, "sumcheck_poly");
        ds_op(b, 'S', 1, "folding_randomness");
    }
}
static void ds_whir(bytebuf *b, const whir_cfg *c) {
    ds_op(b, 'S', 1, "initial_combination_randomness");
    ds_sumcheck(b, FOLD);
    for (int r = 0; r < c->n_rounds; r++) {
        int nb = ceil_div(c->rounds[r].domain_log - FOLD, 8);
        ds_op(b, 'A', 1, "merkle_digest");
        ds_op(b, 'S', 1, "ood_query");
        ds_op(b, 'A', 1, "ood_ans");
        ds_pow(b, c->rounds[r].pow_bits);
        ds_op(b, 'S', ceil_div(c->rounds[r].num_queries * nb, 15), "stir_queries");
        ds_op(b, 'H', 0, "stir_answers");
        ds_op(b, 'H', 0, "merkle_proof");
        ds_op(b, 'S', 1, "combination_randomness");
        ds_sumcheck(b, FOLD);
    }
    int nb = ceil_div(c->final_domain_log - FOLD, 8);
    ds_op(b, 'A', (size_t)1 << c->final_sumcheck_rounds, "final_coeffs");
    ds_pow(b, c->final_pow_bits);
    ds_op(b, 'S', ceil_div(c->final_queries * nb, 15), "final_queries");
    ds_op(b, 'H', 0, "stir_answers");
    ds_op(b, 'H', 0, "merkle_proof");
    ds_sumcheck(b, c->final_sumcheck_rounds);
    ds_op(b, 'H', 0, "deferred_weight_evaluations");
}
static void build_domsep(bytebuf *b, const whir_cfg *cw, const whir_cfg *ch, int m0) {
    bb_str(b, "\xF0\x9F\x8C\xAA\xEF\xB8\x8F"); /* "🌪️" */
    ds_commit(b, cw);
    ds_op(b, 'S', (size_t)m0, "rand");
    ds_commit(b, ch);
    ds_op(b, 'A', 1, "Sum of G over boolean hypercube");
    ds_op(b, 'S', 1, "Rho");
    for (int i = 0; i < m0; i++) {
        ds_op(b, 'A', 4, "Sumcheck Polynomials");
        ds_op(b, 'S', 1, "Sumcheck Random");
    }
    ds_op(b, 'A', 2, "Polynomial sums");
    ds_whir(b, ch);
    ds_op(b, 'H', 0, "claimed_evaluations");
    ds_whir(b, cw);
}

/* ---- helpers ---- */
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static __thread double T[8];
void orc_last_timings(double out[8]) { memcpy(out, T, sizeof T); }
void orc_free(void *p) { free(p); }

static int next_pow2_log(uint64_t n) { /* provekit/common/src/utils/mod.rs:71-79 */
    int a = 0;
    uint64_t p = 1;
    while (p < n) {
        p <<= 1;
        a++;
    }
    return a;
}
static fr_t *fr_alloc(size_t n) { return (fr_t *)calloc(n ? n : 1, sizeof(fr_t)); }
static fr_t eval_cubic(const fr_t c[4], fr_t x) { /* sumcheck.rs:174-176 */
    return fr_add(c[0], fr_mul(x, fr_add(c[1], fr_mul(x, fr_add(c[2], fr_mul(x, c[3]))))));
}
static void expand_from_univariate(fr_t z, int n, fr_t *out) { /* utilities.go:182-190 */
    fr_t acc = z;
    for (int i = 0; i < n; i++) {
        out[n - 1 - i] = acc;
        acc = fr_sqr(acc);
    }
}
/* MultivarPoly (utilities.go:15-22): vars[j] binds bit j of the coefficient index */
static fr_t multivar_poly(const fr_t *coefs, int k, const fr_t *vars) {
    fr_t tmp[64];
    size_t len = (size_t)1 << k;
    memcpy(tmp, coefs, len * sizeof(fr_t));
    for (int v = 0; v < k; v++) {
        len >>= 1;
        for (size_t j = 0; j < len; j++) tmp[j] = fr_add(tmp[2 * j], fr_mul(vars[v], tmp[2 * j + 1]));
    }
    return tmp[0];
}
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
/* [EXT] get_challenge_stir_queries, restated by whir_utilities.go:48-77 (+ prover-side sort/dedup,
 * pinned by the strictly increasing leaf_indexes of the fixture) */
static size_t stir_queries(fs_state *fs, int domain_log, int num_queries, uint64_t *idx) {
    int folded_log = domain_log - FOLD;
    int nb = ceil_div(folded_log, 8);
    size_t nbytes = (size_t)num_queries * nb;
    uint8_t *bytes = (uint8_t *)malloc(nbytes ? nbytes : 1);
    fs_challenge_bytes(fs, bytes, nbytes);
    for (int i = 0; i < num_queries; i++) {
        uint64_t v = 0;
        for (int j = 0; j < nb; j++) v = (v << 8) | bytes[i * nb + j];
        idx[i] = v & (((uint64_t)1 << folded_log) - 1);
    }
    free(bytes);
    qsort(idx, num_queries, 8, cmp_u64);
    size_t n = 0;
    for (int i = 0; i < num_queries; i++)
        if (n == 0 || idx[n - 1] != idx[i]) idx[n++] = idx[i];
    return n;
}
/* spongefish-pow challenge_pow [EXT] + SkyscraperPoW (provekit/common/src/skyscraper/pow.rs:14-30) */
static void pow_prove(fs_state *fs, double bits) {
    if (bits <= 0) return;
    double t0 = now_s();
    uint8_t ch[32], nb[8];
    uint64_t c[4];
    fs_challenge_bytes(fs, ch, 32);
    memcpy(c, ch, 32);
    uint64_t nonce = orc_pow_solve(c, bits);
    for (int i = 0; i < 8; i++) nb[i] = (uint8_t)(nonce >> (56 - 8 * i));
    fs_add_bytes(fs, nb, 8);
    T[4] += now_s() - t0;
}
static int pow_check(fs_state *fs, double bits) {
    if (bits <= 0) return 1;
    uint8_t ch[32], nb[8];
    uint64_t c[4], nonce = 0;
    fs_challenge_bytes(fs, ch, 32);
    memcpy(c, ch, 32);
    fs_next_bytes(fs, nb, 8);
    for (int i = 0; i < 8; i++) nonce = (nonce << 8) | nb[i];
    return orc_pow_verify(c, bits, nonce);
}

/* ---- commitment ---- */
typedef struct {
    int n, batch;
    fr_t *poly;   /* batched coefficients, 2^n */
    fr_t *leaves; /* L x w */
    fr_t *nodes;  /* 2L, Montgomery digests */
    size_t L, w;
    int domain_log;
    fr_t ood_point, ood_answer /* batched */, batching;
} commitment;
static void commitment_free(commitment *c) {
    free(c->poly);
    free(c->leaves);
    free(c->nodes);
}
static void serialize_multipath(bytebuf *out, const fr_t *nodes, size_t L, const uint64_t *idx, size_t n) {
    int depth = next_pow2_log(L);
    int plen = depth - 1;
    bytebuf suf = {0}, pre = {0};
    bb_u64(out, n);
    for (size_t q = 0; q < n; q++) bb_fr(out, nodes[(L + idx[q]) ^ 1]);
    fr_t *prev = fr_alloc(plen > 0 ? plen : 1), *cur = fr_alloc(plen > 0 ? plen : 1);
    bb_u64(&pre, n);
    bb_u64(&suf, n);
    for (size_t q = 0; q < n; q++) {
        size_t pos = (L + idx[q]) >> 1;
        for (int d = plen - 1; d >= 0; d--) { /* root->leaf order */
            cur[d] = nodes[pos ^ 1];
            pos >>= 1;
        }
        int k = 0;
        if (q > 0)
            while (k < plen && fr_eq(&prev[k], &cur[k])) k++;
        bb_u64(&pre, (uint64_t)k);
        bb_u64(&suf, (uint64_t)(plen - k));
        for (int d = k; d < plen; d++) bb_fr(&suf, cur[d]);
        fr_t *t = prev;
        prev = cur;
        cur = t;
    }
    bb_push(out, pre.p, pre.len);
    bb_push(out, suf.p, suf.len);
    bb_u64(out, n);
    for (size_t q = 0; q < n; q++) bb_u64(out, idx[q]);
    free(prev);
    free(cur);
    free(pre.p);
    free(suf.p);
}

/* [EXT] CommitmentWriter::commit_batch (call site whir_r1cs.rs:200-206; transcript effect pinned by the
 * fixture: root, batch x OOD answers; order pinned by circuit/mtUtilities.go parseBatchedCommitment) */
static void commit_batch(fs_state *fs, fr_t **polys, int batch, int n, int log_inv_rate, int hv, commitment *c,
                         int timed) {
    memset(c, 0, sizeof *c);
    c->n = n;
    c->batch = batch;
    c->domain_log = n + log_inv_rate;
    c->L = (size_t)1 << (c->domain_log - FOLD);
    c->w = (size_t)batch << FOLD;
    c->leaves = fr_alloc(c->L * c->w);
    c->nodes = fr_alloc(2 * c->L);
    double t0 = now_s();
    for (int b = 0; b < batch; b++)
        orc_rs_encode((const uint64_t *)polys[b], n, log_inv_rate, FOLD, (uint64_t *)c->leaves, c->w, (size_t)b << FOLD);
    double t1 = now_s();
    orc_merkle_build((const uint64_t *)c->leaves, c->L, c->w, (uint64_t *)c->nodes, hv);
    double t2 = now_s();
    if (timed) {
        T[0] += t1 - t0;
        T[1] += t2 - t1;
    }
    fs_add_scalars(fs, &c->nodes[1], 1);
    fs_challenge_scalars(fs, &c->ood_point, 1);
    fr_t ans[4];
    for (int b = 0; b < batch; b++)
        orc_eval_univariate((const uint64_t *)polys[b], (size_t)1 << n, c->ood_point.l, ans[b].l);
    fs_add_scalars(fs, ans, batch);
    c->poly = fr_alloc((size_t)1 << n);
    memcpy(c->poly, polys[0], sizeof(fr_t) << n);
    c->ood_answer = ans[0];
    c->batching = FR_ZERO;
    if (batch > 1) {
        fs_challenge_scalars(fs, &c->batching, 1);
        fr_t g = c->batching;
        for (int b = 1; b < batch; b++) {
            orc_axpy((uint64_t *)c->poly, (const uint64_t *)polys[b], g.l, (size_t)1 << n);
            c->ood_answer = fr_add(c->ood_answer, fr_mul(g, ans[b]));
            g = fr_mul(g, c->batching);
        }
    }
}

/* whir sumcheck rounds: sends [h(0),h(1),h(2)], receives the folding challenge (fused fold+next round) */
static void whir_sumcheck_rounds(fs_state *fs, fr_t *P, fr_t *W, int *cur_log, int rounds, fr_t *rs, int *pending,
                                 fr_t *pending_r) {
    double t0 = now_s();
    for (int i = 0; i < rounds; i++) {
        fr_t h[3];
        orc_whir_sumcheck_round((uint64_t *)P, (uint64_t *)W, *cur_log, *pending ? pending_r->l : NULL, (uint64_t *)h);
        if (*pending) (*cur_log)--;
        fs_add_scalars(fs, h, 3);
        fs_challenge_scalars(fs, &rs[i], 1);
        *pending = 1;
        *pending_r = rs[i];
    }
    T[3] += now_s() - t0;
}
static void apply_pending_fold(fr_t *P, fr_t *W, int *cur_log, int *pending, const fr_t *r) {
    if (!*pending) return;
    size_t h = (size_t)1 << (*cur_log - 1);
    for (size_t i = 0; i < h; i++) {
        P[i] = fr_add(P[2 * i], fr_mul(*r, fr_sub(P[2 * i + 1], P[2 * i])));
        W[i] = fr_add(W[2 * i], fr_mul(*r, fr_sub(W[2 * i + 1], W[2 * i])));
    }
    (*cur_log)--;
    *pending = 0;
}

/* [EXT] whir Prover::prove (call site whir_r1cs.rs:431-434) */
static void whir_prove(fs_state *fs, const whir_cfg *cfg, commitment *cm, fr_t **weights, const fr_t *sums, int n_weights,
                       int hv) {
    const int n = cfg->num_variables;
    size_t N = (size_t)1 << n;
    fr_t gamma, g = FR_ONE, sum = FR_ZERO;
    fs_challenge_scalars(fs, &gamma, 1);
    fr_t *W = fr_alloc(N), *P = fr_alloc(N);
    fr_t pt[64];
    /* constraint 0: the OOD evaluation (added in front), then the caller's linear constraints */
    expand_from_univariate(cm->ood_point, n, pt);
    orc_eval_eq_accumulate((const uint64_t *)pt, n, g.l, (uint64_t *)W);
    sum = fr_mul(g, cm->ood_answer);
    for (int j = 0; j < n_weights; j++) {
        g = fr_mul(g, gamma);
        orc_axpy((uint64_t *)W, (const uint64_t *)weights[j], g.l, N);
        sum = fr_add(sum, fr_mul(g, sums[j]));
    }
    memcpy(P, cm->poly, sizeof(fr_t) * N);
    orc_coeffs_to_evals((uint64_t *)P, n);

    fr_t all_r[64];
    int n_r = 0, cur_log = n, pending = 0;
    fr_t pending_r = FR_ZERO, fold_r[FOLD];
    whir_sumcheck_rounds(fs, P, W, &cur_log, FOLD, fold_r, &pending, &pending_r);
    for (int i = 0; i < FOLD; i++) all_r[n_r++] = fold_r[i];

    fr_t *coeffs = cm->poly; /* owned by cm until replaced */
    int nv = n, domain_log = cm->domain_log;
    commitment prev = *cm;
    int prev_owned = 0;
    for (int ri = 0; ri <= cfg->n_rounds; ri++) {
        int nvp = nv - FOLD;
        fr_t *folded = fr_alloc((size_t)1 << nvp);
        orc_fold_coeffs((const uint64_t *)coeffs, nv, (const uint64_t *)fold_r, FOLD, (uint64_t *)folded);
        int is_final = ri == cfg->n_rounds;
        commitment next;
        memset(&next, 0, sizeof next);
        fr_t ood_pt = FR_ZERO, ood_ans = FR_ZERO;
        double pow_bits;
        int nq;
        if (!is_final) {
            const round_cfg *rc = &cfg->rounds[ri];
            int new_dl = domain_log - 1;
            next.n = nvp;
            next.batch = 1;
            next.domain_log = new_dl;
            next.L = (size_t)1 << (new_dl - FOLD);
            next.w = 16;
            next.leaves = fr_alloc(next.L * 16);
            next.nodes = fr_alloc(2 * next.L);
            orc_rs_encode((const uint64_t *)folded, nvp, new_dl - nvp, FOLD, (uint64_t *)next.leaves, 16, 0);
            orc_merkle_build((const uint64_t *)next.leaves, next.L, 16, (uint64_t *)next.nodes, hv);
            fs_add_scalars(fs, &next.nodes[1], 1);
            fs_challenge_scalars(fs, &ood_pt, 1);
            orc_eval_univariate((const uint64_t *)folded, (size_t)1 << nvp, ood_pt.l, ood_ans.l);
            fs_add_scalars(fs, &ood_ans, 1);
            pow_bits = rc->pow_bits;
            nq = rc->num_queries;
        } else {
            fs_add_scalars(fs, folded, (size_t)1 << nvp);
            pow_bits = cfg->final_pow_bits;
            nq = cfg->final_queries;
        }
        pow_prove(fs, pow_bits);
        uint64_t *idx = (uint64_t *)malloc(8 * (size_t)nq);
        size_t nidx = stir_queries(fs, domain_log, nq, idx);
        /* hints: stir_answers (Vec<Vec<F>>), merkle_proof (ark MultiPath) */
        bytebuf hb = {0};
        bb_u64(&hb, nidx);
        for (size_t q = 0; q < nidx; q++) {
            bb_u64(&hb, prev.w);
            for (size_t k = 0; k < prev.w; k++) bb_fr(&hb, prev.leaves[idx[q] * prev.w + k]);
        }
        fs_hint(fs, hb.p, hb.len);
        hb.len = 0;
        serialize_multipath(&hb, prev.nodes, prev.L, idx, nidx);
        fs_hint(fs, hb.p, hb.len);
        free(hb.p);
        if (!is_final) {
            /* new equality constraints: OOD point + STIR points, combination randomness = powers of gamma */
            fr_t gam, gp = FR_ONE;
            fs_challenge_scalars(fs, &gam, 1);
            apply_pending_fold(P, W, &cur_log, &pending, &pending_r);
            fr_t gen = fr_root_of_unity(domain_log);
            for (int i = 0; i < FOLD; i++) gen = fr_sqr(gen);
            fr_t *pts = fr_alloc((nidx + 1) * (size_t)nvp);
            fr_t *sc = fr_alloc(nidx + 1);
            expand_from_univariate(ood_pt, nvp, pts);
            sc[0] = gp;
            sum = fr_add(sum, fr_mul(gp, ood_ans));
            for (size_t q = 0; q < nidx; q++) {
                gp = fr_mul(gp, gam);
                /* collapse the batched first-round leaf with the batching randomness (rlcBatchedLeaves) */
                fr_t leaf[16];
                for (int k = 0; k < 16; k++) {
                    fr_t v = prev.leaves[idx[q] * prev.w + k], bp = prev.batching;
                    for (int b = 1; b < prev.batch; b++) {
                        v = fr_add(v, fr_mul(bp, prev.leaves[idx[q] * prev.w + 16 * b + k]));
                        bp = fr_mul(bp, prev.batching);
                    }
                    leaf[k] = v;
                }
                fr_t ev = multivar_poly(leaf, FOLD, fold_r);
                expand_from_univariate(fr_pow_u64(gen, idx[q]), nvp, pts + (q + 1) * nvp);
                sc[q + 1] = gp;
                sum = fr_add(sum, fr_mul(gp, ev));
            }
            for (size_t q = 0; q <= nidx; q++)
                orc_eval_eq_accumulate((const uint64_t *)(pts + q * nvp), nvp, sc[q].l, (uint64_t *)W);
            free(pts);
            free(sc);
            whir_sumcheck_rounds(fs, P, W, &cur_log, FOLD, fold_r, &pending, &pending_r);
            for (int i = 0; i < FOLD; i++) all_r[n_r++] = fold_r[i];
        } else {
            fr_t fr_r[FOLD];
            if (cfg->final_sumcheck_rounds > 0) apply_pending_fold(P, W, &cur_log, &pending, &pending_r);
            whir_sumcheck_rounds(fs, P, W, &cur_log, cfg->final_sumcheck_rounds, fr_r, &pending, &pending_r);
            for (int i = 0; i < cfg->final_sumcheck_rounds; i++) all_r[n_r++] = fr_r[i];
        }
        free(idx);
        if (prev_owned) commitment_free(&prev);
        if (coeffs != cm->poly) free(coeffs);
        coeffs = folded;
        nv = nvp;
        if (!is_final) {
            prev = next;
            prev_owned = 1;
            domain_log -= 1;
        }
    }
    if (coeffs != cm->poly) free(coeffs);
    /* deferred weight evaluations at the reversed folding randomness */
    fr_t R[64];
    for (int i = 0; i < n; i++) R[i] = i < n_r ? all_r[n_r - 1 - i] : FR_ZERO;
    bytebuf hb = {0};
    bb_u64(&hb, (uint64_t)n_weights);
    for (int j = 0; j < n_weights; j++) {
        fr_t v;
        orc_mle_eval((const uint64_t *)weights[j], n, (const uint64_t *)R, v.l);
        bb_fr(&hb, v);
    }
    fs_hint(fs, hb.p, hb.len);
    free(hb.p);
    free(W);
    free(P);
    (void)sum;
}

/* ---- R1CS helpers (provekit/common/src/sparse_matrix.rs:148-184) ---- */
static void csr_mul_vec(const orc_csr *m, const uint64_t *interned, const fr_t *x, fr_t *out) {
#pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < m->num_rows; r++) {
        uint64_t s = m->row_start[r], e = r + 1 < m->num_rows ? m->row_start[r + 1] : m->nnz;
        fr_t acc = FR_ZERO;
        for (uint64_t k = s; k < e; k++) {
            fr_t v;
            memcpy(v.l, interned + 4 * m->val[k], 32);
            acc = fr_add(acc, fr_mul(v, x[m->col[k]]));
        }
        out[r] = acc;
    }
}
static void vec_mul_csr(const orc_csr *m, const uint64_t *interned, const fr_t *x, fr_t *out /* num_cols, zeroed */) {
    for (uint64_t r = 0; r < m->num_rows; r++) {
        uint64_t s = m->row_start[r], e = r + 1 < m->num_rows ? m->row_start[r + 1] : m->nnz;
        for (uint64_t k = s; k < e; k++) {
            fr_t v;
            memcpy(v.l, interned + 4 * m->val[k], 32);
            out[m->col[k]] = fr_add(out[m->col[k]], fr_mul(v, x[r]));
        }
    }
}
/* exported for the kernel-level SpMV parity tests: out = M x (transposed == 0) or out = x^T M (out zeroed here) */
void orc_r1cs_matvec(const orc_csr *m, const uint64_t *interned, const uint64_t *x, uint64_t *out, int transposed) {
    if (transposed) {
        memset(out, 0, m->num_cols * sizeof(fr_t));
        vec_mul_csr(m, interned, (const fr_t *)x, (fr_t *)out);
    } else {
        csr_mul_vec(m, interned, (const fr_t *)x, (fr_t *)out);
    }
}
/* compute_blinding_coefficients_for_round, whir_r1cs.rs:103-171 */
static void blinding_coeffs_for_round(const fr_t *g /* n x 4 */, int n, int compute_for, const fr_t *alphas, fr_t out[4]) {
    int all_fixed = 0;
    if (compute_for == n) {
        all_fixed = 1;
        compute_for = n - 1;
    }
    fr_t prefix = FR_ZERO, suffix = FR_ZERO;
    for (int i = 0; i < compute_for; i++) prefix = fr_add(prefix, eval_cubic(g + 4 * i, alphas[i]));
    for (int i = compute_for + 1; i < n; i++)
        suffix = fr_add(suffix, fr_add(eval_cubic(g + 4 * i, FR_ZERO), eval_cubic(g + 4 * i, FR_ONE)));
    fr_t pm = FR_ONE;
    for (int i = 0; i < n - 1 - compute_for; i++) pm = fr_dbl(pm);
    fr_t sm = fr_mul(pm, fr_inv(fr_from_u64(2)));
    fr_t cst = fr_add(fr_mul(pm, prefix), fr_mul(sm, suffix));
    const fr_t *c = g + 4 * compute_for;
    fr_t r[4] = {fr_add(fr_mul(pm, c[0]), cst), fr_mul(pm, c[1]), fr_mul(pm, c[2]), fr_mul(pm, c[3])};
    if (all_fixed) {
        out[0] = eval_cubic(r, alphas[compute_for]);
        out[1] = out[2] = out[3] = FR_ZERO;
    } else {
        memcpy(out, r, sizeof r);
    }
}

/* builds [f || mask] evaluations -> coefficient form, and g -> coefficient form; commits the pair
 * (batch_commit_to_polynomial, whir_r1cs.rs:182-209) */
static void batch_commit(fs_state *fs, int m, const fr_t *f_evals /* 2^(m-1) */, const fr_t *mask, const fr_t *g, int hv,
                         commitment *cm, fr_t **masked_evals, fr_t **g_evals, int timed) {
    size_t half = (size_t)1 << (m - 1), N = (size_t)1 << m;
    fr_t *me = fr_alloc(N), *ge = fr_alloc(N);
    memcpy(me, f_evals, half * sizeof(fr_t));
    memcpy(me + half, mask, half * sizeof(fr_t));
    memcpy(ge, g, N * sizeof(fr_t));
    fr_t *mc = fr_alloc(N), *gc = fr_alloc(N);
    memcpy(mc, me, N * sizeof(fr_t));
    memcpy(gc, ge, N * sizeof(fr_t));
    orc_evals_to_coeffs((uint64_t *)mc, m);
    orc_evals_to_coeffs((uint64_t *)gc, m);
    fr_t *polys[2] = {mc, gc};
    commit_batch(fs, polys, 2, m, 1, hv, cm, timed);
    free(mc);
    free(gc);
    *masked_evals = me;
    *g_evals = ge;
}

/* vt == NULL: the in-tree sponge (proof string returned in *out); otherwise the caller's transcript */
static int64_t prove_impl(const orc_r1cs *r1cs, const uint64_t *witness_in, const orc_rand *rnd, int hv, uint8_t **out,
                          const orc_transcript_vtbl *vt, void *user) {
    memset(T, 0, sizeof T);
    double t_start = now_s();
    /* scheme shapes: provekit/r1cs-compiler/src/whir_r1cs.rs:15-36 */
    int m = next_pow2_log(r1cs->num_witnesses) + 1;
    int m0 = next_pow2_log(r1cs->num_constraints);
    if (m0 < 1 || m < FOLD + 1) return -1;
    int mh = next_pow2_log(4 * (uint64_t)m0) + 1;
    whir_cfg cw, ch;
    whir_cfg_new(&cw, m, 2);
    whir_cfg_new(&ch, mh, 2);
    fs_state fs;
    if (vt) {
        fs_init_foreign(&fs, vt, user, 0);
    } else {
        bytebuf ds = {0};
        build_domsep(&ds, &cw, &ch, m0);
        fs_init(&fs, ds.p, ds.len, NULL, 0);
        free(ds.p);
    }

    const fr_t *witness = (const fr_t *)witness_in;
    size_t half = (size_t)1 << (m - 1), N = (size_t)1 << m, N0 = (size_t)1 << m0;
    fr_t *z = fr_alloc(half);
    memcpy(z, witness, r1cs->num_witnesses * sizeof(fr_t));
    commitment cmw;
    fr_t *masked_w, *g_w;
    batch_commit(&fs, m, z, (const fr_t *)rnd->mask_w, (const fr_t *)rnd->g_w, hv, &cmw, &masked_w, &g_w, 1);

    /* zk-sumcheck (run_zk_sumcheck_prover, whir_r1cs.rs:228-369) */
    fr_t *r = fr_alloc(m0);
    fs_challenge_scalars(&fs, r, m0);
    fr_t *a = fr_alloc(N0), *b = fr_alloc(N0), *c = fr_alloc(N0), *eq = fr_alloc(N0);
    double t0 = now_s();
    csr_mul_vec(&r1cs->a, r1cs->interned, witness, a);
    csr_mul_vec(&r1cs->b, r1cs->interned, witness, b);
    orc_fr_mul((const uint64_t *)a, (const uint64_t *)b, (uint64_t *)c, r1cs->num_constraints);
    orc_eval_eq_accumulate((const uint64_t *)r, m0, FR_ONE.l, (uint64_t *)eq);
    T[5] += now_s() - t0;
    const fr_t *blind = (const fr_t *)rnd->blind;
    size_t halfh = (size_t)1 << (mh - 1);
    fr_t *blind_evals = fr_alloc(halfh);
    memcpy(blind_evals, blind, 4 * (size_t)m0 * sizeof(fr_t));
    commitment cmh;
    fr_t *masked_h, *g_h;
    batch_commit(&fs, mh, blind_evals, (const fr_t *)rnd->mask_h, (const fr_t *)rnd->g_h, hv, &cmh, &masked_h, &g_h, 0);
    fr_t c0[4], sum_g;
    blinding_coeffs_for_round(blind, m0, 0, NULL, c0);
    sum_g = fr_add(eval_cubic(c0, FR_ZERO), eval_cubic(c0, FR_ONE)); /* sum_over_hypercube :173-180 */
    fs_add_scalars(&fs, &sum_g, 1);
    fr_t rho;
    fs_challenge_scalars(&fs, &rho, 1);
    fr_t saved = fr_mul(rho, sum_g);
    fr_t *alpha = fr_alloc(m0);
    const fr_t HALF = fr_inv(fr_from_u64(2));
    int cur = m0;
    t0 = now_s();
    for (int idx = 0; idx < m0; idx++) {
        fr_t h3[3];
        orc_zk_sumcheck_round((uint64_t *)a, (uint64_t *)b, (uint64_t *)c, (uint64_t *)eq, cur,
                              idx ? alpha[idx - 1].l : NULL, (uint64_t *)h3);
        if (idx) cur--;
        fr_t gp[4], cf[4];
        blinding_coeffs_for_round(blind, m0, idx, alpha, gp);
        cf[0] = fr_add(h3[0], fr_mul(rho, gp[0]));
        fr_t g_m1 = fr_sub(fr_add(fr_sub(gp[0], gp[1]), gp[2]), gp[3]);
        fr_t c_m1 = fr_add(h3[1], fr_mul(rho, g_m1));
        cf[2] = fr_mul(HALF, fr_sub(fr_sub(fr_sub(fr_add(saved, c_m1), cf[0]), cf[0]), cf[0]));
        cf[3] = fr_add(h3[2], fr_mul(rho, gp[3]));
        cf[1] = fr_sub(fr_sub(fr_sub(fr_sub(saved, cf[0]), cf[0]), cf[3]), cf[2]);
        fs_add_scalars(&fs, cf, 4);
        fs_challenge_scalars(&fs, &alpha[idx], 1);
        saved = eval_cubic(cf, alpha[idx]);
    }
    T[2] += now_s() - t0;
    /* statement over the blinding commitment: weight = expand_powers(alpha) (whir_r1cs.rs:347-377) */
    {
        size_t Nh = (size_t)1 << mh;
        fr_t *wt = fr_alloc(Nh);
        for (int i = 0; i < m0; i++) {
            wt[4 * i] = FR_ONE;
            wt[4 * i + 1] = alpha[i];
            wt[4 * i + 2] = fr_sqr(alpha[i]);
            wt[4 * i + 3] = fr_mul(wt[4 * i + 2], alpha[i]);
        }
        fr_t fs_[2], stmt;
        orc_dot((const uint64_t *)wt, (const uint64_t *)masked_h, Nh, fs_[0].l);
        orc_dot((const uint64_t *)wt, (const uint64_t *)g_h, Nh, fs_[1].l);
        stmt = fr_add(fs_[0], fr_mul(cmh.batching, fs_[1]));
        fs_add_scalars(&fs, fs_, 2);
        fr_t *ws[1] = {wt};
        whir_prove(&fs, &ch, &cmh, ws, &stmt, 1, hv);
        free(wt);
    }
    /* weights from the R1CS instance (calculate_external_row_of_r1cs_matrices, sumcheck.rs:207-218;
     * create_combined_statement_over_two_polynomials, whir_r1cs.rs:382-412) */
    t0 = now_s();
    fr_t *eq_alpha = fr_alloc(N0);
    orc_eval_eq_accumulate((const uint64_t *)alpha, m0, FR_ONE.l, (uint64_t *)eq_alpha);
    fr_t *wts[3], f_sums[3], g_sums[3], stmts[3];
    const orc_csr *mats[3] = {&r1cs->a, &r1cs->b, &r1cs->c};
    for (int j = 0; j < 3; j++) {
        wts[j] = fr_alloc(N);
        vec_mul_csr(mats[j], r1cs->interned, eq_alpha, wts[j]);
        orc_dot((const uint64_t *)wts[j], (const uint64_t *)masked_w, N, f_sums[j].l);
        orc_dot((const uint64_t *)wts[j], (const uint64_t *)g_w, N, g_sums[j].l);
        stmts[j] = fr_add(f_sums[j], fr_mul(cmw.batching, g_sums[j]));
    }
    T[5] += now_s() - t0;
    bytebuf hb = {0};
    bb_u64(&hb, 3);
    for (int j = 0; j < 3; j++) bb_fr(&hb, f_sums[j]);
    bb_u64(&hb, 3);
    for (int j = 0; j < 3; j++) bb_fr(&hb, g_sums[j]);
    fs_hint(&fs, hb.p, hb.len);
    free(hb.p);
    whir_prove(&fs, &cw, &cmw, wts, stmts, 3, hv);

    for (int j = 0; j < 3; j++) free(wts[j]);
    free(eq_alpha);
    free(alpha);
    free(blind_evals);
    free(a);
    free(b);
    free(c);
    free(eq);
    free(r);
    free(z);
    free(masked_w);
    free(g_w);
    free(masked_h);
    free(g_h);
    commitment_free(&cmw);
    commitment_free(&cmh);
    T[6] = now_s() - t_start;
    if (vt) return fs.failed ? -2 : 0;
    *out = fs.narg.p;
    return (int64_t)fs.narg.len;
}
int64_t orc_prove(const orc_r1cs *r1cs, const uint64_t *witness_in, const orc_rand *rnd, int hv, uint8_t **out) {
    return prove_impl(r1cs, witness_in, rnd, hv, out, NULL, NULL);
}
int orc_prove_with_transcript(const orc_r1cs *r1cs, const uint64_t *witness, const orc_rand *rnd, int hv,
                              const orc_transcript_vtbl *vt, void *user) {
    if (!vt || !vt->add_scalars || !vt->challenge_scalars || !vt->add_bytes || !vt->challenge_bytes || !vt->hint) return -1;
    return (int)prove_impl(r1cs, witness, rnd, hv, NULL, vt, user);
}

/* ================================ verifier ================================================= */
typedef struct {
    fr_t root, ood_point, ood_answers[4], batching;
} parsed_commitment;
static void parse_commitment(fs_state *fs, int batch, parsed_commitment *pc) { /* mtUtilities.go:53-82 */
    memset(pc, 0, sizeof *pc);
    fs_next_scalars(fs, &pc->root, 1);
    fs_challenge_scalars(fs, &pc->ood_point, 1);
    fs_next_scalars(fs, pc->ood_answers, batch);
    if (batch > 1) fs_challenge_scalars(fs, &pc->batching, 1);
}
typedef struct {
    const uint8_t *p;
    size_t len, pos;
    int bad;
} rdr;
static uint64_t rd_u64(rdr *r) {
    uint64_t v = 0;
    if (r->pos + 8 > r->len) {
        r->bad = 1;
        return 0;
    }
    memcpy(&v, r->p + r->pos, 8);
    r->pos += 8;
    return v;
}
static fr_t rd_fr(rdr *r) {
    uint64_t c[4] = {0, 0, 0, 0};
    if (r->pos + 32 > r->len) {
        r->bad = 1;
        return FR_ZERO;
    }
    memcpy(c, r->p + r->pos, 32);
    r->pos += 32;
    if (fr_raw_geq_p(c)) r->bad = 1;
    return fr_from_canonical(c);
}
static fr_t compress_v(fr_t l, fr_t r, int hv) {
    if (hv != 1) return orc_compress_fr(l, r);
    uint64_t a[4], b[4], h[4];
    fr_to_canonical(l, a);
    fr_to_canonical(r, b);
    orc_sky_compress_v1(a, b, h);
    return fr_from_canonical(h);
}
/* reads stir_answers + merkle_proof hints, checks the paths against `root` (whir_utilities.go:13-46) and
 * that the opened indexes are exactly the expected sorted/deduped query set; returns leaves (malloc'd) */
static int read_and_check_openings(fs_state *fs, fr_t root, size_t L, size_t w, const uint64_t *idx, size_t nidx, int hv,
                                   fr_t **leaves_out) {
    size_t n1, n2;
    const uint8_t *h1 = fs_next_hint(fs, &n1);
    const uint8_t *h2 = fs_next_hint(fs, &n2);
    if (fs->failed) return -20;
    rdr a = {h1, n1, 0, 0}, p = {h2, n2, 0, 0};
    if (rd_u64(&a) != nidx) return -21;
    fr_t *leaves = fr_alloc(nidx * w);
    for (size_t q = 0; q < nidx; q++) {
        if (rd_u64(&a) != w) {
            free(leaves);
            return -22;
        }
        for (size_t k = 0; k < w; k++) leaves[q * w + k] = rd_fr(&a);
    }
    int depth = next_pow2_log(L), plen = depth - 1;
    int rc = 0;
    if (rd_u64(&p) != nidx) rc = -23;
    fr_t *sib = fr_alloc(nidx);
    for (size_t q = 0; q < nidx && !rc; q++) sib[q] = rd_fr(&p);
    uint64_t *pre = (uint64_t *)calloc(nidx + 1, 8);
    if (!rc && rd_u64(&p) != nidx) rc = -24;
    for (size_t q = 0; q < nidx && !rc; q++) pre[q] = rd_u64(&p);
    if (!rc && rd_u64(&p) != nidx) rc = -25;
    fr_t *path = fr_alloc(plen > 0 ? plen : 1);
    for (size_t q = 0; q < nidx && !rc; q++) {
        uint64_t sl = rd_u64(&p);
        if (pre[q] + sl != (uint64_t)plen || (q == 0 && pre[q] != 0)) {
            rc = -26;
            break;
        }
        for (uint64_t d = pre[q]; d < (uint64_t)plen; d++) path[d] = rd_fr(&p); /* prefix stays from prev */
        /* leaf -> root */
        fr_t h = leaves[q * w];
        for (size_t k = 1; k < w; k++) h = compress_v(h, leaves[q * w + k], hv);
        uint64_t pos = idx[q];
        h = (pos & 1) ? compress_v(sib[q], h, hv) : compress_v(h, sib[q], hv);
        pos >>= 1;
        for (int d = plen - 1; d >= 0; d--) {
            h = (pos & 1) ? compress_v(path[d], h, hv) : compress_v(h, path[d], hv);
            pos >>= 1;
        }
        if (!fr_eq(&h, &root)) rc = -27;
    }
    if (!rc && rd_u64(&p) != nidx) rc = -28;
    for (size_t q = 0; q < nidx && !rc; q++)
        if (rd_u64(&p) != idx[q]) rc = -29;
    if (!rc && (a.bad || p.bad || a.pos != a.len || p.pos != p.len)) rc = -30;
    free(sib);
    free(pre);
    free(path);
    if (rc) {
        free(leaves);
        return rc;
    }
    *leaves_out = leaves;
    return 0;
}
static fr_t eq_outside(const fr_t *a, const fr_t *b, int n) { /* utilities.go:140-146 */
    fr_t acc = FR_ONE;
    for (int i = 0; i < n; i++) {
        fr_t t = fr_add(fr_mul(a[i], b[i]), fr_mul(fr_sub(FR_ONE, a[i]), fr_sub(FR_ONE, b[i])));
        acc = fr_mul(acc, t);
    }
    return acc;
}
static int verify_sumcheck_rounds(fs_state *fs, int rounds, fr_t *last, fr_t *rs) { /* whir_utilities.go:107-131 */
    const fr_t inv2 = fr_inv(fr_from_u64(2));
    for (int i = 0; i < rounds; i++) {
        fr_t h[3];
        fs_next_scalars(fs, h, 3);
        fs_challenge_scalars(fs, &rs[i], 1);
        fr_t s = fr_add(h[0], h[1]);
        if (!fr_eq(&s, last)) return -40;
        /* utilities.go:148-154 */
        fr_t four_h1 = fr_dbl(fr_dbl(h[1])), three_h0 = fr_add(fr_dbl(h[0]), h[0]);
        fr_t b1 = fr_mul(fr_sub(fr_sub(four_h1, h[2]), three_h0), inv2);
        fr_t b2 = fr_mul(fr_add(fr_sub(h[2], fr_dbl(h[1])), h[0]), inv2);
        *last = fr_add(fr_add(fr_mul(fr_mul(rs[i], rs[i]), b2), fr_mul(rs[i], b1)), h[0]);
    }
    return 0;
}
/* RunZKWhir, recursive-verifier/app/circuit/whir.go:51-220.  lin_evals[b][j]: claimed <w_j, poly_b>.
 * On success returns the reversed folding randomness R (n entries) and the deferred values. */
static int whir_verify(fs_state *fs, const whir_cfg *cfg, const parsed_commitment *pc, fr_t lin_evals[][3], int n_lin,
                       int hv, fr_t *R_out, fr_t *deferred_out) {
    const int n = cfg->num_variables, batch = cfg->batch_size;
    fr_t ood0 = pc->ood_answers[0], bp = pc->batching;
    for (int b = 1; b < batch; b++) { /* oodAnswers, mt.go:76-100 */
        ood0 = fr_add(ood0, fr_mul(bp, pc->ood_answers[b]));
        bp = fr_mul(bp, pc->batching);
    }
    fr_t gamma0, init_comb[8], g = FR_ONE, last = FR_ZERO;
    fs_challenge_scalars(fs, &gamma0, 1);
    for (int j = 0; j < 1 + n_lin; j++) {
        init_comb[j] = g;
        g = fr_mul(g, gamma0);
    }
    last = fr_mul(init_comb[0], ood0);
    for (int j = 0; j < n_lin; j++) { /* initialSumcheck, mtUtilities.go:12-51 */
        fr_t s = FR_ZERO, mult = FR_ONE;
        for (int b = 0; b < batch; b++) {
            s = fr_add(s, fr_mul(lin_evals[b][j], mult));
            mult = fr_mul(mult, pc->batching);
        }
        last = fr_add(last, fr_mul(init_comb[1 + j], s));
    }
    fr_t all_r[64], fold_r[FOLD];
    int n_r = 0, rc;
    if ((rc = verify_sumcheck_rounds(fs, FOLD, &last, fold_r))) return rc;
    for (int i = 0; i < FOLD; i++) all_r[n_r++] = fold_r[i];

    /* per-round data for computeWPoly */
    fr_t *round_pts[MAX_ROUNDS];
    fr_t *round_comb[MAX_ROUNDS];
    size_t round_npts[MAX_ROUNDS];
    int n_round_data = 0;
    int domain_log = cfg->starting_domain_log;
    fr_t prev_root = pc->root;
    size_t prev_w = (size_t)batch << FOLD;
    int prev_batch = batch;
    fr_t *computed_fold = NULL;
    size_t n_fold = 0;
    rc = 0;
    for (int ri = 0; ri <= cfg->n_rounds && !rc; ri++) {
        int is_final = ri == cfg->n_rounds;
        fr_t root = FR_ZERO, ood_pt = FR_ZERO, ood_ans = FR_ZERO;
        fr_t final_coeffs[16];
        double pow_bits;
        int nq;
        if (!is_final) {
            fs_next_scalars(fs, &root, 1);
            fs_challenge_scalars(fs, &ood_pt, 1);
            fs_next_scalars(fs, &ood_ans, 1);
            pow_bits = cfg->rounds[ri].pow_bits;
            nq = cfg->rounds[ri].num_queries;
        } else {
            fs_next_scalars(fs, final_coeffs, (size_t)1 << cfg->final_sumcheck_rounds);
            pow_bits = cfg->final_pow_bits;
            nq = cfg->final_queries;
        }
        if (!pow_check(fs, pow_bits)) {
            rc = -50;
            break;
        }
        uint64_t *idx = (uint64_t *)malloc(8 * (size_t)nq);
        size_t nidx = stir_queries(fs, domain_log, nq, idx);
        fr_t *leaves = NULL;
        size_t L = (size_t)1 << (domain_log - FOLD);
        rc = read_and_check_openings(fs, prev_root, L, prev_w, idx, nidx, hv, &leaves);
        if (rc) {
            free(idx);
            break;
        }
        /* fold values of the opened leaves with the last folding randomness (computeFold) */
        fr_t *folds = fr_alloc(nidx);
        for (size_t q = 0; q < nidx; q++) {
            fr_t leaf[16];
            for (int k = 0; k < 16; k++) {
                fr_t v = leaves[q * prev_w + k], bq = pc->batching;
                for (int b = 1; b < prev_batch; b++) { /* rlcBatchedLeaves */
                    v = fr_add(v, fr_mul(bq, leaves[q * prev_w + 16 * b + k]));
                    bq = fr_mul(bq, pc->batching);
                }
                leaf[k] = v;
            }
            folds[q] = multivar_poly(leaf, FOLD, fold_r);
        }
        free(leaves);
        fr_t gen = fr_root_of_unity(domain_log);
        for (int i = 0; i < FOLD; i++) gen = fr_sqr(gen);
        if (!is_final) {
            fr_t gam, gp = FR_ONE;
            fs_challenge_scalars(fs, &gam, 1);
            size_t np = nidx + 1;
            fr_t *pts = fr_alloc(np), *comb = fr_alloc(np);
            pts[0] = ood_pt;
            comb[0] = gp;
            last = fr_add(last, fr_mul(gp, ood_ans)); /* calculateShiftValue */
            for (size_t q = 0; q < nidx; q++) {
                gp = fr_mul(gp, gam);
                pts[q + 1] = fr_pow_u64(gen, idx[q]);
                comb[q + 1] = gp;
                last = fr_add(last, fr_mul(gp, folds[q]));
            }
            round_pts[n_round_data] = pts;
            round_comb[n_round_data] = comb;
            round_npts[n_round_data] = np;
            n_round_data++;
            rc = verify_sumcheck_rounds(fs, FOLD, &last, fold_r);
            for (int i = 0; i < FOLD; i++) all_r[n_r++] = fold_r[i];
            prev_root = root;
            prev_w = 16;
            prev_batch = 1;
            domain_log -= 1;
        } else {
            /* final: folded leaves must equal the final polynomial at the query points */
            size_t nc = (size_t)1 << cfg->final_sumcheck_rounds;
            for (size_t q = 0; q < nidx && !rc; q++) {
                fr_t z = fr_pow_u64(gen, idx[q]), acc = FR_ZERO;
                for (size_t i = nc; i-- > 0;) acc = fr_add(fr_mul(acc, z), final_coeffs[i]);
                if (!fr_eq(&acc, &folds[q])) rc = -51;
            }
            fr_t fin_r[FOLD];
            if (!rc) rc = verify_sumcheck_rounds(fs, cfg->final_sumcheck_rounds, &last, fin_r);
            for (int i = 0; i < cfg->final_sumcheck_rounds; i++) all_r[n_r++] = fin_r[i];
            /* deferred hint */
            size_t dn;
            const uint8_t *dh = fs_next_hint(fs, &dn);
            rdr d = {dh, dn, 0, 0};
            if (!rc && (fs->failed || rd_u64(&d) != (uint64_t)n_lin)) rc = -52;
            for (int j = 0; j < n_lin && !rc; j++) deferred_out[j] = rd_fr(&d);
            if (!rc && d.bad) rc = -52;
            if (!rc) {
                /* computeWPoly, whir_utilities.go:133-166 */
                fr_t R[64];
                for (int i = 0; i < n; i++) R[i] = all_r[n_r - 1 - i];
                fr_t pt[64], value;
                expand_from_univariate(pc->ood_point, n, pt);
                value = fr_mul(init_comb[0], eq_outside(pt, R, n));
                for (int j = 0; j < n_lin; j++) value = fr_add(value, fr_mul(init_comb[1 + j], deferred_out[j]));
                int nvars = n;
                for (int r2 = 0; r2 < n_round_data; r2++) {
                    nvars -= FOLD;
                    for (size_t i = 0; i < round_npts[r2]; i++) {
                        expand_from_univariate(round_pts[r2][i], nvars, pt);
                        value = fr_add(value, fr_mul(round_comb[r2][i], eq_outside(pt, R, nvars)));
                    }
                }
                /* MultivarPoly(finalCoefficients, finalSumcheckRandomness) */
                fr_t fv = multivar_poly(final_coeffs, cfg->final_sumcheck_rounds, fin_r);
                fr_t rhs = fr_mul(value, fv);
                if (!fr_eq(&last, &rhs)) rc = -53;
                memcpy(R_out, R, sizeof(fr_t) * n);
            }
        }
        free(folds);
        free(idx);
        free(computed_fold);
        computed_fold = NULL;
        (void)n_fold;
    }
    for (int i = 0; i < n_round_data; i++) {
        free(round_pts[i]);
        free(round_comb[i]);
    }
    if (!rc && fs->failed) rc = -54;
    return rc;
}

static int verify_impl(const orc_r1cs *r1cs, const uint8_t *transcript, size_t len, int hv, const orc_transcript_vtbl *vt,
                       void *user) {
    int m = next_pow2_log(r1cs->num_witnesses) + 1;
    int m0 = next_pow2_log(r1cs->num_constraints);
    int mh = next_pow2_log(4 * (uint64_t)m0) + 1;
    whir_cfg cw, ch;
    whir_cfg_new(&cw, m, 2);
    whir_cfg_new(&ch, mh, 2);
    fs_state fs;
    if (vt) {
        fs_init_foreign(&fs, vt, user, 1);
    } else {
        bytebuf ds = {0};
        build_domsep(&ds, &cw, &ch, m0);
        fs_init(&fs, ds.p, ds.len, transcript, len);
        free(ds.p);
    }
    /* provekit/verifier/src/whir_r1cs.rs:40-100 and circuit.go:43-82 */
    parsed_commitment pcw, pch;
    parse_commitment(&fs, 2, &pcw);
    fr_t *r = fr_alloc(m0), *alpha = fr_alloc(m0);
    fs_challenge_scalars(&fs, r, m0);
    parse_commitment(&fs, 2, &pch);
    fr_t sum_g, rho;
    fs_next_scalars(&fs, &sum_g, 1);
    fs_challenge_scalars(&fs, &rho, 1);
    fr_t saved = fr_mul(rho, sum_g);
    int rc = 0;
    for (int i = 0; i < m0; i++) {
        fr_t h[4];
        fs_next_scalars(&fs, h, 4);
        fs_challenge_scalars(&fs, &alpha[i], 1);
        fr_t s = fr_add(eval_cubic(h, FR_ZERO), eval_cubic(h, FR_ONE));
        if (!fr_eq(&s, &saved)) {
            rc = -10;
            break;
        }
        saved = eval_cubic(h, alpha[i]);
    }
    fr_t sums[2], Rh[64], Rw[64], def_h[3], def_w[3];
    if (!rc) {
        fs_next_scalars(&fs, sums, 2);
        fr_t lin[2][3] = {{sums[0]}, {sums[1]}};
        rc = whir_verify(&fs, &ch, &pch, lin, 1, hv, Rh, def_h);
    }
    if (!rc) {
        /* the blinding weight is public: check its deferred evaluation ourselves (the Go circuit takes
         * it as HidingSpartanLinearStatementEvaluations; it equals the MLE of expand_powers(alpha) at Rh) */
        size_t Nh = (size_t)1 << mh;
        fr_t *wt = fr_alloc(Nh), v;
        for (int i = 0; i < m0; i++) {
            wt[4 * i] = FR_ONE;
            wt[4 * i + 1] = alpha[i];
            wt[4 * i + 2] = fr_sqr(alpha[i]);
            wt[4 * i + 3] = fr_mul(wt[4 * i + 2], alpha[i]);
        }
        orc_mle_eval((const uint64_t *)wt, mh, (const uint64_t *)Rh, v.l);
        if (!fr_eq(&v, &def_h[0])) rc = -11;
        free(wt);
    }
    fr_t f_at_alpha = fr_sub(saved, fr_mul(rho, sums[0]));
    if (!rc) {
        size_t hn;
        const uint8_t *hp = fs_next_hint(&fs, &hn);
        rdr d = {hp, hn, 0, 0};
        fr_t lin[2][3];
        for (int b = 0; b < 2 && !rc; b++) {
            if (rd_u64(&d) != 3) rc = -12;
            for (int j = 0; j < 3; j++) lin[b][j] = rd_fr(&d);
        }
        if (!rc && (d.bad || fs.failed)) rc = -12;
        if (!rc) rc = whir_verify(&fs, &cw, &pcw, lin, 3, hv, Rw, def_w);
        if (!rc) {
            /* Spartan relation, verifier/src/whir_r1cs.rs:84-96 */
            fr_t lhs = fr_mul(fr_sub(fr_mul(lin[0][0], lin[0][1]), lin[0][2]), eq_outside(r, alpha, m0));
            if (!fr_eq(&lhs, &f_at_alpha)) rc = -13;
        }
        if (!rc) {
            /* evaluateR1CSMatrixExtension, matrix_evaluation.go:47-79 */
            size_t N0 = (size_t)1 << m0, N = (size_t)1 << m;
            fr_t *row = fr_alloc(N0), *col = fr_alloc(N);
            orc_eval_eq_accumulate((const uint64_t *)alpha, m0, FR_ONE.l, (uint64_t *)row);
            orc_eval_eq_accumulate((const uint64_t *)Rw, m, FR_ONE.l, (uint64_t *)col);
            const orc_csr *mats[3] = {&r1cs->a, &r1cs->b, &r1cs->c};
            for (int j = 0; j < 3 && !rc; j++) {
                fr_t acc = FR_ZERO;
                for (uint64_t rr = 0; rr < mats[j]->num_rows; rr++) {
                    uint64_t s = mats[j]->row_start[rr], e = rr + 1 < mats[j]->num_rows ? mats[j]->row_start[rr + 1] : mats[j]->nnz;
                    for (uint64_t k = s; k < e; k++) {
                        fr_t v;
                        memcpy(v.l, r1cs->interned + 4 * mats[j]->val[k], 32);
                        acc = fr_add(acc, fr_mul(v, fr_mul(row[rr], col[mats[j]->col[k]])));
                    }
                }
                if (!fr_eq(&acc, &def_w[j])) rc = -14;
            }
            free(row);
            free(col);
        }
    }
    if (!rc && (fs.failed || (!vt && fs.rd != len))) rc = -15; /* a foreign transcript checks its own end of input */
    free(r);
    free(alpha);
    return rc;
}
/* ---- the in-tree sponge packaged as a foreign transcript (tests: the product driven by the oracle's transcript
 * through pk_prove_with_transcript must reproduce orc_prove byte for byte) ---- */
static int ft_add_scalars(void *u, const uint64_t *x, size_t n) { fs_add_scalars((fs_state *)u, (const fr_t *)x, n); return 0; }
static int ft_challenge_scalars(void *u, uint64_t *o, size_t n) { fs_challenge_scalars((fs_state *)u, (fr_t *)o, n); return 0; }
static int ft_add_bytes(void *u, const uint8_t *b, size_t n) { fs_add_bytes((fs_state *)u, b, n); return 0; }
static int ft_challenge_bytes(void *u, uint8_t *o, size_t n) { fs_challenge_bytes((fs_state *)u, o, n); return 0; }
static int ft_hint(void *u, const uint8_t *b, size_t n) { fs_hint((fs_state *)u, b, n); return 0; }
static int ft_next_scalars(void *u, uint64_t *o, size_t n) { fs_next_scalars((fs_state *)u, (fr_t *)o, n); return ((fs_state *)u)->failed; }
static int ft_next_bytes(void *u, uint8_t *o, size_t n) { fs_next_bytes((fs_state *)u, o, n); return ((fs_state *)u)->failed; }
static int ft_next_hint(void *u, const uint8_t **p, size_t *n) { *p = fs_next_hint((fs_state *)u, n); return ((fs_state *)u)->failed; }
const orc_transcript_vtbl *orc_fs_vtbl(void) {
    static const orc_transcript_vtbl vt = {ft_add_scalars, ft_challenge_scalars, ft_add_bytes, ft_challenge_bytes, ft_hint,
                                           ft_next_scalars, ft_next_bytes, ft_next_hint};
    return &vt;
}
void *orc_fs_create(uint64_t num_constraints, uint64_t num_witnesses, const uint8_t *proof, size_t proof_len) {
    int m = next_pow2_log(num_witnesses) + 1, m0 = next_pow2_log(num_constraints);
    int mh = next_pow2_log(4 * (uint64_t)m0) + 1;
    whir_cfg cw, ch;
    whir_cfg_new(&cw, m, 2);
    whir_cfg_new(&ch, mh, 2);
    bytebuf ds = {0};
    build_domsep(&ds, &cw, &ch, m0);
    fs_state *fs = (fs_state *)malloc(sizeof *fs);
    fs_init(fs, ds.p, ds.len, proof, proof_len);
    free(ds.p);
    return fs;
}
const uint8_t *orc_fs_narg(const void *fs, size_t *len) {
    *len = ((const fs_state *)fs)->narg.len;
    return ((const fs_state *)fs)->narg.p;
}
void orc_fs_free(void *p) {
    fs_state *fs = (fs_state *)p;
    if (fs && !fs->is_verifier) free(fs->narg.p);
    free(fs);
}
int orc_verify(const orc_r1cs *r1cs, const uint8_t *transcript, size_t len, int hv) {
    return verify_impl(r1cs, transcript, len, hv, NULL, NULL);
}
int orc_verify_with_transcript(const orc_r1cs *r1cs, int hv, const orc_transcript_vtbl *vt, void *user) {
    if (!vt || !vt->next_scalars || !vt->challenge_scalars || !vt->next_bytes || !vt->challenge_bytes || !vt->next_hint) return -1;
    return verify_impl(r1cs, NULL, 0, hv, vt, user);
}
