/* oracle/kernels.c — CPU restatement of the data-parallel stages (K0..K8 of SURVEY §3).
 * TEST INFRASTRUCTURE ONLY — see oracle/pk_oracle.h.
 */
#include <stdlib.h>

#include "fr.h"
#include "pk_oracle.h"

fr_t orc_compress_fr(fr_t l, fr_t r); /* skyscraper.c */

/* OpenMP team size of every parallel region of the oracle (bench.py --impl reference: torchrun exports OMP_NUM_THREADS=1,
 * which silently made the CPU arm single-threaded).  n <= 0 leaves the setting alone.  Returns the team size in force. */
#ifdef _OPENMP
#include <omp.h>
int orc_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#else
int orc_set_threads(int n) {
    (void)n;
    return 1;
}
#endif

static inline fr_t ld(const uint64_t *p, size_t i) {
    fr_t x;
    memcpy(x.l, p + 4 * i, 32);
    return x;
}
static inline void st(uint64_t *p, size_t i, fr_t x) { memcpy(p + 4 * i, x.l, 32); }

void orc_to_montgomery(const uint64_t *canon, uint64_t *mont, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) st(mont, i, fr_from_canonical(canon + 4 * i));
}
void orc_from_montgomery(const uint64_t *mont, uint64_t *canon, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) fr_to_canonical(ld(mont, i), canon + 4 * i);
}
void orc_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) st(out, i, fr_mul(ld(a, i), ld(b, i)));
}

/* [EXT] whir EvaluationsList::to_coeffs / CoefficientList -> EvaluationsList (call sites
 * provekit/prover/src/whir_r1cs.rs:195,198): the subset-sum ("wavelet") transform and its inverse.
 * evals[idx] = sum of coeffs[s] over s whose bits are a subset of idx's bits. */
static void wavelet(uint64_t *a, int log_n, int inverse) {
    size_t n = (size_t)1 << log_n;
    for (size_t h = 1; h < n; h <<= 1) {
#pragma omp parallel for schedule(static)
        for (size_t t = 0; t < n / 2; t++) {
            size_t i = ((t / h) * 2 * h) + (t % h);
            fr_t lo = ld(a, i), hi = ld(a, i + h);
            st(a, i + h, inverse ? fr_sub(hi, lo) : fr_add(hi, lo));
        }
    }
}
void orc_evals_to_coeffs(uint64_t *a, int log_n) { wavelet(a, log_n, 1); }
void orc_coeffs_to_evals(uint64_t *a, int log_n) { wavelet(a, log_n, 0); }

/* in-place radix-2 DIT NTT, natural in / natural out, tw[j] = zeta^j for j < n/2 */
static void ntt_inplace(fr_t *a, int log_n, const fr_t *tw) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0, j = 0; i < n; i++) {
        if (i < j) { fr_t t = a[i]; a[i] = a[j]; a[j] = t; }
        size_t m = n >> 1;
        while (m >= 1 && (j & m)) { j ^= m; m >>= 1; }
        j |= m;
    }
    for (int s = 1; s <= log_n; s++) {
        size_t half = (size_t)1 << (s - 1), step = n >> s;
        for (size_t k = 0; k < n; k += 2 * half)
            for (size_t j = 0; j < half; j++) {
                fr_t u = a[k + j], v = fr_mul(a[k + j + half], tw[j * step]);
                a[k + j] = fr_add(u, v);
                a[k + j + half] = fr_sub(u, v);
            }
    }
}

/* [EXT] whir commit-time RS encoding ("prover helps" coefficient layout), pinned by the
 * reference-produced fixture and the Go verifier (whir.go:139-142,
 * whir_utilities.go:180-186).  f(X) = sum_k X^k f_k(X^w), w = 2^fold; g generates the domain of
 * size D = n << log_inv_rate; leaf i entry k = f_k((g^w)^i), i < D/w.
 * Each column is an M = D/w point evaluation of a polynomial with n' = n/w coefficients; like the
 * reference's expand_from_coeff it is computed as E = M/n' coset NTTs of size n':
 *   f_k(om^(s + E t)) = NTT_{n'}[ c_j om^(s j) ](t),  om = g^w, zeta = om^E. */
void orc_rs_encode(const uint64_t *coeffs, int log_n, int log_inv_rate, int fold, uint64_t *out,
                   size_t leaf_stride, size_t col_offset) {
    int w = 1 << fold;
    int log_np = log_n - fold; /* n' */
    size_t np = (size_t)1 << log_np;
    size_t E = (size_t)1 << log_inv_rate;
    fr_t g = fr_root_of_unity(log_n + log_inv_rate);
    fr_t om = g;
    for (int i = 0; i < fold; i++) om = fr_sqr(om);
    fr_t zeta = om;
    for (int i = 0; i < log_inv_rate; i++) zeta = fr_sqr(zeta);
    fr_t *tw = (fr_t *)malloc(sizeof(fr_t) * (np / 2 + 1));
    tw[0] = FR_ONE;
    for (size_t j = 1; j < np / 2; j++) tw[j] = fr_mul(tw[j - 1], zeta);
    long jobs = (long)(w * E);
#pragma omp parallel
    {
        fr_t *buf = (fr_t *)malloc(sizeof(fr_t) * np);
#pragma omp for schedule(dynamic, 1)
        for (long job = 0; job < jobs; job++) {
            int k = (int)(job % w);
            size_t s = (size_t)(job / w);
            fr_t oms = fr_pow_u64(om, s), acc = FR_ONE;
            for (size_t j = 0; j < np; j++) {
                buf[j] = fr_mul(ld(coeffs, j * w + k), acc);
                acc = fr_mul(acc, oms);
            }
            ntt_inplace(buf, log_np, tw);
            for (size_t t = 0; t < np; t++) st(out, (s + E * t) * leaf_stride + col_offset + k, buf[t]);
        }
        free(buf);
    }
    free(tw);
}

/* UnivarPoly (recursive-verifier/app/utilities/utilities.go:24-38): Horner; the OOD answer of
 * commit_batch [EXT] is the polynomial at (z^(2^(n-1)),...,z^2,z) = the univariate value at z. */
void orc_eval_univariate(const uint64_t *coeffs, size_t n, const uint64_t z[4], uint64_t out[4]) {
    fr_t zz = ld(z, 0);
    /* blocked Horner so the baseline can use all cores: sum_b z^(b*B) * block_b(z) */
    size_t B = n > 4096 ? 4096 : n;
    size_t nb = n / B;
    fr_t *part = (fr_t *)malloc(sizeof(fr_t) * nb);
#pragma omp parallel for schedule(static)
    for (size_t b = 0; b < nb; b++) {
        fr_t acc = FR_ZERO;
        for (size_t i = B; i-- > 0;) acc = fr_add(fr_mul(acc, zz), ld(coeffs, b * B + i));
        part[b] = acc;
    }
    fr_t zB = fr_pow_u64(zz, B), acc = FR_ZERO;
    for (size_t b = nb; b-- > 0;) acc = fr_add(fr_mul(acc, zB), part[b]);
    free(part);
    memcpy(out, acc.l, 32);
}

/* [EXT] CoefficientList::fold, pinned by computeFold = MultivarPoly(leaf, r)
 * (whir_utilities.go:180-186, utilities.go:15-22): block of 2^k consecutive coefficients ->
 * multilinear evaluation where r[j] binds bit j of the in-block index. */
void orc_fold_coeffs(const uint64_t *coeffs, int log_n, const uint64_t *r, int k, uint64_t *out) {
    size_t nout = (size_t)1 << (log_n - k);
    size_t w = (size_t)1 << k;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nout; i++) {
        fr_t tmp[64];
        for (size_t j = 0; j < w; j++) tmp[j] = ld(coeffs, i * w + j);
        size_t len = w;
        for (int v = 0; v < k; v++) { /* bind bit 0 first with r[0] */
            fr_t rv = ld(r, v);
            len >>= 1;
            for (size_t j = 0; j < len; j++)
                tmp[j] = fr_add(tmp[2 * j], fr_mul(rv, tmp[2 * j + 1]));
        }
        st(out, i, tmp[0]);
    }
}

/* eval_eq, provekit/common/src/utils/sumcheck.rs:145-171 (same convention as whir's eval_eq [EXT]):
 * out[idx] += scalar * prod_j (bit_j(idx) ? x_j : 1-x_j), x_0 <-> most significant bit. */
void orc_eval_eq_accumulate(const uint64_t *point, int n, const uint64_t scalar[4], uint64_t *out) {
    size_t N = (size_t)1 << n;
    fr_t *tmp = (fr_t *)malloc(sizeof(fr_t) * N);
    tmp[0] = ld(scalar, 0);
    /* grow from the last variable (LSB) up so that x_0 ends as the MSB */
    size_t len = 1;
    for (int j = n - 1; j >= 0; j--) {
        fr_t x = ld(point, j);
        /* new[idx + len*bit]: bit of variable j sits above the already-placed lower variables */
#pragma omp parallel for schedule(static) if (len > 4096)
        for (size_t i = 0; i < len; i++) {
            fr_t s1 = fr_mul(tmp[i], x);
            tmp[i + len] = s1;
            tmp[i] = fr_sub(tmp[i], s1);
        }
        len <<= 1;
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) st(out, i, fr_add(ld(out, i), tmp[i]));
    free(tmp);
}

/* [EXT] Weights::linear(w).compute(point) / EvaluationsList::evaluate: sum_i evals[i]*eq(point,i);
 * verifier side: matrix_evaluation.go:47-79. */
void orc_mle_eval(const uint64_t *evals, int log_n, const uint64_t *point, uint64_t out[4]) {
    size_t N = (size_t)1 << log_n;
    fr_t *tmp = (fr_t *)malloc(sizeof(fr_t) * N);
    memcpy(tmp, evals, sizeof(fr_t) * N);
    /* bind x_0 (MSB) first: halves */
    for (int j = 0; j < log_n; j++) {
        size_t h = N >> (j + 1);
        fr_t x = ld(point, j);
#pragma omp parallel for schedule(static) if (h > 4096)
        for (size_t i = 0; i < h; i++) tmp[i] = fr_add(tmp[i], fr_mul(x, fr_sub(tmp[i + h], tmp[i])));
    }
    memcpy(out, tmp[0].l, 32);
    free(tmp);
}

/* [EXT] Weights::weighted_sum (call site whir_r1cs.rs:398-399) */
void orc_dot(const uint64_t *a, const uint64_t *b, size_t n, uint64_t out[4]) {
    fr_t acc = FR_ZERO;
#pragma omp parallel
    {
        fr_t loc = FR_ZERO;
#pragma omp for schedule(static) nowait
        for (size_t i = 0; i < n; i++) loc = fr_add(loc, fr_mul(ld(a, i), ld(b, i)));
#pragma omp critical
        acc = fr_add(acc, loc);
    }
    memcpy(out, acc.l, 32);
}
void orc_axpy(uint64_t *y, const uint64_t *x, const uint64_t a[4], size_t n) {
    fr_t s = ld(a, 0);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) st(y, i, fr_add(ld(y, i), fr_mul(s, ld(x, i))));
}

/* Merkle tree: leaf digest = SkyscraperCRH (left fold of compress, provekit/common/src/skyscraper/
 * whir.rs:30-48), inner = SkyscraperTwoToOne (:53-74), tree shape [EXT] ark MerkleTree::new.
 * version 1 uses the stale v1 compression (fixture tests only). */
void orc_sky_compress_v1(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
static fr_t compress_fr_v(fr_t l, fr_t r, int version) {
    if (version != 1) return orc_compress_fr(l, r);
    uint64_t a[4], b[4], h[4];
    fr_to_canonical(l, a);
    fr_to_canonical(r, b);
    orc_sky_compress_v1(a, b, h);
    return fr_from_canonical(h);
}
void orc_merkle_build(const uint64_t *leaves, size_t L, size_t w, uint64_t *nodes, int version) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < L; i++) {
        fr_t d = ld(leaves, i * w);
        for (size_t k = 1; k < w; k++) d = compress_fr_v(d, ld(leaves, i * w + k), version);
        st(nodes, L + i, d);
    }
    for (size_t lvl = L / 2; lvl >= 1; lvl >>= 1) {
#pragma omp parallel for schedule(static) if (lvl > 64)
        for (size_t i = lvl; i < 2 * lvl; i++)
            st(nodes, i, compress_fr_v(ld(nodes, 2 * i), ld(nodes, 2 * i + 1), version));
    }
    memset(nodes, 0, 32);
}

/* sumcheck_fold_map_reduce with the map of run_zk_sumcheck_prover
 * (provekit/common/src/utils/sumcheck.rs:16-104, provekit/prover/src/whir_r1cs.rs:280-292).
 * Pairs element i with i + len/2 (binds the MSB variable).  When fold != NULL the four arrays are
 * first folded in place by x[i] += fold*(x[i+len/2]-x[i]) and logically truncated to len/2
 * (log_n is the length BEFORE folding). out3 = [f(0), f(-1), f(inf)] (Montgomery). */
void orc_zk_sumcheck_round(uint64_t *a, uint64_t *b, uint64_t *c, uint64_t *eq, int log_n,
                           const uint64_t *fold, uint64_t out3[12]) {
    size_t n = (size_t)1 << log_n;
    uint64_t *arr[4] = {a, b, c, eq};
    if (fold) {
        fr_t f = ld(fold, 0);
        size_t h = n / 2;
        for (int k = 0; k < 4; k++) {
            uint64_t *x = arr[k];
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < h; i++) {
                fr_t lo = ld(x, i);
                st(x, i, fr_add(lo, fr_mul(f, fr_sub(ld(x, i + h), lo))));
            }
        }
        n = h;
    }
    size_t h = n / 2;
    fr_t s0 = FR_ZERO, s1 = FR_ZERO, s2 = FR_ZERO;
#pragma omp parallel
    {
        fr_t l0 = FR_ZERO, l1 = FR_ZERO, l2 = FR_ZERO;
#pragma omp for schedule(static) nowait
        for (size_t i = 0; i < h; i++) {
            fr_t a0 = ld(a, i), a1 = ld(a, i + h), b0 = ld(b, i), b1 = ld(b, i + h);
            fr_t c0 = ld(c, i), c1 = ld(c, i + h), e0 = ld(eq, i), e1 = ld(eq, i + h);
            fr_t f0 = fr_mul(e0, fr_sub(fr_mul(a0, b0), c0));
            fr_t am = fr_sub(fr_dbl(a0), a1), bm = fr_sub(fr_dbl(b0), b1);
            fr_t cm = fr_sub(fr_dbl(c0), c1), em = fr_sub(fr_dbl(e0), e1);
            fr_t fm = fr_mul(em, fr_sub(fr_mul(am, bm), cm));
            fr_t fi = fr_mul(fr_mul(fr_sub(e1, e0), fr_sub(a1, a0)), fr_sub(b1, b0));
            l0 = fr_add(l0, f0);
            l1 = fr_add(l1, fm);
            l2 = fr_add(l2, fi);
        }
#pragma omp critical
        {
            s0 = fr_add(s0, l0);
            s1 = fr_add(s1, l1);
            s2 = fr_add(s2, l2);
        }
    }
    st(out3, 0, s0);
    st(out3, 1, s1);
    st(out3, 2, s2);
}

/* [EXT] whir SumcheckSingle::compute_sumcheck_polynomial + compress: pairs adjacent elements
 * (binds the LSB variable), sends h(0), h(1), h(2) (recursive-verifier whir_utilities.go:107-131,
 * app/utilities/utilities.go:148-154).  fold != NULL: first p'[i] = p[2i] + fold*(p[2i+1]-p[2i]),
 * same for w, in place into the first half (log_n = length BEFORE folding). */
void orc_whir_sumcheck_round(uint64_t *p, uint64_t *w, int log_n, const uint64_t *fold,
                             uint64_t out3[12]) {
    size_t n = (size_t)1 << log_n;
    if (fold) {
        fr_t f = ld(fold, 0);
        size_t h = n / 2;
        uint64_t *arr[2] = {p, w};
        for (int k = 0; k < 2; k++) {
            uint64_t *x = arr[k];
            /* in-place safe only sequentially per index order; use a scratch to stay parallel */
            fr_t *tmp = (fr_t *)malloc(sizeof(fr_t) * h);
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < h; i++) {
                fr_t lo = ld(x, 2 * i);
                tmp[i] = fr_add(lo, fr_mul(f, fr_sub(ld(x, 2 * i + 1), lo)));
            }
            memcpy(x, tmp, sizeof(fr_t) * h);
            free(tmp);
        }
        n = h;
    }
    size_t h = n / 2;
    fr_t s0 = FR_ZERO, s1 = FR_ZERO, s2 = FR_ZERO;
#pragma omp parallel
    {
        fr_t l0 = FR_ZERO, l1 = FR_ZERO, l2 = FR_ZERO;
#pragma omp for schedule(static) nowait
        for (size_t i = 0; i < h; i++) {
            fr_t p0 = ld(p, 2 * i), p1 = ld(p, 2 * i + 1), w0 = ld(w, 2 * i), w1 = ld(w, 2 * i + 1);
            l0 = fr_add(l0, fr_mul(p0, w0));
            l1 = fr_add(l1, fr_mul(p1, w1));
            l2 = fr_add(l2, fr_mul(fr_sub(fr_dbl(p1), p0), fr_sub(fr_dbl(w1), w0)));
        }
#pragma omp critical
        {
            s0 = fr_add(s0, l0);
            s1 = fr_add(s1, l1);
            s2 = fr_add(s2, l2);
        }
    }
    st(out3, 0, s0);
    st(out3, 1, s1);
    st(out3, 2, s2);
}
