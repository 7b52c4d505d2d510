/* oracle/pk_oracle.h — public surface of the CPU oracle (libpkoracle.so).
 *
 * TEST INFRASTRUCTURE ONLY.  This library restates, on the CPU, the reference algorithms on the
 * `noir-r1cs prove` WHIR hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product (provekit_b200/, libpkwhir.so)
 * never links or calls it.
 *
 * PARITY STATUS: Skyscraper/PoW pinned by the reference KATs (skyscraper/core/src/reference.rs:100-188);
 * Merkle layout, MultiPath encoding, leaf order, RS-encode layout, transcript framing and WHIR
 * parameters pinned by the reference-produced proof fixture (tests/golden/);
 * verifier algebra restated from the in-tree Go verifier (recursive-verifier/app/circuit/ *.go files) AND pinned against
 * the reference-produced proof without any Fiat-Shamir layer (tests/test_fixture_algebra.py: challenges recovered as roots
 * of the sumcheck links; sumcheck message formats, fold order, RS-encode + batch layout on a complete codeword, wavelet /
 * masked-polynomial layout, OOD, batching, the blinding scheme, and the full verifier equations of the blinding WHIR and
 * of the zk-sumcheck all hold for exactly one candidate tuple).
 * Fiat-Shamir challenge DERIVATION (spongefish internals, whir label strings): parity unpinned.
 *
 * Field elements cross this API as arkworks' in-memory form: 4 x u64 little-endian limbs,
 * Montgomery form (R = 2^256), unless a parameter says "canonical".
 */
#ifndef PK_ORACLE_H
#define PK_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* OpenMP team size of the oracle's parallel loops; n <= 0 only queries.  Returns the size in force. */
int orc_set_threads(int n);

/* ---- Skyscraper (canonical LE integers) ---- */
void orc_sky_permute(const uint64_t l[4], const uint64_t r[4], uint64_t lo[4], uint64_t ro[4]);
void orc_sky_compress(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
void orc_sky_compress_v1(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
int orc_sky_compress_many(const uint8_t *messages, uint8_t *hashes, size_t n, int version);
void orc_pow_threshold(double difficulty, uint64_t out[4]);
int orc_pow_verify(const uint64_t challenge[4], double difficulty, uint64_t nonce);
uint64_t orc_pow_solve(const uint64_t challenge[4], double difficulty);

/* ---- field helpers (Montgomery <-> canonical), n elements ---- */
void orc_to_montgomery(const uint64_t *canon, uint64_t *mont, size_t n);
void orc_from_montgomery(const uint64_t *mont, uint64_t *canon, size_t n);
void orc_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);

/* ---- multilinear / univariate ---- */
void orc_evals_to_coeffs(uint64_t *a, int log_n);
void orc_coeffs_to_evals(uint64_t *a, int log_n);
/* RS-encode one polynomial of 2^log_n coefficients onto the domain of size 2^(log_n+log_inv_rate),
 * "prover helps" layout: out[i*leaf_stride + col_offset + k] = f_k((g^(2^fold))^i). */
void orc_rs_encode(const uint64_t *coeffs, int log_n, int log_inv_rate, int fold, uint64_t *out,
                   size_t leaf_stride, size_t col_offset);
void orc_eval_univariate(const uint64_t *coeffs, size_t n, const uint64_t z[4], uint64_t out[4]);
void orc_fold_coeffs(const uint64_t *coeffs, int log_n, const uint64_t *r, int k, uint64_t *out);
void orc_eval_eq_accumulate(const uint64_t *point, int n, const uint64_t scalar[4], uint64_t *out);
void orc_mle_eval(const uint64_t *evals, int log_n, const uint64_t *point, uint64_t out[4]);
void orc_dot(const uint64_t *a, const uint64_t *b, size_t n, uint64_t out[4]);
void orc_axpy(uint64_t *y, const uint64_t *x, const uint64_t a[4], size_t n); /* y += a*x */

/* ---- Merkle (heap order: nodes[1] root, leaf digests nodes[L..2L)) ---- */
void orc_merkle_build(const uint64_t *leaves, size_t L, size_t w, uint64_t *nodes, int version);

/* ---- sumchecks (in place; arrays shrink logically by half when `fold` != NULL) ---- */
void orc_zk_sumcheck_round(uint64_t *a, uint64_t *b, uint64_t *c, uint64_t *eq, int log_n,
                           const uint64_t *fold, uint64_t out3[12]);
void orc_whir_sumcheck_round(uint64_t *p, uint64_t *w, int log_n, const uint64_t *fold,
                             uint64_t out3[12]);

/* ---- mask generator (oracle/rng.c): ChaCha12 counter stream -> uniform Fr, the CPU twin of pk_rng_fill ---- */
void orc_chacha_block(const uint32_t in[16], int rounds, uint32_t out[16]);
void orc_rng_fill(uint64_t *out, size_t n, const uint8_t seed[32], uint32_t stream);
void orc_rng_masks(const uint8_t seed[32], int m, int m0, int mh, uint64_t *mask_w, uint64_t *g_w, uint64_t *blind,
                   uint64_t *mask_h, uint64_t *g_h);

/* ---- full prover / verifier (oracle/prover.c) ---- */
typedef struct {
    uint64_t num_rows, num_cols, nnz;
    const uint64_t *row_start; /* num_rows entries (provekit/common/src/sparse_matrix.rs:19-26) */
    const uint32_t *col;       /* nnz */
    const uint32_t *val;       /* nnz indices into `interned` */
} orc_csr;
typedef struct {
    uint64_t num_constraints, num_witnesses, num_interned;
    const uint64_t *interned; /* Montgomery */
    orc_csr a, b, c;
} orc_r1cs;
/* provekit/common/src/sparse_matrix.rs:148-165 (M x, transposed == 0) and :168-184 (x^T M) */
void orc_r1cs_matvec(const orc_csr *m, const uint64_t *interned, const uint64_t *x, uint64_t *out, int transposed);
/* Randomness the reference draws from thread_rng is an input here. */
typedef struct {
    const uint64_t *mask_w;   /* 2^(m-1) */
    const uint64_t *g_w;      /* 2^m */
    const uint64_t *blind;    /* 4*m_0 cubic coefficients */
    const uint64_t *mask_h;   /* 2^(mh-1) */
    const uint64_t *g_h;      /* 2^mh */
} orc_rand;
/* The spongefish ProverState / VerifierState surface the path uses (provekit/prover/src/whir_r1cs.rs:240-242,268-272,
 * 335-337; provekit/common/src/whir_r1cs.rs:28-39), as callbacks: same shape as pk_transcript_vtbl of include/pkwhir.h
 * (prover entries) plus the three verifier-side reads.  Scalars are Montgomery 4 x u64.  0 = ok. */
typedef struct {
    int (*add_scalars)(void *user, const uint64_t *scalars, size_t n);
    int (*challenge_scalars)(void *user, uint64_t *out, size_t n);
    int (*add_bytes)(void *user, const uint8_t *bytes, size_t n);
    int (*challenge_bytes)(void *user, uint8_t *out, size_t n);
    int (*hint)(void *user, const uint8_t *payload, size_t n);
    /* verifier only */
    int (*next_scalars)(void *user, uint64_t *out, size_t n);
    int (*next_bytes)(void *user, uint8_t *out, size_t n);
    int (*next_hint)(void *user, const uint8_t **payload, size_t *n); /* payload stays valid until the next call */
} orc_transcript_vtbl;
/* returns transcript length written to *out (malloc'd; free with orc_free), <0 on error */
int64_t orc_prove(const orc_r1cs *r1cs, const uint64_t *witness, const orc_rand *rnd, int hash_version,
                  uint8_t **out);
/* 0 = accept; negative = first failed check */
int orc_verify(const orc_r1cs *r1cs, const uint8_t *transcript, size_t len, int hash_version);
/* the same prover / verifier with every transcript operation forwarded to the caller (the proof string lives on the
 * caller's side): 0 = ok / accept */
int orc_prove_with_transcript(const orc_r1cs *r1cs, const uint64_t *witness, const orc_rand *rnd, int hash_version,
                              const orc_transcript_vtbl *vt, void *user);
int orc_verify_with_transcript(const orc_r1cs *r1cs, int hash_version, const orc_transcript_vtbl *vt, void *user);
/* the in-tree sponge as a foreign transcript: state for the scheme's domain separator (prover: proof == NULL;
 * verifier: wraps `proof`), its vtable (user = the state), the proof string accumulated so far */
void *orc_fs_create(uint64_t num_constraints, uint64_t num_witnesses, const uint8_t *proof, size_t proof_len);
const orc_transcript_vtbl *orc_fs_vtbl(void);
const uint8_t *orc_fs_narg(const void *fs, size_t *len);
void orc_fs_free(void *fs);
void orc_free(void *p);
/* stage timers of the last orc_prove on this thread, seconds: [commit_w_ntt, commit_w_merkle,
 * zk_sumcheck, whir_sumcheck, pow, other, total] */
void orc_last_timings(double out[8]);

#ifdef __cplusplus
}
#endif
#endif
