/* include/pkwhir.h — C-ABI of libpkwhir.so: the B200 (sm_100a) WHIR hot path of `noir-r1cs prove`.
 *
 * The reference (worldfnd/provekit, Rust) has no FFI today; its plug-in seams for this path are Rust
 * traits / fn types.  Each entry point below names the reference interface it replaces.
 * A Rust host binds these with a `extern "C"` block (INTEGRATION.md shows the shim) and keeps its
 * CLI, witness solving, .nps/.np files and the Fiat-Shamir ProverState.
 *
 * Conventions
 *   - Field elements cross the ABI in arkworks' in-memory form: 4 x u64 little-endian limbs,
 *     Montgomery form (R = 2^256), i.e. `FieldElement = ark_ff::Fp<MontBackend<BN254Config,4>,4>`
 *     (provekit/common/src/lib.rs:19) can be passed by pointer without conversion.  Parameters
 *     documented as "canonical" are plain little-endian 256-bit integers (the hash-input / wire form,
 *     provekit/common/src/skyscraper/whir.rs:21-24).
 *   - Every function returns PK_OK (0) or a negative pk_status; pk_last_error(ctx) gives text.
 *     Where the reference panics (assert!/expect, e.g. provekit/common/src/utils/sumcheck.rs:21-24,
 *     provekit/prover/src/whir_r1cs.rs:206) this ABI returns PK_ERR_INVALID_ARG instead of unwinding.
 *   - One pk_ctx per host thread (owns one CUDA stream); calls are synchronous from the caller's
 *     view unless documented otherwise.  No CPU fallback exists: with no usable CUDA device
 *     pk_ctx_create fails with PK_ERR_NO_DEVICE.
 *   - pk_buf is a device-resident array of field elements; "host" pointers are ordinary memory.
 */
#ifndef PKWHIR_H
#define PKWHIR_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PK_OK = 0,
    PK_ERR_INVALID_ARG = -1,   /* the reference's assert!/panic conditions */
    PK_ERR_NO_DEVICE = -2,
    PK_ERR_CUDA = -3,
    PK_ERR_OOM = -4,
    PK_ERR_INTERNAL = -5,
    PK_ERR_EMPTY_INPUT = -6    /* ark Error::IncorrectInputLength(0), skyscraper/whir.rs:47 */
} pk_status;

typedef struct pk_ctx pk_ctx;
typedef struct pk_buf pk_buf;
typedef struct pk_commitment pk_commitment;

/* ---- context ------------------------------------------------------------------------------- */
/* optional, BEFORE the first CUDA call on `device` in this process: host threads waiting for the device sleep
 * (cudaDeviceScheduleBlockingSync) instead of spinning — for many proofs in flight on many GPUs of one box */
int pk_set_blocking_sync(int device, int on);
int pk_ctx_create(int device, pk_ctx **out);
void pk_ctx_destroy(pk_ctx *ctx);
const char *pk_last_error(const pk_ctx *ctx);
const char *pk_version(void);
/* number of CUDA kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t pk_launch_count(const pk_ctx *ctx);
/* raw cudaStream_t of the ctx (for CUDA-event timing by the caller) */
void *pk_ctx_stream(pk_ctx *ctx);
int pk_ctx_sync(pk_ctx *ctx);

/* ---- device buffers of field elements -------------------------------------------------------- */
int pk_buf_alloc(pk_ctx *ctx, size_t n_elems, pk_buf **out);
void pk_buf_free(pk_ctx *ctx, pk_buf *buf);
size_t pk_buf_len(const pk_buf *buf);
void *pk_buf_device_ptr(pk_buf *buf);
int pk_buf_upload(pk_ctx *ctx, pk_buf *dst, size_t dst_off, const uint64_t *host, size_t n_elems);
int pk_buf_download(pk_ctx *ctx, const pk_buf *src, size_t src_off, uint64_t *host, size_t n_elems);
int pk_buf_copy(pk_ctx *ctx, pk_buf *dst, size_t dst_off, const pk_buf *src, size_t src_off, size_t n_elems);
int pk_buf_zero(pk_ctx *ctx, pk_buf *dst, size_t off, size_t n_elems);

/* ---- seam: skyscraper::CompressManyFn ---------------------------------------------------------
 * `pub type CompressManyFn = fn(&[u8], &mut [u8])` (skyscraper/core/src/lib.rs:26; contract
 * generic.rs:14-37): n messages of 64 B (two canonical LE 256-bit integers, any value < 2^256)
 * -> n hashes of 32 B (canonical, < p).  Host buffers. */
int pk_skyscraper_compress_many(pk_ctx *ctx, const uint8_t *messages, uint8_t *hashes, size_t n);
/* same on device-resident data (messages: 2n elems, hashes: n elems, canonical form) */
int pk_skyscraper_compress_many_dev(pk_ctx *ctx, const pk_buf *messages, pk_buf *hashes, size_t n);

/* ---- seam: PowStrategy (provekit/common/src/skyscraper/pow.rs:14-30 -> skyscraper::pow::solve,
 * skyscraper/core/src/pow.rs:33-41): smallest nonce with compress(challenge,[nonce,0,0,0]) <
 * threshold(bits + 0.01).  challenge: canonical 4 x u64.  bits must be in [0, 60). */
int pk_pow_solve(pk_ctx *ctx, const uint64_t challenge[4], double bits, uint64_t *nonce);

/* ---- seam: EvaluationsList::to_coeffs / CoefficientList -> EvaluationsList [whir] -------------
 * call sites provekit/prover/src/whir_r1cs.rs:195,198.  In place on 2^log_n elements. */
int pk_evals_to_coeffs(pk_ctx *ctx, pk_buf *buf, int log_n);
int pk_coeffs_to_evals(pk_ctx *ctx, pk_buf *buf, int log_n);

/* ---- seam: CommitmentWriter::commit_batch [whir], call site whir_r1cs.rs:200-206 -------------
 * RS-encodes `batch` coefficient vectors of 2^log_n on the domain 2^(log_n+log_inv_rate)
 * ("prover helps" leaves of 2^fold per polynomial, stacked per leaf), builds the Skyscraper Merkle
 * tree (SkyscraperMerkleConfig, provekit/common/src/skyscraper/whir.rs:27-86) and returns the root
 * (Montgomery form, = MerkleConfig::InnerDigest).  The commitment keeps leaves + tree on device. */
int pk_commit_batch(pk_ctx *ctx, const pk_buf *const *coeffs, int batch, int log_n, int log_inv_rate,
                    int fold, pk_commitment **out, uint64_t root_out[4]);
void pk_commit_free(pk_ctx *ctx, pk_commitment *c);
size_t pk_commit_num_leaves(const pk_commitment *c);
size_t pk_commit_leaf_width(const pk_commitment *c);
/* the two stages of pk_commit_batch, exposed for measurement: codeword only / tree only */
int pk_rs_encode(pk_ctx *ctx, const pk_buf *coeffs, int log_n, int log_inv_rate, int fold, pk_buf *leaves,
                 size_t leaf_stride, size_t col_offset);
/* nodes: 2L elements, heap order (nodes[1] = root, leaf digests at [L,2L)), canonical digests */
int pk_merkle_build(pk_ctx *ctx, const pk_buf *leaves, size_t num_leaves, size_t leaf_width, pk_buf *nodes);

/* ---- multi-GPU commit (one process per GPU): the codeword's columns are sharded across ranks for the
 * NTT, its rows (Merkle leaves) across ranks for hashing.  The exchange between the two layouts is FUSED into the last
 * NTT pass: every rank stores its columns of row r straight into the leaf block of the rank that owns r, through
 * CUDA-IPC mapped peer memory (NVLink P2P stores) — no staging buffer, no separate all-to-all kernel.  Sub-tree roots are
 * then all-gathered by the caller (NCCL, 32 B per rank) and combined with pk_merkle_combine_roots.
 *   pk_buf_alloc_shared : a leaf block other processes can map;  pk_ipc_export/open/close : 64-byte cudaIpcMemHandle_t
 *   pk_rs_encode_sharded: columns [col_first, col_first+n_cols) of one polynomial -> peer_leaves[row / (rows/n_peers)] */
int pk_buf_alloc_shared(pk_ctx *ctx, size_t n_elems, pk_buf **out);
int pk_ipc_export(pk_ctx *ctx, const pk_buf *buf, uint8_t handle_out[64]);
int pk_ipc_open(pk_ctx *ctx, const uint8_t handle[64], void **dptr_out);
int pk_ipc_close(pk_ctx *ctx, void *dptr);
int pk_rs_encode_sharded(pk_ctx *ctx, const pk_buf *coeffs, int log_n, int log_inv_rate, int fold, int col_first,
                         int n_cols, void *const *peer_leaves, int n_peers, size_t leaf_stride, size_t col_offset);
/* n canonical sub-tree roots (rank order, n a power of two <= 8) -> canonical root of the whole tree */
int pk_merkle_combine_roots(pk_ctx *ctx, const uint64_t *roots, int n, uint64_t root_out[4]);

/* ---- seam: MerkleTree::generate_multi_proof + STIR answers [whir/ark-crypto-primitives] -------
 * sorted_idx: strictly increasing leaf indexes.  leaves_out: n_idx * leaf_width elements (Montgomery).
 * ark MultiPath pieces (canonical 32-byte digests): sibling_out[n_idx*4]; prefix_len_out[n_idx];
 * suffix_out holds the concatenated suffixes (root->leaf order), suffix_len_out[n_idx] their lengths;
 * suffix_cap = capacity of suffix_out in digests (n_idx * depth always suffices). */
int pk_commit_open(pk_ctx *ctx, const pk_commitment *c, const uint64_t *sorted_idx, size_t n_idx,
                   uint64_t *leaves_out, uint64_t *sibling_out, uint64_t *prefix_len_out,
                   uint64_t *suffix_out, uint64_t *suffix_len_out, size_t suffix_cap);

/* the two halves of pk_commit_open, for a SHARDED opening ("pk_commit_open routes each queried index to its owner
 * rank"): every rank opens ITS rows with pk_commit_open_paths on a view of its leaf block and sub-tree (pk_commit_wrap, local
 * row indexes), the uncompressed paths are gathered in rank order, extended by the levels above the sub-trees (computed from
 * the all-gathered sub-roots) and compressed once with pk_multipath_build.  paths[q][level]: level 0 = sibling leaf digest,
 * level depth-1 = the sibling below the root; canonical digests. */
int pk_commit_wrap(pk_ctx *ctx, const pk_buf *leaves, const pk_buf *nodes, size_t num_leaves, size_t leaf_width, pk_commitment **out);
int pk_commit_open_paths(pk_ctx *ctx, const pk_commitment *c, const uint64_t *sorted_idx, size_t n_idx, uint64_t *leaves_out,
                         uint64_t *paths_out);
int pk_multipath_build(pk_ctx *ctx, const uint64_t *paths, size_t n_idx, int depth, uint64_t *sibling_out, uint64_t *prefix_len_out,
                       uint64_t *suffix_out, uint64_t *suffix_len_out, size_t suffix_cap);

/* ---- univariate / multilinear helpers used by commit_batch and Prover::prove [whir] ---------- */
/* OOD answer: coefficient vector evaluated at (z^(2^(n-1)),..,z^2,z) = Horner at z */
int pk_eval_univariate(pk_ctx *ctx, const pk_buf *coeffs, size_t n, const uint64_t z[4], uint64_t out[4]);
/* the same point for k <= 3 polynomials in one pass (commit_batch evaluates every polynomial of the batch) */
int pk_eval_univariate_batch(pk_ctx *ctx, const pk_buf *const *coeffs, int k, size_t n, const uint64_t z[4], uint64_t *out);
/* y[i] += a * x[i]  (batching p0 + b*p1; linear-weight accumulation) */
int pk_axpy(pk_ctx *ctx, pk_buf *y, const pk_buf *x, const uint64_t a[4], size_t n);
/* Weights::weighted_sum: <a, b> over n elements (call site whir_r1cs.rs:398-399) */
int pk_dot(pk_ctx *ctx, const pk_buf *a, const pk_buf *b, size_t n, uint64_t out[4]);
/* out[ja*nb + jb] = <a_ja, b_jb> in ONE pass; shapes (na, nb) = (3, 2) or (1, 2): the f_sums / g_sums of
 * create_combined_statement_over_two_polynomials (whir_r1cs.rs:382-412) */
int pk_multi_dot(pk_ctx *ctx, const pk_buf *const *a, int na, const pk_buf *const *b, int nb, size_t n, uint64_t *out);
/* eval_eq (provekit/common/src/utils/sumcheck.rs:145-171): out[idx] += scalar * eq(point, idx),
 * point[0] <-> most significant index bit; n variables */
int pk_eval_eq(pk_ctx *ctx, const uint64_t *point, int n, const uint64_t scalar[4], pk_buf *out);
/* several points at once: out[idx] += sum_k scalars[k] * eq(points[k], idx)  (STIR/OOD constraints) */
int pk_eval_eq_batch(pk_ctx *ctx, const uint64_t *points, size_t k, int n, const uint64_t *scalars, pk_buf *out);
/* the same sum for UNIVARIATE points on a multiplicative subgroup — the STIR constraints of a WHIR round
 * (recursive-verifier/app/circuit/whir.go:139-142: z_k = expDomainGen^leafIndex, expanded by ExpandFromUnivariate):
 * out[x] += sum_k scalars[k] * eq((z_k^(2^(n-1)), .., z_k^2, z_k), x) with z_k = omega_D^(exps[k]), D = 2^log_d the
 * arkworks 2-adic subgroup.  Large batches are evaluated as M^T(DFT_D(sparse scalars)) in O(D log D) multiplications
 * instead of k * 2^n (exact arithmetic: bit-identical to pk_eval_eq_batch on the expanded points). */
int pk_eval_eq_roots_batch(pk_ctx *ctx, const uint64_t *exps, size_t k, int log_d, int n, const uint64_t *scalars, pk_buf *out);
/* Weights::linear(w).compute(point): sum_idx evals[idx] * eq(point, idx) */
int pk_mle_eval(pk_ctx *ctx, const pk_buf *evals, int log_n, const uint64_t *point, uint64_t out[4]);
/* k <= 3 weight vectors at the same point in one pass (the deferred_weight_evaluations hint) */
int pk_mle_eval_batch(pk_ctx *ctx, const pk_buf *const *evals, int k, int log_n, const uint64_t *point, uint64_t *out);
/* the same for arrays known to vanish beyond their first n_prefix elements (the R1CS weight vectors are zero-extended
 * from #witnesses to 2^m, whir_r1cs.rs:382-412): only the prefix is read */
int pk_mle_eval_batch_prefix(pk_ctx *ctx, const pk_buf *const *evals, int k, int log_n, size_t n_prefix,
                             const uint64_t *point, uint64_t *out);
/* CoefficientList::fold: 2^k consecutive coefficients -> 1; r[j] binds bit j of the in-block index */
int pk_fold_coeffs(pk_ctx *ctx, const pk_buf *coeffs, int log_n, const uint64_t *r, int k, pk_buf *out);

/* ---- seam: sumcheck_fold_map_reduce (provekit/common/src/utils/sumcheck.rs:16-39) -------------
 * with the map of run_zk_sumcheck_prover (provekit/prover/src/whir_r1cs.rs:284-291).  The four
 * arrays hold 2^log_n elements; if fold != NULL they are first folded in place by
 * x[i] += fold*(x[i+len/2]-x[i]) and the caller treats them as 2^(log_n-1) long afterwards
 * (the reference truncates, whir_r1cs.rs:293-298).  out3 = [f(0), f(-1), f(inf)] (Montgomery).
 * Asserts of the reference (power of two, >= 2, >= 4 when folding) -> PK_ERR_INVALID_ARG. */
int pk_zk_sumcheck_round(pk_ctx *ctx, pk_buf *a, pk_buf *b, pk_buf *c, pk_buf *eq, int log_n,
                         const uint64_t *fold_or_null, uint64_t out3[12]);

/* ---- seam: whir SumcheckSingle::compute_sumcheck_polynomial + compress ------------------------
 * LSB pairing (elements 2i, 2i+1).  If fold != NULL: p_out[i] = p_in[2i] + fold*(p_in[2i+1]-p_in[2i])
 * (same for w) over 2^(log_n-1) outputs, and the round polynomial is computed on the folded arrays;
 * otherwise it is computed on p_in/w_in directly (p_out/w_out may be NULL).  out3 = [h(0),h(1),h(2)]. */
int pk_whir_sumcheck_round(pk_ctx *ctx, const pk_buf *p_in, const pk_buf *w_in, pk_buf *p_out, pk_buf *w_out,
                           int log_n, const uint64_t *fold_or_null, uint64_t out3[12]);

/* ---- multi-GPU sumchecks (one process per GPU) --------------------------------------
 * The zk-sumcheck arrays are sharded by the LOW index bits (rank g owns x[(j << log2 G) | g]: the MSB pairs of
 * sumcheck.rs:26-38 stay local), the WHIR arrays by the HIGH bits (contiguous blocks: the LSB pairs stay local).  A
 * sharded round is the single-GPU kernel on the local shard followed by ONE one-warp kernel that stores the three
 * partial sums into every peer's mailbox (CUDA-IPC mapped peer memory, NVLink P2P stores), waits for the peers'
 * stores and adds them up: out3 is the GLOBAL round message on every rank, with no host-side collective.
 *   mailbox: pk_buf_alloc_shared(ctx, pk_shard_mailbox_elems()); mailboxes[r] = rank r's mailbox (own device pointer, or
 *   pk_ipc_open of the peer's handle).  pk_shard_group_set wipes the rank's OWN mailbox and restarts the sequence at 0, so a
 *   mailbox can serve group after group; the caller puts a barrier between group_set and the first round.  After a
 *   timeout the group is disabled on that rank until the next pk_shard_group_set (the sequence numbers are out of step).
 * All ranks must call the sharded rounds in lock step; a peer that never arrives -> PK_ERR_CUDA after ~10 s, not a hang.
 * When a shard is down to two elements the caller gathers the 2G survivors and finishes with the unsharded entry points
 * (host logic: provekit_b200/sharded.py). */
size_t pk_shard_mailbox_elems(void);
int pk_shard_group_set(pk_ctx *ctx, int rank, int world, void *const *mailboxes);
int pk_shard_group_clear(pk_ctx *ctx);
/* The same mailboxes carry the two collectives of the SHARDED COMMITMENT, so that it needs no host collective at all:
 *   pk_shard_barrier   : barrier on the stream (asynchronous): kernels enqueued behind it start once every rank has reached
 *                        its own barrier — between pk_rs_encode_sharded (peer stores into the other ranks' rows) and
 *                        pk_merkle_build of the local rows.
 *   pk_shard_allgather : out[r] = rank r's src[off] (one field element = one sub-tree root per rank, rank order); synchronises.
 * Every rank must issue the same sequence of sharded calls (rounds, barriers, all-gathers). */
int pk_shard_barrier(pk_ctx *ctx);
int pk_shard_allgather(pk_ctx *ctx, const pk_buf *src, size_t off, uint64_t *out);
int pk_zk_sumcheck_round_sharded(pk_ctx *ctx, pk_buf *a, pk_buf *b, pk_buf *c, pk_buf *eq, int log_n,
                                 const uint64_t *fold_or_null, uint64_t out3[12]);
int pk_whir_sumcheck_round_sharded(pk_ctx *ctx, const pk_buf *p_in, const pk_buf *w_in, pk_buf *p_out, pk_buf *w_out,
                                   int log_n, const uint64_t *fold_or_null, uint64_t out3[12]);

/* ---- seam: WhirR1CSProver::prove (provekit/prover/src/whir_r1cs.rs:42-100) --------------------
 * The whole hot path.  Two ways to drive it:
 *   pk_prove_with_transcript  THE DROP-IN: every Fiat-Shamir operation is a callback into the host's own spongefish
 *                             ProverState, so the proof bytes are the reference's by construction.
 *   pk_prove / pk_prove_seeded / *_enqueue   the same flow over this library's IN-TREE restatement of the sponge
 *                             (on the device by default).  spongefish is not vendored in the reference and its challenge
 *                             derivation could not be pinned here (DESIGN.md "parity unpinned"): these entry
 *                             points are for measurement and for self-contained use with this repo's verifier; a reference
 *                             or gnark verifier is only guaranteed to accept proofs made through the callback entry point.
 * R1CS in the reference's interned-CSR form (provekit/common/src/sparse_matrix.rs:19-26, interner.rs). */
typedef struct {
    uint64_t num_rows, num_cols, nnz;
    const uint64_t *row_start; /* num_rows offsets into col/val */
    const uint32_t *col;
    const uint32_t *val;       /* indices into pk_r1cs.interned */
} pk_csr;
typedef struct {
    uint64_t num_constraints, num_witnesses, num_interned;
    const uint64_t *interned;  /* Montgomery field elements */
    pk_csr a, b, c;
} pk_r1cs;
/* The reference draws these from thread_rng (zk_utils.rs:13-22, whir_r1cs.rs:211-225); the host
 * supplies them so that proofs are reproducible. */
typedef struct {
    const uint64_t *mask_w;  /* 2^(m-1) */
    const uint64_t *g_w;     /* 2^m */
    const uint64_t *blind;   /* 4*m_0 cubic coefficients */
    const uint64_t *mask_h;  /* 2^(mh-1) */
    const uint64_t *g_h;     /* 2^mh */
} pk_rand;
typedef struct pk_prover pk_prover;
/* uploads the R1CS once (scheme-time work, like reading the .nps) */
int pk_prover_create(pk_ctx *ctx, const pk_r1cs *r1cs, pk_prover **out);
void pk_prover_destroy(pk_prover *p);
/* scheme shapes (provekit/r1cs-compiler/src/whir_r1cs.rs:15-36): the witness passed to pk_prove must hold exactly
 * r1cs.num_witnesses elements (the reference asserts witness.len() == r1cs.num_witnesses(), whir_r1cs.rs:48-54) and the
 * pk_rand arrays 2^(m-1), 2^m, 4*m_0, 2^(mh-1), 2^mh elements: the library reads exactly that many and cannot check a
 * bare pointer, so callers size their buffers from these numbers */
void pk_prover_shapes(const pk_prover *p, int *m, int *m0, int *mh);
/* seam: `impl Mul<&[FieldElement]> for HydratedSparseMatrix` / `impl Mul<HydratedSparseMatrix> for &[FieldElement]`
 * (provekit/common/src/sparse_matrix.rs:148-184) on the uploaded R1CS; which = 0 A, 1 B, 2 C.  transposed = 0: out = M x
 * (x: num_witnesses -> out: num_constraints; not for C, the path forms c = a o b instead); transposed = 1: out = x^T M
 * (x: num_constraints -> out: num_witnesses), the rows of calculate_external_row_of_r1cs_matrices (sumcheck.rs:207-218). */
int pk_prover_matvec(pk_prover *p, int which, int transposed, const pk_buf *x, pk_buf *out);
/* Where the in-tree sponge of pk_prove / pk_prove_seeded / pk_prove_staged runs.  Default (0): ON THE DEVICE
 * (provekit/common/src/skyscraper/sponge.rs:24-58 restated as device code, csrc/devts.cuh): prover messages are absorbed
 * and challenges squeezed by one-warp kernels, hints are serialised in HBM, and the host only enqueues work and reads the
 * proof string with ONE synchronisation at the end.  1: on the host (every challenge is a host round trip; the mode the
 * callback entry point pk_prove_with_transcript necessarily uses).  Both give the same bytes.  The environment variable
 * PK_HOST_TRANSCRIPT=1 sets the default for new provers.  pk_prover_host_syncs: stream synchronisations of the last proof. */
int pk_prover_set_host_transcript(pk_prover *p, int on);
uint64_t pk_prover_host_syncs(const pk_prover *p);
/* returns the spongefish NARG string (= WhirR1CSProof.transcript); *out is malloc'd, free with pk_free */
int pk_prove(pk_prover *p, const uint64_t *witness, const pk_rand *rnd, uint8_t **out, size_t *out_len);
/* the two halves of pk_prove, exposed so that a caller can keep one proof's inputs resident in HBM:
 * H2D staging of witness + masks, then the proof itself from the staged inputs (repeatable). */
int pk_prover_upload_inputs(pk_prover *p, const uint64_t *witness, const pk_rand *rnd);
int pk_prove_staged(pk_prover *p, uint8_t **out, size_t *out_len);
/* Asynchronous form (device transcript only): *_enqueue returns as soon as the whole proof — input upload, every kernel,
 * the proof string's D2H — is enqueued on the prover's stream; pk_prove_collect waits for it and hands out the proof
 * string.  One host thread can keep many provers (one ctx / stream each) busy this way; the witness / mask arrays must
 * stay valid and unchanged until collect returns.  One proof per prover at a time. */
int pk_prove_enqueue(pk_prover *p, const uint64_t *witness, const pk_rand *rnd);
int pk_prove_seeded_enqueue(pk_prover *p, const uint64_t *witness, const uint8_t seed[32]);
int pk_prove_staged_enqueue(pk_prover *p);
int pk_prove_collect(pk_prover *p, uint8_t **out, size_t *out_len);
/* Masks drawn on the device: the reference draws them inside prove from thread_rng (a ChaCha12 stream; Fp::rand
 * rejection-samples 254-bit strings; provekit/common/src/utils/zk_utils.rs:13-22, provekit/prover/src/whir_r1cs.rs:211-225).
 * pk_rng_fill writes n uniform field elements: element i = first candidate < p among the 256-bit halves (top word
 * masked to 30 bits) of ChaCha12 blocks keyed by `seed` with state words 12..15 = (i lo, i hi, stream, attempt).
 * pk_prove_seeded = pk_prove with streams 0..4 of `seed` as mask_w, g_w, blind, mask_h, g_h: only the witness is
 * copied host->device.  The caller supplies 32 fresh bytes of OS entropy per proof. */
int pk_rng_fill(pk_ctx *ctx, pk_buf *dst, size_t off, size_t n, const uint8_t seed[32], uint32_t stream);
int pk_prover_upload_inputs_seeded(pk_prover *p, const uint64_t *witness, const uint8_t seed[32]);
int pk_prove_seeded(pk_prover *p, const uint64_t *witness, const uint8_t seed[32], uint8_t **out, size_t *out_len);
/* ---- the host's own Fiat-Shamir transcript: WhirR1CSProver::prove with every transcript operation forwarded to the
 * caller.  The table is exactly the spongefish ProverState surface the reference uses on this path
 * (provekit/prover/src/whir_r1cs.rs:240-242 challenge_scalars, :268-272 add_scalars + challenge_scalars, :335-337,
 * :92 hint; [whir] add_digest = add_scalars(1), challenge_pow = challenge_bytes(32) + add_bytes(8 BE nonce),
 * stir queries = challenge_bytes; IO pattern provekit/common/src/whir_r1cs.rs:28-39): the Rust host passes trampolines
 * over its `ProverState<SkyscraperSponge, FieldElement>` and keeps sponge, codecs, domain separator and
 * the proof string (`narg_string()`) on its side, so the bytes are the reference's own by construction.  pk_prove /
 * pk_prove_seeded are this entry point with the in-tree C++ sponge as the transcript.
 * Scalars cross as Montgomery 4 x u64; `hint` receives the ark-serialised payload WITHOUT the u32 length prefix
 * (= `ProverState::hint_bytes`).  Callbacks return 0 on success; any other value aborts the proof with
 * PK_ERR_INVALID_ARG.  They are called on the calling thread, in protocol order. */
typedef struct {
    int (*add_scalars)(void *user, const uint64_t *scalars, size_t n);
    int (*challenge_scalars)(void *user, uint64_t *out, size_t n);
    int (*add_bytes)(void *user, const uint8_t *bytes, size_t n);
    int (*challenge_bytes)(void *user, uint8_t *out, size_t n);
    int (*hint)(void *user, const uint8_t *payload, size_t n);
} pk_transcript_vtbl;
int pk_prove_with_transcript(pk_prover *p, const uint64_t *witness, const pk_rand *rnd, const pk_transcript_vtbl *vt, void *user);
int pk_prove_staged_with_transcript(pk_prover *p, const pk_transcript_vtbl *vt, void *user);
void pk_free(void *p);
/* host wall-clock seconds per stage of the last pk_prove: [0] witness commit (NTT+Merkle), [1] H2D staging of the inputs,
 * [2] zk-sumcheck, [3] WHIR sumcheck rounds, [4] PoW, [5] STIR openings, [6] R1CS mat-vec + weights,
 * [7] everything else, [8] total */
void pk_prover_timings(const pk_prover *p, double out[9]);

/* ---- `.np` proof container (provekit/common/src/file/bin.rs:16-60, file/mod.rs:33-37): 20-byte header
 * (magic, "NPSProof", version 0.0) + zstd(postcard(NoirProof{WhirR1CSProof{transcript}})).  No device needed.
 * Outputs are malloc'd (pk_free).  decode accepts what `noir-r1cs prove` writes; encode writes what `verify` reads. */
int pk_np_encode(const uint8_t *transcript, size_t len, uint8_t **file_out, size_t *file_len);
int pk_np_decode(const uint8_t *file, size_t len, uint8_t **transcript_out, size_t *transcript_len);

/* ---- `.nps` scheme container: zstd(postcard(NoirProofScheme)), format tag "NrProScm"
 * (provekit/common/src/file/mod.rs:27-29, noir_proof_scheme.rs:16-23).  Extracts the R1CS (r1cs.rs:7-13: interner +
 * three interned-CSR matrices, sparse_matrix.rs:10-27) in exactly the form pk_prover_create consumes (constants converted
 * to Montgomery form); the ACIR program in front of it and the witness builders behind it are skipped (witness solving
 * stays on the host, out of scope).  The returned object owns the arrays pk_nps_r1cs points into. */
typedef struct pk_nps pk_nps;
int pk_nps_read_r1cs(const uint8_t *file, size_t len, pk_nps **out);
const pk_r1cs *pk_nps_r1cs(const pk_nps *s);
int64_t pk_nps_num_public_inputs(const pk_nps *s); /* -1 when not decodable */
void pk_nps_free(pk_nps *s);

/* ---- measurement: CUDA-event timing per kernel class on the ctx stream (no reference counterpart).
 * Between begin and end every launch group is bracketed by an event pair.  Classes: 0 RS-encode NTT
 * passes, 1 Merkle leaf hashing, 2 Merkle upper levels, 3 zk-sumcheck rounds, 4 WHIR sumcheck rounds,
 * 5 wavelet, 6 PoW scan, 7 unused.  launches_by_class counts bracketed launch groups; max_ms_by_class is
 * the longest single group of the class (the dominant launch, e.g. the witness-commit leaf hashing). */
int pk_profile_begin(pk_ctx *ctx);
int pk_profile_end(pk_ctx *ctx, double ms_by_class[8], uint64_t launches_by_class[8], double max_ms_by_class[8]);

/* ---- measurement helper (no reference counterpart): `iters` dependent Montgomery multiplications
 * in two chains per thread on n_threads threads; *ms_out = device time.  Gives the modmul/s ceiling
 * of the integer pipe that DESIGN.md quotes next to the HBM roofline. */
int pk_modmul_bench(pk_ctx *ctx, size_t n_threads, int iters, float *ms_out);
int pk_modsqr_bench(pk_ctx *ctx, size_t n_threads, int iters, float *ms_out); /* same with the dedicated squaring */

#ifdef __cplusplus
}
#endif
#endif
