// provekit_b200/csrc/kernels.cuh — launchers of the sm_100a kernels (K0..K10 of SURVEY §3, PoW).
// All pointers are device pointers to 32-byte field elements unless noted.  Every launcher enqueues on
// `st` and returns the number of kernels it launched (for pk_launch_count).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace pk {

struct fr;

// fold parameter for the sumchecks and scalars in general are passed by value as 8 limbs
struct fr_arg {
    uint32_t v[8];
};

int launch_compress_many(cudaStream_t st, const void* msgs, void* hashes, size_t n);
int launch_to_mont(cudaStream_t st, void* data, size_t n, bool to_mont);

// K0 wavelet transforms, in place
int launch_wavelet(cudaStream_t st, void* a, int log_n, bool inverse);
// mode 0 forward, 1 inverse (the evaluations->coefficients map M), 2 M transposed (monomial weights -> eq weights)
int launch_wavelet_mode(cudaStream_t st, void* a, int log_n, int mode);
// pieces of the sparse-DFT eq-weight batch (see kernels.cu): s[exps[i]] += scalars[i];  u = radix-16 recombination of an
// RS-encode (rate 1) leaf array into the plain size-2^log_d transform, first n_out outputs;  y += x
// k_ptr (device) overrides k when non-null
int launch_scatter_add(cudaStream_t st, void* s, const uint64_t* exps, const void* scalars, size_t k, const uint32_t* k_ptr = nullptr);
int launch_dft16_combine(cudaStream_t st, const void* lv, void* u, size_t n_out, int log_d, const void* table, int table_log_m);
int launch_add_inplace(cudaStream_t st, void* y, const void* x, size_t n);

// twiddle table W[e] = omega_M^e (Montgomery), e < M/2, omega_M = arkworks 2-adic root of order M = 2^log_m
// host_pow2: 28 x 8 limbs, omega^(2^b) for b = 0..27 (Montgomery)
int launch_twiddle_table(cudaStream_t st, void* table, int log_m, const uint32_t* host_pow2);

// K1 RS encode: coeffs (2^log_n) -> out[(row)*leaf_stride + col_offset + k]; scratch: 2^(log_n+log_inv_rate)
// table: twiddles for M = 2^(log_n + log_inv_rate - fold), table_log_m >= that (strided use)
// twist_canonical (optional): table of CANONICAL twiddles of the same order; when given, the codeword comes out as
// canonical integers instead of Montgomery-form elements
int launch_rs_encode(cudaStream_t st, const void* coeffs, int log_n, int log_inv_rate, int fold, void* out,
                     size_t leaf_stride, size_t col_offset, void* scratch, const void* table, int table_log_m,
                     const void* twist_canonical = nullptr);

// ---- ntt.cu: the TMA-staged radix-8 RS-encode (2^S points x 4 columns per tile, tile-major scratch between passes) ----
constexpr int NTT8_MAX_S = 9;  // 2^9 points x 128 B = 64 KB tile (+ 16 KB twiddle slice): two CTAs per SM
constexpr int NTT8_MIN_L = 7;  // shorter columns (2^L points) stay on k_ntt_pass
struct NttR8 {
    const fr* in;      // first pass: coefficients (point q = 16 columns, 512 B); later passes: this pass's tile-major scratch
    fr* out;           // the next pass's tile-major scratch, or the leaves (last pass)
    const fr* tw;      // per-pass twiddle slices [Lo][2^S] (launch_ntt_tile_twiddles)
    const fr* Wtwist;  // coset-twist table of order 2^(logM + tbl_shift): Montgomery, or canonical when `canonical`
    int tbl_shift;
    int L, l, S, logE, logM;
    int first, last, canonical;
    int l_next, S_next;
    size_t leaf_stride, col_offset;
};
int launch_ntt_tile_twiddles(cudaStream_t st, void* out, const void* W, int table_log_m, int L, int l, int S);
int launch_ntt_r8_pass(cudaStream_t st, const NttR8& P);
size_t ntt_r8_smem_bytes(int S);
cudaError_t init_ntt_attributes();

// sharded variant: only columns [col0, col0 + 2^nc_log) of the polynomial; row r of the codeword is stored into
// peers[r / (rows / n_peers)] (device pointers of this or of peer GPUs) at its local row index
int launch_rs_encode_cols(cudaStream_t st, const void* coeffs, int log_n, int log_inv_rate, int fold, int col0, int nc_log,
                          void* const* peers, int n_peers, size_t leaf_stride, size_t col_offset, void* scratch,
                          const void* table, int table_log_m, const void* twist_canonical = nullptr);

// K2 Merkle: leaves Montgomery, nodes canonical heap order
// canonical: the leaf elements are canonical integers already (no Montgomery conversion before hashing)
int launch_merkle_leaves(cudaStream_t st, const void* leaves, size_t L, size_t w, void* nodes, bool canonical);
// root_mont_out (device, optional): the root as a Montgomery-form field element (= MerkleConfig::InnerDigest)
int launch_merkle_upper(cudaStream_t st, size_t L, void* nodes, void* root_mont_out = nullptr);

// tensor-product tables of K points with nv_hi + nv_lo variables each (point k at points[k*pt_stride ..]):
//   hi_k[idx] = scale_k * prod_{j < nv_hi} (bit_{nv_hi-1-j}(idx) ? f1 : f0)   over the first nv_hi variables,
//   lo_k[idx] =           prod over the remaining nv_lo variables;            eq: (f0,f1) = (1-x, x);  pow: (1, x)
// so that T_k[idx] = hi_k[idx >> nv_lo] * lo_k[idx & mask].  scales may be null (= 1).  One launch builds both.
// Where the K points come from:
//   TENSOR_PTS_EXPLICIT  points[k * pt_stride + j], j < nv_hi + nv_lo
//   TENSOR_PTS_UNIVARIATE  point k = (z^(2^(n-1)), .., z^2, z) with z = points[k]  (expand_from_univariate)
//   TENSOR_PTS_ROOTS       the same with z = omega_D^(exps[k]), D = 2^log_d, read from the twiddle table W
// count (device, optional): points k >= *count get scale 0, i.e. contribute nothing (batch size known only on the device)
enum { TENSOR_PTS_EXPLICIT = 0, TENSOR_PTS_UNIVARIATE = 1, TENSOR_PTS_ROOTS = 2 };
struct TensorSrc {
    int mode;
    const uint64_t* exps;
    const uint32_t* count;
    const void* W;
    int tbl_shift, log_d;
};
int launch_tensor_tables(cudaStream_t st, const void* points, size_t K, int pt_stride, int nv_hi, int nv_lo,
                         const void* scales, bool eq_mode, void* out_hi, void* out_lo, TensorSrc src = TensorSrc());
// out[idx] += sum_k hi_k[idx >> lo_bits] * lo_k[idx & mask]
int launch_tensor_accumulate(cudaStream_t st, void* out, int log_n, const void* hi, const void* lo, size_t K,
                             int lo_bits);
// result[0] = sum_idx a[idx] * hi[idx >> lo_bits] * lo[idx & mask]   (hi/lo single tables)
int launch_tensor_dot(cudaStream_t st, const void* a, size_t n, const void* hi, const void* lo, int lo_bits,
                      void* partials, void* result);
// result[0] = <a, b>
int launch_dot(cudaStream_t st, const void* a, const void* b, size_t n, void* partials, void* result);
// result[ja*nb + jb] = <a_ja, b_jb>  ((na, nb) in {(3,2), (1,2)});  result[j] = sum a_j[i] * (hi (x) lo)[i], na <= 3
int launch_multi_dot(cudaStream_t st, const void* const* a, int na, const void* const* b, int nb, size_t n, void* partials,
                     void* result);
int launch_multi_tensor_dot(cudaStream_t st, const void* const* a, int na, size_t n, const void* hi, const void* lo,
                            int lo_bits, void* partials, void* result);
// a_ptr (device) overrides the by-value scalar when non-null
int launch_axpy(cudaStream_t st, void* y, const void* x, fr_arg a, const void* a_ptr, size_t n);
// r_dev[j * r_stride] binds bit j of the in-block index (stride -1: challenges stored newest-first)
int launch_fold_coeffs(cudaStream_t st, const void* coeffs, int log_n, const void* r_dev, int r_stride, int k, void* out);

// K5 / K7 sumcheck rounds; result: 3 field elements on device
// fold_ptr (device, Montgomery) overrides the by-value `fold` when non-null: the challenge may never have left the device
int launch_zk_sumcheck_round(cudaStream_t st, void* a, void* b, void* c, void* eq, int log_n, bool has_fold,
                             fr_arg fold, const void* fold_ptr, void* partials, void* result);
int launch_whir_sumcheck_round(cudaStream_t st, const void* p_in, const void* w_in, void* p_out, void* w_out,
                               int log_n, bool has_fold, fr_arg fold, const void* fold_ptr, void* partials, void* result);

// Sharded sumchecks: all-gather + field sum of the three partial sums of a round, fused behind the reduction as
// NVLink peer stores.  Every rank owns a mailbox of SHARD_SLOTS x SHARD_MAX_WORLD cells of 128 B (96 B payload + sequence
// flag) that its peers map through CUDA IPC; one warp publishes result[0..3) into cell [seq % SLOTS][rank] of every peer,
// waits until its own cells carry `seq`, and replaces result[0..3) by the sum over ranks.  status: 0 ok, 1 timed out.
constexpr int SHARD_MAX_WORLD = 8;
constexpr int SHARD_SLOTS = 4;
constexpr int SHARD_CELL_BYTES = 128;
constexpr size_t SHARD_MAILBOX_BYTES = (size_t)SHARD_SLOTS * SHARD_MAX_WORLD * SHARD_CELL_BYTES;
struct ShardGroup {
    uint8_t* mbox[SHARD_MAX_WORLD];
    int rank, world;
};
int launch_shard_exchange(cudaStream_t st, void* result, ShardGroup g, uint32_t seq, uint32_t* status);
// all-gather of 3 elements per rank into gathered[3 * world] + stream barrier across the group (see k_shard_exchange)
int launch_shard_gather(cudaStream_t st, void* payload3, ShardGroup g, uint32_t seq, uint32_t* status, void* gathered);

// interned-CSR sparse matrix x vector (provekit/common/src/sparse_matrix.rs:148-184): out[r] = sum_k
// interned[val[k]] * x[col[k]] over row r.  The transposed product uses the same kernel on the CSC arrays.
// Rows longer than SPMV_LONG_ROW are skipped by launch_spmv and handled by launch_spmv_long, which splits them
// into chunks of <= SPMV_CHUNK entries (one block per chunk, then one thread per long row sums its chunks):
// a constant-one witness column makes one row of the transposed matrix hold millions of entries.
constexpr int SPMV_LONG_ROW = 64;
constexpr int SPMV_CHUNK = 2048;
int launch_spmv(cudaStream_t st, const uint64_t* row_start, const uint32_t* col, const uint32_t* val,
                const void* interned, const void* x, void* out, size_t num_rows, size_t nnz);
int launch_spmv_long(cudaStream_t st, const uint32_t* col, const uint32_t* val, const void* interned, const void* x,
                     void* out, const uint64_t* chunk_start, const uint64_t* chunk_end, size_t n_chunks,
                     const uint32_t* long_row, const uint32_t* long_first, const uint32_t* long_cnt, size_t n_long,
                     void* chunk_partials);
int launch_mul(cudaStream_t st, const void* a, const void* b, void* out, size_t n);

// K10 gathers
int launch_gather_rows(cudaStream_t st, const void* leaves, size_t w, const uint64_t* idx_dev, size_t n_idx,
                       void* out);
// out[q*depth + d]: d = 0 leaf sibling, then siblings going up (leaf -> root), canonical->canonical
int launch_gather_paths(cudaStream_t st, const void* nodes, size_t L, const uint64_t* idx_dev, size_t n_idx,
                        int depth, void* out);

// PoW: scans nonces [base, base+count) and atomically mins accepted ones into *best (u64 on device)
int launch_pow_scan(cudaStream_t st, fr_arg challenge, fr_arg threshold, uint64_t base, uint64_t count,
                    unsigned long long* best);

// mask generator: n uniform field elements from the ChaCha12 counter stream (key, stream id); see kernels.cu
int launch_rng_fill(cudaStream_t st, void* out, size_t n, const uint32_t key[8], uint32_t stream);

// microbenchmark: chains `iters` dependent Montgomery multiplications per thread; returns launches
int launch_modmul_bench(cudaStream_t st, void* data, size_t n_threads, int iters, bool square);

// ---- glue.cu: protocol flow on the device (transcript, zk-sumcheck bookkeeping, STIR, hints, PoW) ----
int launch_ts_init(cudaStream_t st, void* ts, fr_arg iv_canonical, uint32_t cap_words);
int launch_ts_exchange(cudaStream_t st, void* ts, const void* absorb, int na, bool canonical, void* squeeze, int ns);
int launch_ts_challenge_bytes(cudaStream_t st, void* ts, uint32_t* out_words, int n_bytes);
int launch_ts_hint(cudaStream_t st, void* ts, const uint32_t* payload, uint32_t len_words, const uint32_t* len_ptr);
constexpr size_t DEVTS_HEADER_BYTES = 96;
struct ZkGlue {
    const fr* blind;  // 4 * m0 cubic coefficients of the blinding polynomials (Montgomery)
    fr* suffix;       // m0
    fr* state;        // [0] rho, [1] saved, [2] prefix, [3] sum_g
    fr* alpha;        // m0 challenges
    fr* h3;           // the round kernel's [f(0), f(-1), f(inf)]
    fr* cf;           // the 4 coefficients sent
    int m0;
};
int launch_zk_init(cudaStream_t st, ZkGlue g, fr_arg half);
int launch_zk_glue(cudaStream_t st, ZkGlue g, int idx, fr_arg half, void* ts_or_null);
int launch_powers(cudaStream_t st, void* out, const void* base, int n, int first_exp, const uint32_t* count_ptr);
int launch_expand_powers(cudaStream_t st, void* out, const void* alpha, int m0, size_t len);
constexpr int OPEN_MAX_QUERIES = 128;  // queries per round <= protocol security level (128) at rate <= 1/2
int launch_stir_indices(cudaStream_t st, const uint32_t* bytes_words, int nq, int nb, int folded_log, uint64_t* idx, uint32_t* n_idx);
int launch_open_hints(cudaStream_t st, const void* leaves, bool leaves_canonical, int w, const void* nodes, size_t L, int depth,
                      const uint64_t* idx, const uint32_t* n_idx, uint32_t* hint1, uint32_t* hint2, uint32_t* lens);
int launch_hint_scalars(cudaStream_t st, uint32_t* out, const void* src, int n_groups, int group_len, int estride, int gstride);
struct PowCtrl {
    uint32_t challenge[8];
    unsigned long long ticket, best;
};
int launch_pow_begin(cudaStream_t st, void* ts, PowCtrl* c);
int launch_pow_grind(cudaStream_t st, PowCtrl* c, fr_arg threshold, int blocks);
int launch_pow_end(cudaStream_t st, void* ts, PowCtrl* c);

constexpr int REDUCE_MAX_BLOCKS = 1184;  // 148 SMs x 8
// The multiplier-heavy reduction kernels (sumcheck rounds, batched dot products: 128 registers, two 256-thread blocks per SM)
// launch ONE resident wave and loop: with 1184 blocks a thread of round 0 ran 1.7 iterations before a block reduction of
// three 256-bit sums (5 shuffle levels of 8 SHFL + a field addition each, ~500 instructions) — 13-19 % of the kernel.
#ifndef PK_HEAVY_REDUCE_BLOCKS
#define PK_HEAVY_REDUCE_BLOCKS (148 * 2)
#endif
constexpr int REDUCE_HEAVY_BLOCKS = PK_HEAVY_REDUCE_BLOCKS;
constexpr int REDUCE_MAX_SUMS = 6;       // field sums per reduction kernel
// bytes of the partials work area: per-block partial sums + the ticket counter of the single-launch reduction
constexpr size_t REDUCE_AREA_BYTES = (size_t)REDUCE_MAX_BLOCKS * REDUCE_MAX_SUMS * 32 + 64;

}  // namespace pk
