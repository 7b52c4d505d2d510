// provekit_b200/csrc/kernels.cu — hand-written sm_100a kernels for the WHIR hot path.
// See kernels.cuh for the launcher contracts and DESIGN.md for layouts / rooflines.
#include "kernels.cuh"

#include "fr.cuh"
#include "skyscraper.cuh"

namespace pk {

static __device__ __forceinline__ fr arg_fr(const fr_arg& a) {
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = a.v[i];
    return r;
}

// Shared-memory tile of field elements in SPLIT layout: low and high 16-byte halves in two planes, so that a warp's
// 128-bit accesses to consecutive elements are bank-conflict free (a 32-byte element stride is a 2-way conflict).
struct SmTile {
    uint4* lo;
    uint4* hi;
    __device__ __forceinline__ SmTile(uint4* base, int n_elems) : lo(base), hi(base + n_elems) {}
    __device__ __forceinline__ fr get(int i) const {
        uint4 a = lo[i], b = hi[i];
        fr r;
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void put(int i, const fr& x) const {
        lo[i] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
        hi[i] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    }
};

static inline int grid_for(size_t work, int threads, int max_blocks) {
    size_t b = (work + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (size_t)max_blocks) b = max_blocks;
    return (int)b;
}

// ------------------------------------------------------------------------------------------------
// reductions of field sums
// ------------------------------------------------------------------------------------------------
static __device__ __forceinline__ fr fr_shfl_down(const fr& x, int delta) {
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, x.v[i], delta);
    return r;
}

// every thread contributes acc[NS]; thread 0 of the block writes the block sum to out[NS]
template <int NS>
static __device__ __forceinline__ void block_reduce(fr (&acc)[NS], fr* out) {
    __shared__ fr warp_sums[32 * NS];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc[s] = fr_add(acc[s], fr_shfl_down(acc[s], d));
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    if (lane == 0)
#pragma unroll
        for (int s = 0; s < NS; s++) warp_sums[warp * NS + s] = acc[s];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            fr v = lane < nwarps ? warp_sums[lane * NS + s] : fr_zero();
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v = fr_add(v, fr_shfl_down(v, d));
            if (lane == 0) fr_store(&out[s], v);
        }
    }
}

static __device__ __forceinline__ fr fr_load_cg(const void* p) {  // L2 (cache-global) load: sees other blocks' stores
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldcg(q), b = __ldcg(q + 1);
    fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
// Whole-grid reduction in ONE launch: every block publishes its partial sums, the block that arrives last
// (atomic ticket) adds them up and writes `result`.  Field addition is exact, so the order does not matter.
// The ticket counter lives right behind the partials area and is reset by the last block.
__device__ __forceinline__ unsigned int* reduce_counter(fr* partials) {
    return reinterpret_cast<unsigned int*>(partials + (size_t)REDUCE_MAX_BLOCKS * REDUCE_MAX_SUMS);
}
template <int NS>
static __device__ __forceinline__ void grid_reduce(fr (&acc)[NS], fr* partials, fr* result) {
    __shared__ bool is_last;
    block_reduce<NS>(acc, &partials[(size_t)blockIdx.x * NS]);
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int ticket = atomicAdd(reduce_counter(partials), 1u);
        is_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    fr tot[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) tot[s] = fr_zero();
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
        for (int s = 0; s < NS; s++) tot[s] = fr_add(tot[s], fr_load_cg(&partials[(size_t)b * NS + s]));
    block_reduce<NS>(tot, result);
    if (threadIdx.x == 0) *reduce_counter(partials) = 0;
}

// ------------------------------------------------------------------------------------------------
// Skyscraper compress_many  (seam: CompressManyFn)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_compress_many(const fr* __restrict__ msgs, fr* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr l = sky_reduce(fr_load_nc(&msgs[2 * i]));
    fr r = sky_reduce(fr_load_nc(&msgs[2 * i + 1]));
    fr_store(&out[i], sky_compress(l, r));
}
int launch_compress_many(cudaStream_t st, const void* msgs, void* hashes, size_t n) {
    if (n == 0) return 0;
    k_compress_many<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((const fr*)msgs, (fr*)hashes, n);
    return 1;
}

__global__ void k_convert(fr* data, size_t n, bool to_mont) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr x = fr_load(&data[i]);
    fr_store(&data[i], to_mont ? fr_to_mont(fr_reduce_any(x)) : fr_from_mont(x));
}
int launch_to_mont(cudaStream_t st, void* data, size_t n, bool to_mont) {
    if (n == 0) return 0;
    k_convert<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((fr*)data, n, to_mont);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// K0 wavelet (evaluations <-> coefficients): add/sub butterflies only.  A 2^21 vector is two or three passes over
// HBM: every pass stages a tile of 2^T elements in shared memory and runs the stages of the index bits the tile spans.
// One kernel serves both tile shapes:
//   flat tile : 2^T contiguous elements                                (cbits = T; stages over tile bits [0, T))
//   row tile  : 2^(T-cbits) rows at stride 2^l x 2^cbits contiguous    (stages over tile bits [cbits, T) = index bits [l, ..))
// The stage loop is radix-8 in registers: a thread owns the 8 elements that differ in three consecutive stage bits, runs
// 12 butterflies and writes 7 elements back (element 0 of a group never changes), i.e. 15 shared-memory accesses per
// three stages instead of 36 and one barrier instead of three -- the radix-2 version was shared-memory-bandwidth bound.
// Shared layout: split lo/hi planes (SmTile) with the element index XOR-swizzled by its bits [3,6), which makes the
// stride-8 / stride-64 / ... register-group accesses and the linear load/store phases all bank-conflict free.
// ------------------------------------------------------------------------------------------------
#ifndef PK_WAVELET_TILE_BITS
#define PK_WAVELET_TILE_BITS 11   // 2^11 x 32 B = 64 KB tile
#endif
#ifndef PK_WAVELET_MIN_CBITS
#define PK_WAVELET_MIN_CBITS 1    // shortest contiguous run of a row tile: 2^1 elements = 64 B
#endif
constexpr int WAVELET_TILE_BITS = PK_WAVELET_TILE_BITS;
constexpr int WAVELET_MIN_CBITS = PK_WAVELET_MIN_CBITS;
constexpr int WAVELET_THREADS = 256;

static __device__ __forceinline__ int wv_swz(int e) { return e ^ ((e >> 3) & 7); }

// MODE 0: evaluations <- coefficients (hi += lo); 1: coefficients <- evaluations, the Moebius map M (hi -= lo);
// 2: M transposed (lo -= hi), which maps the monomial weights z^j to the evaluation-basis weights eq(pow(z), x)
template <int MODE, int G>
static __device__ __forceinline__ void wavelet_group(const SmTile& sm, int T, int hb0) {
    constexpr int R = 1 << G;
    const int items = 1 << (T - G);
    for (int t = threadIdx.x; t < items; t += blockDim.x) {
        const int e0 = ((t >> hb0) << (hb0 + G)) | (t & ((1 << hb0) - 1));
        fr x[R];
#pragma unroll
        for (int j = 0; j < R; j++) x[j] = sm.get(wv_swz(e0 + (j << hb0)));
#pragma unroll
        for (int sbit = 0; sbit < G; sbit++)
#pragma unroll
            for (int j = 0; j < R; j++)
                if (!(j & (1 << sbit))) {
                    const int hi = j | (1 << sbit);
                    if (MODE == 2)
                        x[j] = fr_sub(x[j], x[hi]);
                    else
                        x[hi] = MODE == 1 ? fr_sub(x[hi], x[j]) : fr_add(x[hi], x[j]);
                }
        // the element a stage never touches: index 0 of the group (modes 0/1), index R-1 (mode 2)
#pragma unroll
        for (int j = (MODE == 2 ? 0 : 1); j < (MODE == 2 ? R - 1 : R); j++) sm.put(wv_swz(e0 + (j << hb0)), x[j]);
    }
}

template <int MODE>
__global__ void __launch_bounds__(WAVELET_THREADS, 3) k_wavelet_tile(fr* a, int T, int cbits, int l) {
    extern __shared__ uint4 smem_raw[];
    const int n = 1 << T;
    SmTile sm(smem_raw, n);
    const size_t tile = blockIdx.x;
    const size_t lowgrp = tile & (((size_t)1 << (l - cbits)) - 1);
    const size_t hi = tile >> (l - cbits);
    const size_t base = (hi << (l + T - cbits)) | (lowgrp << cbits);
    const int cmask = (1 << cbits) - 1;
    for (int e = threadIdx.x; e < n; e += blockDim.x)
        sm.put(wv_swz(e), fr_load(&a[base + ((size_t)(e >> cbits) << l) + (e & cmask)]));
    __syncthreads();
    // INV (evaluations -> coefficients) and forward both commute across stages: any stage order gives the same result
    for (int hb = (cbits == T ? 0 : cbits); hb < T;) {
        const int g = T - hb >= 3 ? 3 : T - hb;
        if (g == 3)
            wavelet_group<MODE, 3>(sm, T, hb);
        else if (g == 2)
            wavelet_group<MODE, 2>(sm, T, hb);
        else
            wavelet_group<MODE, 1>(sm, T, hb);
        hb += g;
        __syncthreads();
    }
    for (int e = threadIdx.x; e < n; e += blockDim.x)
        fr_store(&a[base + ((size_t)(e >> cbits) << l) + (e & cmask)], sm.get(wv_swz(e)));
}
int launch_wavelet(cudaStream_t st, void* a, int log_n, bool inverse) { return launch_wavelet_mode(st, a, log_n, inverse ? 1 : 0); }
int launch_wavelet_mode(cudaStream_t st, void* a, int log_n, int mode) {
    int launches = 0;
    auto pass = [&](int T, int cbits, int l) {
        size_t smem = (size_t)32 << T;
        unsigned grid = 1u << (log_n - T);
        int threads = (1 << T) / 8 < WAVELET_THREADS ? ((1 << T) / 8 < 32 ? 32 : (1 << T) / 8) : WAVELET_THREADS;
        if (mode == 2)
            k_wavelet_tile<2><<<grid, threads, smem, st>>>((fr*)a, T, cbits, l);
        else if (mode == 1)
            k_wavelet_tile<1><<<grid, threads, smem, st>>>((fr*)a, T, cbits, l);
        else
            k_wavelet_tile<0><<<grid, threads, smem, st>>>((fr*)a, T, cbits, l);
        launches++;
    };
    const int flat = log_n < WAVELET_TILE_BITS ? log_n : WAVELET_TILE_BITS;
    pass(flat, flat, flat);
    // remaining index bits [flat, log_n): as few row passes as the shortest allowed contiguous run permits, balanced
    const int rest = log_n - flat, max_s = WAVELET_TILE_BITS - WAVELET_MIN_CBITS;
    const int npass = rest > 0 ? (rest + max_s - 1) / max_s : 0;
    int l = flat;
    for (int p = 0; p < npass; p++) {
        int S = (log_n - l + (npass - p) - 1) / (npass - p);
        pass(WAVELET_TILE_BITS, WAVELET_TILE_BITS - S, l);
        l += S;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// twiddle table: W[e] = omega^e, omega = 2-adic root of order 2^log_m; pow2[b] = omega^(2^b) from host
// ------------------------------------------------------------------------------------------------
struct TwPow2 {
    uint32_t v[28][8];  // omega^(2^b), Montgomery form; passed by value (kernel parameter space): no device-global state
};
__global__ void k_twiddle_table(fr* table, int log_m, TwPow2 pw) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)1 << (log_m - 1);
    if (e >= n) return;
    fr acc = fr_one();
    for (int b = 0; b < log_m - 1; b++)
        if ((e >> b) & 1) {
            fr w;
#pragma unroll
            for (int k = 0; k < 8; k++) w.v[k] = pw.v[b][k];
            acc = fr_mul(acc, w);
        }
    fr_store(&table[e], acc);
}
int launch_twiddle_table(cudaStream_t st, void* table, int log_m, const uint32_t* host_pow2) {
    if (log_m < 1) return 0;
    TwPow2 pw;
    for (int b = 0; b < 28; b++)
        for (int k = 0; k < 8; k++) pw.v[b][k] = host_pow2[b * 8 + k];
    size_t n = (size_t)1 << (log_m - 1);
    k_twiddle_table<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((fr*)table, log_m, pw);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// K1 Reed-Solomon encode: 16 column NTTs ("prover helps" layout), E coset NTTs of size N' per column,
// multi-pass radix-2 DIF in shared-memory tiles of 2^S points x 16 columns.
// ------------------------------------------------------------------------------------------------
constexpr int NTT_TILE_ELEMS_LOG = 11;  // at most 2^11 elements x 32 B = 64 KB per tile (2^7 points x 16 columns, 2^9 x 4, ...)
constexpr int NTT_MAX_PEERS = 8;
struct NttPass {
    const fr* in;
    fr* out;           // scratch (non-last pass)
    const fr* W;       // twiddle table, order 2^tbl_log
    const fr* Wcan;    // the same table as canonical integers, or null: when set the codeword is emitted canonical
    int tbl_shift;     // omega_M^e = W[e << tbl_shift]
    int L, l, S, logE, logM;
    int first, last;
    int nc_log;        // this launch handles 2^nc_log of the polynomial's 16 columns, starting at col0
    int col0;
    // last pass: row r of the codeword goes to peer[r >> rows_per_peer_log] (this GPU's own block or a peer GPU's
    // block mapped through CUDA IPC) at local row r & mask: the all-to-all of the sharded commit is fused into the
    // final NTT pass as NVLink peer stores
    fr* peer[NTT_MAX_PEERS];
    int rows_per_peer_log;
    size_t leaf_stride, col_offset;
};
__global__ void __launch_bounds__(256) k_ntt_pass(NttPass P) {
    extern __shared__ uint4 smem_raw[];
    const int S = P.S, l = P.l, L = P.L, ncl = P.nc_log;
    const int NC = 1 << ncl;
    SmTile sm(smem_raw, NC << S);
    // twiddles of this tile, staged once: stage hb (half size h = 2^hb) owns tw[h-1 .. 2h-2]
    fr* tw = reinterpret_cast<fr*>(smem_raw + (2 * NC << S));
    const uint32_t tile = blockIdx.x;
    const uint32_t s = tile >> (L - S);
    const uint32_t t = tile & ((1u << (L - S)) - 1u);
    const uint32_t Lo = t & ((1u << l) - 1u);
    const uint32_t H = t >> l;
    const uint32_t qbase = (H << (l + S)) | Lo;
    const int npts = 1 << S;
    const uint32_t halfM = 1u << (P.logM - 1);
    for (int idx = threadIdx.x; idx < npts * NC; idx += blockDim.x) {
        uint32_t mid = idx >> ncl, k = idx & (NC - 1);
        uint32_t q = qbase | (mid << l);
        fr x;
        if (P.first) {
            x = fr_load_nc(&P.in[(size_t)q * 16 + P.col0 + k]);
            if (s || P.Wcan) {  // coset twist omega_M^(s*q); a canonical twiddle also strips the Montgomery factor
                uint32_t e = (s * q) & ((halfM << 1) - 1u);
                bool neg = e >= halfM;
                e &= halfM - 1;
                if (e)
                    x = fr_mul(x, fr_load_nc(&(P.Wcan ? P.Wcan : P.W)[(size_t)e << P.tbl_shift]));
                else if (P.Wcan)
                    x = fr_from_mont(x);
                if (neg) x = fr_neg(x);
            }
        } else {
            x = fr_load(&P.in[((((size_t)s << L) + q) << ncl) + k]);
        }
        sm.put(idx, x);
    }
    for (int t = threadIdx.x; t < npts - 1; t += blockDim.x) {
        int hb = 31 - __clz(t + 1);
        int j = t + 1 - (1 << hb);
        uint32_t e = (((uint32_t)j << l) | Lo) << (L - 1 - hb - l + P.logE);
        fr_store(&tw[t], fr_load_nc(&P.W[(size_t)e << P.tbl_shift]));
    }
    __syncthreads();
    for (int hb = S - 1; hb >= 0; hb--) {
        const int h = 1 << hb;
        const bool unit_twiddle = (Lo == 0) && (hb == 0);  // the only stage whose twiddle is 1 for the whole tile
        for (int bf = threadIdx.x; bf < (npts / 2) * NC; bf += blockDim.x) {
            int k = bf & (NC - 1), b = bf >> ncl;
            int j = b & (h - 1);
            int i0 = ((b >> hb) << (hb + 1)) | j;
            int i1 = i0 + h;
            fr a = sm.get((i0 << ncl) + k), bb = sm.get((i1 << ncl) + k);
            sm.put((i0 << ncl) + k, fr_add(a, bb));
            fr d = fr_sub(a, bb);
            if (!unit_twiddle && (j | Lo)) d = fr_mul(d, fr_load(&tw[h - 1 + j]));
            sm.put((i1 << ncl) + k, d);
        }
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < npts * NC; idx += blockDim.x) {
        uint32_t mid = idx >> ncl, k = idx & (NC - 1);
        uint32_t q = qbase | (mid << l);
        if (P.last) {
            uint32_t tq = L ? (__brev(q) >> (32 - L)) : 0u;
            size_t row = (size_t)s + ((size_t)tq << P.logE);
            fr* base = P.peer[row >> P.rows_per_peer_log];
            size_t lrow = row & (((size_t)1 << P.rows_per_peer_log) - 1);
            fr_store(&base[lrow * P.leaf_stride + P.col_offset + P.col0 + k], sm.get(idx));
        } else {
            fr_store(&P.out[((((size_t)s << L) + q) << ncl) + k], sm.get(idx));
        }
    }
}
// columns [col0, col0 + 2^nc_log) of one polynomial; peers: n_peers leaf blocks of (rows / n_peers) rows each
int launch_rs_encode_cols(cudaStream_t st, const void* coeffs, int log_n, int log_inv_rate, int fold, int col0, int nc_log,
                          void* const* peers, int n_peers, size_t leaf_stride, size_t col_offset, void* scratch,
                          const void* table, int table_log_m, const void* twist_canonical) {
    // fold must be 4 (16 columns): FoldingFactor::Constant(4), provekit/r1cs-compiler/src/whir_r1cs.rs:44
    int L = log_n - fold;
    int logE = log_inv_rate;
    int logM = L + logE;
    int launches = 0;
    int done = 0;  // stage bits processed, from the top
    int tile_bits = NTT_TILE_ELEMS_LOG - nc_log;
    int npass = L == 0 ? 1 : (L + tile_bits - 1) / tile_bits;
    int peers_log = 0;
    while ((1 << peers_log) < n_peers) peers_log++;
    for (int p = 0; p < npass; p++) {
        // balanced split (17 -> 6+6+5 rather than 7+7+3): tiny tiles waste the load/store phases
        int S = (L - done + (npass - p) - 1) / (npass - p);
        NttPass P = {};
        P.first = p == 0;
        P.last = p == npass - 1;
        P.in = (const fr*)(P.first ? coeffs : scratch);
        P.out = (fr*)scratch;
        P.W = (const fr*)table;
        P.Wcan = (const fr*)twist_canonical;
        P.tbl_shift = table_log_m - logM;
        P.L = L;
        P.S = S;
        P.l = L - done - S;
        P.logE = logE;
        P.logM = logM < 1 ? 1 : logM;
        P.nc_log = nc_log;
        P.col0 = col0;
        for (int i = 0; i < n_peers; i++) P.peer[i] = (fr*)peers[i];
        P.rows_per_peer_log = logM - peers_log;
        P.leaf_stride = leaf_stride;
        P.col_offset = col_offset;
        unsigned grid = 1u << (logE + L - S);
        int elems_log = S + nc_log;
        size_t smem = ((size_t)32 << elems_log) + ((size_t)32 << S);
        k_ntt_pass<<<grid, elems_log >= 9 ? 256 : (elems_log >= 8 ? 128 : 32), smem, st>>>(P);
        launches++;
        done += S;
    }
    return launches;
}
int launch_rs_encode(cudaStream_t st, const void* coeffs, int log_n, int log_inv_rate, int fold, void* out,
                     size_t leaf_stride, size_t col_offset, void* scratch, const void* table, int table_log_m,
                     const void* twist_canonical) {
    void* peers[1] = {out};
    return launch_rs_encode_cols(st, coeffs, log_n, log_inv_rate, fold, 0, 4, peers, 1, leaf_stride, col_offset, scratch, table,
                                 table_log_m, twist_canonical);
}

// ------------------------------------------------------------------------------------------------
// K2 Merkle tree: leaf digest = left fold of compress over the leaf (SkyscraperCRH, provekit/common/
// src/skyscraper/whir.rs:30-48), inner nodes = compress(left, right) (:53-74).  Digests are kept
// CANONICAL on device (that is the form compress consumes and the transcript carries).
// ------------------------------------------------------------------------------------------------
// Launch structure: one leaf kernel (pure leaf folds: the ALU-bound bulk, kept free of any block-level tail), then the
// inner levels in at most two launches of k_merkle_reduce, each block taking an aligned sub-tree through 8-10 levels in
// shared memory (every inner node is also stored: pk_commit_open needs the whole tree).  One launch per level, as before,
// cost a launch gap + one compress latency (~11 us) per level with the GPU almost empty; hashing the block's levels inside
// the leaf kernel was measured slower (the divergent tail holds the block's registers: 2.16 vs 1.98 ms on 2^18 x 32).
constexpr int MERKLE_LEAF_BLOCK = 128;
constexpr int MERKLE_TOP_MAX_IN = 2048;  // inputs one block of 1024 threads reduces to the root
// CANON: the leaf elements are already canonical integers (the commit path's RS-encode emits them that way, see
// NttPass::canonical); otherwise they are Montgomery-form field elements (provekit/common/src/skyscraper/whir.rs:21-24
// converts each one with into_bigint)
template <bool CANON>
__global__ void __launch_bounds__(MERKLE_LEAF_BLOCK) k_merkle_leaves(const fr* __restrict__ leaves, size_t L, int w, fr* nodes) {
    const size_t i = (size_t)blockIdx.x * MERKLE_LEAF_BLOCK + threadIdx.x;
    if (i >= L) return;
    const fr* leaf = leaves + i * w;
    fr d = fr_load_nc(&leaf[0]);
    if (!CANON) d = fr_from_mont(d);
    for (int k = 1; k < w; k++) {
        fr x = fr_load_nc(&leaf[k]);
        d = sky_compress(d, CANON ? x : fr_from_mont(x));
    }
    fr_store(&nodes[L + i], d);
}
// n_in nodes of one tree level (heap positions [n_in, 2 n_in)) -> every block reduces `per_block` (= 2 * blockDim.x) of them
// through log2(per_block) levels to one node, storing all inner nodes
__global__ void __launch_bounds__(1024) k_merkle_reduce(fr* nodes, size_t n_in, int per_block, fr* root_mont_out) {
    extern __shared__ uint4 smem_raw[];
    fr* sd = reinterpret_cast<fr*>(smem_raw);
    const int tid = threadIdx.x;
    size_t lvl = n_in, off = (size_t)blockIdx.x * per_block;
    for (int n = per_block >> 1; n >= 1; n >>= 1) {
        fr a, b;
        const bool act = tid < n;
        if (act) {
            if (n == per_block >> 1) {  // first level: children come from global memory
                a = fr_load(&nodes[lvl + off + 2 * tid]);
                b = fr_load(&nodes[lvl + off + 2 * tid + 1]);
            } else {
                a = sd[2 * tid];
                b = sd[2 * tid + 1];
            }
        }
        __syncthreads();
        lvl >>= 1;
        off >>= 1;
        if (act) {
            fr h = sky_compress<0>(a, b);  // latency-bound tree levels: register-only round sums
            sd[tid] = h;
            fr_store(&nodes[lvl + off + tid], h);
            if (root_mont_out && lvl == 1) fr_store(root_mont_out, fr_to_mont(h));
        }
        __syncthreads();
    }
}
int launch_merkle_leaves(cudaStream_t st, const void* leaves, size_t L, size_t w, void* nodes, bool canonical) {
    const unsigned grid = (unsigned)((L + MERKLE_LEAF_BLOCK - 1) / MERKLE_LEAF_BLOCK);
    if (canonical)
        k_merkle_leaves<true><<<grid, MERKLE_LEAF_BLOCK, 0, st>>>((const fr*)leaves, L, (int)w, (fr*)nodes);
    else
        k_merkle_leaves<false><<<grid, MERKLE_LEAF_BLOCK, 0, st>>>((const fr*)leaves, L, (int)w, (fr*)nodes);
    return 1;
}
// all inner levels: L leaf digests at heap positions [L, 2L) -> root at nodes[1]
int launch_merkle_upper(cudaStream_t st, size_t L, void* nodes, void* root_mont_out) {
    int launches = 0;
    size_t n = L;
    while (n > (size_t)MERKLE_TOP_MAX_IN) {
        const int per_block = n > ((size_t)1 << 19) ? 1024 : 256;  // keeps any tree of up to 2^21 leaves at three launches
        k_merkle_reduce<<<(unsigned)(n / per_block), per_block / 2, (per_block / 2) * sizeof(fr), st>>>((fr*)nodes, n, per_block, nullptr);
        n /= per_block;
        launches++;
    }
    if (n >= 2) {
        const int threads = (int)(n / 2) < 32 ? 32 : (int)(n / 2);
        k_merkle_reduce<<<1, threads, (n / 2) * sizeof(fr), st>>>((fr*)nodes, n, (int)n, (fr*)root_mont_out);
        launches++;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------------
// tensor-product tables (eq / power tables), accumulate, dot products
// ------------------------------------------------------------------------------------------------
// block (2k + part): part 0 builds the high table of point k (variables [0, nv_hi), scaled), part 1 the low table
// (variables [nv_hi, nv_hi + nv_lo))
__global__ void __launch_bounds__(256) k_tensor_tables(const fr* __restrict__ points, int pt_stride, int nv_hi, int nv_lo,
                                                       const fr* __restrict__ scales, bool eq_mode, fr* out_hi, fr* out_lo,
                                                       TensorSrc src) {
    const size_t k = blockIdx.x >> 1;
    const int part = blockIdx.x & 1;
    const int nv = part ? nv_lo : nv_hi, var_off = part ? nv_hi : 0;
    fr* T = (part ? out_lo : out_hi) + (k << nv);
    const bool live = !src.count || k < (size_t)*src.count;
    if (threadIdx.x == 0) fr_store(&T[0], !live ? fr_zero() : ((!part && scales) ? fr_load(&scales[k]) : fr_one()));
    // univariate points: variable j of n = nv_hi + nv_lo is z^(2^(n-1-j)); the loop below walks j downwards, i.e. the
    // power doubles every step, starting from z^(2^(n - var_off - nv)) (every thread carries the same running power)
    fr zp = fr_zero();
    if (src.mode == TENSOR_PTS_UNIVARIATE) {
        zp = fr_load(&points[k]);
    } else if (src.mode == TENSOR_PTS_ROOTS) {
        const uint32_t D = 1u << src.log_d, halfD = D >> 1;
        uint32_t e = live ? (uint32_t)src.exps[k] & (D - 1) : 0u;
        const bool neg = halfD && e >= halfD;
        if (halfD) e &= halfD - 1;
        zp = e ? fr_load_nc(&reinterpret_cast<const fr*>(src.W)[(size_t)e << src.tbl_shift]) : fr_one();
        if (neg) zp = fr_neg(zp);
    }
    if (src.mode != TENSOR_PTS_EXPLICIT)
        for (int i = nv_hi + nv_lo - var_off - nv; i > 0; i--) zp = fr_sqr(zp);
    __syncthreads();
    int len = 1;
    for (int j = nv - 1; j >= 0; j--) {  // last variable <-> least significant index bit
        fr x;
        if (src.mode == TENSOR_PTS_EXPLICIT) {
            x = fr_load(&points[k * pt_stride + var_off + j]);
        } else {
            x = zp;
            zp = fr_sqr(zp);
        }
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            fr t = fr_load(&T[i]);
            fr s1 = fr_mul(t, x);
            fr_store(&T[i + len], s1);
            if (eq_mode) fr_store(&T[i], fr_sub(t, s1));
        }
        len <<= 1;
        __syncthreads();
    }
}
int launch_tensor_tables(cudaStream_t st, const void* points, size_t K, int pt_stride, int nv_hi, int nv_lo,
                         const void* scales, bool eq_mode, void* out_hi, void* out_lo, TensorSrc src) {
    if (K == 0) return 0;
    k_tensor_tables<<<(unsigned)(2 * K), 256, 0, st>>>((const fr*)points, pt_stride, nv_hi, nv_lo, (const fr*)scales, eq_mode,
                                                     (fr*)out_hi, (fr*)out_lo, src);
    return 1;
}
__global__ void __launch_bounds__(256) k_tensor_accumulate(fr* out, size_t n, const fr* __restrict__ hi,
                                                           const fr* __restrict__ lo, int K, int lo_bits, int hi_bits) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    size_t ih = idx >> lo_bits, il = idx & (((size_t)1 << lo_bits) - 1);
    fr acc = fr_load(&out[idx]);
    for (int k = 0; k < K; k++) {
        fr h = fr_load_nc(&hi[((size_t)k << hi_bits) + ih]);
        fr l = fr_load_nc(&lo[((size_t)k << lo_bits) + il]);
        acc = fr_add(acc, fr_mul(h, l));
    }
    fr_store(&out[idx], acc);
}
int launch_tensor_accumulate(cudaStream_t st, void* out, int log_n, const void* hi, const void* lo, size_t K,
                             int lo_bits) {
    size_t n = (size_t)1 << log_n;
    k_tensor_accumulate<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((fr*)out, n, (const fr*)hi, (const fr*)lo, (int)K,
                                                                     lo_bits, log_n - lo_bits);
    return 1;
}
__global__ void __launch_bounds__(256) k_tensor_dot(const fr* __restrict__ a, size_t n, const fr* __restrict__ hi,
                                                    const fr* __restrict__ lo, int lo_bits, fr* partials, fr* result) {
    fr acc[1] = {fr_zero()};
    size_t mask = ((size_t)1 << lo_bits) - 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        fr t = fr_mul(fr_load_nc(&hi[i >> lo_bits]), fr_load_nc(&lo[i & mask]));
        acc[0] = fr_add(acc[0], fr_mul(fr_load_nc(&a[i]), t));
    }
    grid_reduce<1>(acc, partials, result);
}
int launch_tensor_dot(cudaStream_t st, const void* a, size_t n, const void* hi, const void* lo, int lo_bits,
                      void* partials, void* result) {
    int g = grid_for(n, 256, REDUCE_MAX_BLOCKS);
    k_tensor_dot<<<g, 256, 0, st>>>((const fr*)a, n, (const fr*)hi, (const fr*)lo, lo_bits, (fr*)partials, (fr*)result);
    return 1;
}
__global__ void __launch_bounds__(256) k_dot(const fr* __restrict__ a, const fr* __restrict__ b, size_t n, fr* partials,
                                             fr* result) {
    fr acc[1] = {fr_zero()};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc[0] = fr_add(acc[0], fr_mul(fr_load_nc(&a[i]), fr_load_nc(&b[i])));
    grid_reduce<1>(acc, partials, result);
}
int launch_dot(cudaStream_t st, const void* a, const void* b, size_t n, void* partials, void* result) {
    int g = grid_for(n, 256, REDUCE_MAX_BLOCKS);
    k_dot<<<g, 256, 0, st>>>((const fr*)a, (const fr*)b, n, (fr*)partials, (fr*)result);
    return 1;
}
// result[ja * NB + jb] = <a_ja, b_jb>: one pass over NA + NB arrays instead of NA * NB dot products
// (the six sums <w_j, f>, <w_j, g> of create_combined_statement_over_two_polynomials, whir_r1cs.rs:382-412)
struct DotPtrs {
    const fr* a[3];
    const fr* b[2];
};
template <int NA, int NB>
__global__ void __launch_bounds__(256) k_multi_dot(DotPtrs P, size_t n, fr* partials, fr* result) {
    fr acc[NA * NB];
#pragma unroll
    for (int s = 0; s < NA * NB; s++) acc[s] = fr_zero();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        fr bv[NB];
#pragma unroll
        for (int jb = 0; jb < NB; jb++) bv[jb] = fr_load_nc(&P.b[jb][i]);
#pragma unroll
        for (int ja = 0; ja < NA; ja++) {
            fr av = fr_load_nc(&P.a[ja][i]);
#pragma unroll
            for (int jb = 0; jb < NB; jb++) acc[ja * NB + jb] = fr_add(acc[ja * NB + jb], fr_mul(av, bv[jb]));
        }
    }
    grid_reduce<NA * NB>(acc, partials, result);
}
int launch_multi_dot(cudaStream_t st, const void* const* a, int na, const void* const* b, int nb, size_t n, void* partials,
                     void* result) {
    DotPtrs P = {};
    for (int i = 0; i < na; i++) P.a[i] = (const fr*)a[i];
    for (int i = 0; i < nb; i++) P.b[i] = (const fr*)b[i];
    int g = grid_for(n, 256, REDUCE_HEAVY_BLOCKS);
    if (na == 3 && nb == 2)
        k_multi_dot<3, 2><<<g, 256, 0, st>>>(P, n, (fr*)partials, (fr*)result);
    else if (na == 1 && nb == 2)
        k_multi_dot<1, 2><<<g, 256, 0, st>>>(P, n, (fr*)partials, (fr*)result);
    else
        return -1;
    return 1;
}
// results[j] = sum_idx a_j[idx] * hi[idx >> lo_bits] * lo[idx & mask], j < NA (same tensor point for all arrays)
template <int NA>
__global__ void __launch_bounds__(256) k_multi_tensor_dot(DotPtrs P, size_t n, const fr* __restrict__ hi,
                                                          const fr* __restrict__ lo, int lo_bits, fr* partials, fr* result) {
    fr acc[NA];
#pragma unroll
    for (int s = 0; s < NA; s++) acc[s] = fr_zero();
    size_t mask = ((size_t)1 << lo_bits) - 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        fr t = fr_mul(fr_load_nc(&hi[i >> lo_bits]), fr_load_nc(&lo[i & mask]));
#pragma unroll
        for (int ja = 0; ja < NA; ja++) acc[ja] = fr_add(acc[ja], fr_mul(fr_load_nc(&P.a[ja][i]), t));
    }
    grid_reduce<NA>(acc, partials, result);
}
int launch_multi_tensor_dot(cudaStream_t st, const void* const* a, int na, size_t n, const void* hi, const void* lo,
                            int lo_bits, void* partials, void* result) {
    DotPtrs P = {};
    for (int i = 0; i < na; i++) P.a[i] = (const fr*)a[i];
    int g = grid_for(n, 256, REDUCE_HEAVY_BLOCKS);
    if (na == 1)
        k_multi_tensor_dot<1><<<g, 256, 0, st>>>(P, n, (const fr*)hi, (const fr*)lo, lo_bits, (fr*)partials, (fr*)result);
    else if (na == 2)
        k_multi_tensor_dot<2><<<g, 256, 0, st>>>(P, n, (const fr*)hi, (const fr*)lo, lo_bits, (fr*)partials, (fr*)result);
    else if (na == 3)
        k_multi_tensor_dot<3><<<g, 256, 0, st>>>(P, n, (const fr*)hi, (const fr*)lo, lo_bits, (fr*)partials, (fr*)result);
    else
        return -1;
    return 1;
}
__global__ void __launch_bounds__(256) k_axpy(fr* y, const fr* __restrict__ x, fr_arg a, const fr* a_ptr, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr s = a_ptr ? fr_load(a_ptr) : arg_fr(a);
    fr_store(&y[i], fr_add(fr_load(&y[i]), fr_mul(s, fr_load_nc(&x[i]))));
}
int launch_axpy(cudaStream_t st, void* y, const void* x, fr_arg a, const void* a_ptr, size_t n) {
    if (n == 0) return 0;
    k_axpy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((fr*)y, (const fr*)x, a, (const fr*)a_ptr, n);
    return 1;
}
// K8: 2^k consecutive coefficients -> multilinear value, r[j] binds bit j (k <= 4)
__global__ void __launch_bounds__(128) k_fold_coeffs(const fr* __restrict__ c, size_t nout, const fr* __restrict__ r,
                                                     int r_stride, int k, fr* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout) return;
    int w = 1 << k;
    fr tmp[16];
#pragma unroll
    for (int j = 0; j < 16; j++)
        if (j < w) tmp[j] = fr_load_nc(&c[i * w + j]);
    int len = w;
#pragma unroll
    for (int v = 0; v < 4; v++) {
        if (v < k) {
            fr rv = fr_load(&r[v * r_stride]);
            len >>= 1;
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < len) tmp[j] = fr_add(tmp[2 * j], fr_mul(rv, tmp[2 * j + 1]));
        }
    }
    fr_store(&out[i], tmp[0]);
}
int launch_fold_coeffs(cudaStream_t st, const void* coeffs, int log_n, const void* r_dev, int r_stride, int k, void* out) {
    size_t nout = (size_t)1 << (log_n - k);
    k_fold_coeffs<<<(unsigned)((nout + 127) / 128), 128, 0, st>>>((const fr*)coeffs, nout, (const fr*)r_dev, r_stride, k, (fr*)out);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// eq weights of many UNIVARIATE points that are roots of unity (the STIR constraints of a WHIR round):
//   sum_k s_k eq(pow(z_k), x),  z_k = omega_D^(e_k),   pow(z) = (z^(2^(n-1)), .., z^2, z).
// In the monomial basis the weight of pow(z) is u_j = z^j, and eq(pow(z), .) = M^T u with M the evaluations->coefficients
// map, so the whole batch is  M^T ( DFT_D(sparse s)[0 .. 2^n) ):  a sparse scatter, one NTT of size D (the RS-encode
// kernels at rate 1 plus the radix-16 recombination below) and a transposed wavelet pass -- O(D log D) multiplications
// instead of 2^n per point.  Field arithmetic is exact, so the result equals the per-point tensor method bit for bit.
// ------------------------------------------------------------------------------------------------
// s[e_k] += scalar_k, serially (k is ~100; duplicates of e_k must add up)
__global__ void k_scatter_add(fr* s, const uint64_t* exps, const fr* scalars, size_t k, const uint32_t* k_ptr) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (k_ptr) k = *k_ptr;
        for (size_t i = 0; i < k; i++) fr_store(&s[exps[i]], fr_add(fr_load(&s[exps[i]]), fr_load(&scalars[i])));
    }
}
// u[j] = sum_{c<16} omega_D^(j c) * lv[(j mod D/16) * 16 + c],  j < n_out  (lv: RS-encode leaf layout, leaf i entry c =
// f_c((omega_D^16)^i) with f_c the stride-16 sub-polynomials): the last radix-16 step of the size-D transform
__global__ void __launch_bounds__(128) k_dft16_combine(const fr* __restrict__ lv, fr* u, size_t n_out, int log_d,
                                                       const fr* __restrict__ W, int tbl_shift) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const uint32_t D = 1u << log_d, halfD = D >> 1;
    const fr* row = lv + ((j & ((D >> 4) - 1)) << 4);
    fr acc = fr_load_nc(&row[0]);
#pragma unroll 1
    for (uint32_t c = 1; c < 16; c++) {
        uint32_t e = ((uint32_t)j * c) & (D - 1);
        fr t = fr_load_nc(&row[c]);
        const bool neg = e >= halfD;
        e &= halfD - 1;
        if (e) t = fr_mul(t, fr_load_nc(&W[(size_t)e << tbl_shift]));
        acc = neg ? fr_sub(acc, t) : fr_add(acc, t);
    }
    fr_store(&u[j], acc);
}
__global__ void __launch_bounds__(256) k_add_inplace(fr* y, const fr* __restrict__ x, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_store(&y[i], fr_add(fr_load(&y[i]), fr_load_nc(&x[i])));
}
int launch_scatter_add(cudaStream_t st, void* s, const uint64_t* exps, const void* scalars, size_t k, const uint32_t* k_ptr) {
    if (k == 0 && !k_ptr) return 0;
    k_scatter_add<<<1, 32, 0, st>>>((fr*)s, exps, (const fr*)scalars, k, k_ptr);
    return 1;
}
int launch_dft16_combine(cudaStream_t st, const void* lv, void* u, size_t n_out, int log_d, const void* table, int table_log_m) {
    k_dft16_combine<<<(unsigned)((n_out + 127) / 128), 128, 0, st>>>((const fr*)lv, (fr*)u, n_out, log_d, (const fr*)table,
                                                                   table_log_m - log_d);
    return 1;
}
int launch_add_inplace(cudaStream_t st, void* y, const void* x, size_t n) {
    if (n == 0) return 0;
    k_add_inplace<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((fr*)y, (const fr*)x, n);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// K5 zk-sumcheck round: fused fold + map + reduce (provekit/common/src/utils/sumcheck.rs:16-104 with
// the map of provekit/prover/src/whir_r1cs.rs:284-291).  MSB pairing: i <-> i + len/2.
// ------------------------------------------------------------------------------------------------
template <bool FOLD>
__global__ void __launch_bounds__(256, 2) k_zk_sumcheck(fr* a, fr* b, fr* c, fr* eq, size_t half, fr_arg foldv,
                                                     const fr* fold_ptr, fr* partials, fr* result) {
    // `half` = (length after folding) / 2; before folding the arrays are 4*half long
    fr acc[3] = {fr_zero(), fr_zero(), fr_zero()};
    fr f = fold_ptr ? fr_load(fold_ptr) : arg_fr(foldv);
    // one array at a time, (x0, d = x1 - x0) per array, products formed as soon as their inputs exist: keeps the live set
    // near 90 registers so two 256-thread blocks fit per SM
    auto fetch = [&](fr* x, size_t i, fr& x0, fr& d) {
        if (FOLD) {
            fr q0 = fr_load(&x[i]), q2 = fr_load(&x[i + 2 * half]);
            x0 = fr_add(q0, fr_mul(f, fr_sub(q2, q0)));
            fr_store(&x[i], x0);
            fr q1 = fr_load(&x[i + half]), q3 = fr_load(&x[i + 3 * half]);
            fr x1 = fr_add(q1, fr_mul(f, fr_sub(q3, q1)));
            fr_store(&x[i + half], x1);
            d = fr_sub(x1, x0);
        } else {
            x0 = fr_load(&x[i]);
            d = fr_sub(fr_load(&x[i + half]), x0);
        }
    };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
        fr x0, d, y0, e;
        fetch(a, i, x0, d);
        fetch(b, i, y0, e);
        fr t1 = fr_mul(x0, y0);                          // a0 b0
        fr t2 = fr_mul(fr_sub(x0, d), fr_sub(y0, e));    // (2a0 - a1)(2b0 - b1)
        fr t3 = fr_mul(d, e);                            // (a1 - a0)(b1 - b0)
        fetch(c, i, x0, d);
        t1 = fr_sub(t1, x0);                             // a0 b0 - c0
        t2 = fr_sub(t2, fr_sub(x0, d));                  // ... - (2c0 - c1)
        fetch(eq, i, x0, d);
        acc[0] = fr_add(acc[0], fr_mul(x0, t1));                 // f(0)
        acc[1] = fr_add(acc[1], fr_mul(fr_sub(x0, d), t2));      // f(-1)
        acc[2] = fr_add(acc[2], fr_mul(d, t3));                  // f(inf)
    }
    grid_reduce<3>(acc, partials, result);
}
int launch_zk_sumcheck_round(cudaStream_t st, void* a, void* b, void* c, void* eq, int log_n, bool has_fold,
                             fr_arg fold, const void* fold_ptr, void* partials, void* result) {
    size_t n_after = (size_t)1 << (has_fold ? log_n - 1 : log_n);
    size_t half = n_after / 2;
    int g = grid_for(half, 256, REDUCE_HEAVY_BLOCKS);
    if (has_fold)
        k_zk_sumcheck<true><<<g, 256, 0, st>>>((fr*)a, (fr*)b, (fr*)c, (fr*)eq, half, fold, (const fr*)fold_ptr, (fr*)partials,
                                               (fr*)result);
    else
        k_zk_sumcheck<false><<<g, 256, 0, st>>>((fr*)a, (fr*)b, (fr*)c, (fr*)eq, half, fold, nullptr, (fr*)partials, (fr*)result);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// K7 WHIR sumcheck round [whir SumcheckSingle]: LSB pairing (2i, 2i+1), sends h(0), h(1), h(2).
// ------------------------------------------------------------------------------------------------
template <bool FOLD>
__global__ void __launch_bounds__(256, 2) k_whir_sumcheck(const fr* __restrict__ p_in, const fr* __restrict__ w_in,
                                                       fr* p_out, fr* w_out, size_t pairs, fr_arg foldv, const fr* fold_ptr,
                                                       fr* partials, fr* result) {
    fr acc[3] = {fr_zero(), fr_zero(), fr_zero()};
    fr f = fold_ptr ? fr_load(fold_ptr) : arg_fr(foldv);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (size_t)gridDim.x * blockDim.x) {
        fr p0, p1, w0, w1;
        if (FOLD) {
            fr x0 = fr_load_nc(&p_in[4 * i]), x1 = fr_load_nc(&p_in[4 * i + 1]);
            fr x2 = fr_load_nc(&p_in[4 * i + 2]), x3 = fr_load_nc(&p_in[4 * i + 3]);
            p0 = fr_add(x0, fr_mul(f, fr_sub(x1, x0)));
            p1 = fr_add(x2, fr_mul(f, fr_sub(x3, x2)));
            fr_store(&p_out[2 * i], p0);
            fr_store(&p_out[2 * i + 1], p1);
            x0 = fr_load_nc(&w_in[4 * i]); x1 = fr_load_nc(&w_in[4 * i + 1]);
            x2 = fr_load_nc(&w_in[4 * i + 2]); x3 = fr_load_nc(&w_in[4 * i + 3]);
            w0 = fr_add(x0, fr_mul(f, fr_sub(x1, x0)));
            w1 = fr_add(x2, fr_mul(f, fr_sub(x3, x2)));
            fr_store(&w_out[2 * i], w0);
            fr_store(&w_out[2 * i + 1], w1);
        } else {
            p0 = fr_load_nc(&p_in[2 * i]); p1 = fr_load_nc(&p_in[2 * i + 1]);
            w0 = fr_load_nc(&w_in[2 * i]); w1 = fr_load_nc(&w_in[2 * i + 1]);
        }
        acc[0] = fr_add(acc[0], fr_mul(p0, w0));
        acc[1] = fr_add(acc[1], fr_mul(p1, w1));
        acc[2] = fr_add(acc[2], fr_mul(fr_sub(fr_dbl(p1), p0), fr_sub(fr_dbl(w1), w0)));
    }
    grid_reduce<3>(acc, partials, result);
}
int launch_whir_sumcheck_round(cudaStream_t st, const void* p_in, const void* w_in, void* p_out, void* w_out,
                               int log_n, bool has_fold, fr_arg fold, const void* fold_ptr, void* partials, void* result) {
    size_t n_after = (size_t)1 << (has_fold ? log_n - 1 : log_n);
    size_t pairs = n_after / 2;
    int g = grid_for(pairs, 256, REDUCE_HEAVY_BLOCKS);
    if (has_fold)
        k_whir_sumcheck<true><<<g, 256, 0, st>>>((const fr*)p_in, (const fr*)w_in, (fr*)p_out, (fr*)w_out, pairs, fold,
                                                 (const fr*)fold_ptr, (fr*)partials, (fr*)result);
    else
        k_whir_sumcheck<false><<<g, 256, 0, st>>>((const fr*)p_in, (const fr*)w_in, (fr*)p_out, (fr*)w_out, pairs, fold, nullptr,
                                                  (fr*)partials, (fr*)result);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// Sharded sumchecks: exchange of the round's partial sums over peer memory (see kernels.cuh)
// ------------------------------------------------------------------------------------------------
static __device__ __forceinline__ void st_sys(void* p, const fr& x) {  // peer-visible 128-bit stores
    volatile uint4* q = reinterpret_cast<volatile uint4*>(p);
    q[0].x = x.v[0]; q[0].y = x.v[1]; q[0].z = x.v[2]; q[0].w = x.v[3];
    q[1].x = x.v[4]; q[1].y = x.v[5]; q[1].z = x.v[6]; q[1].w = x.v[7];
}
static __device__ __forceinline__ fr ld_sys(const void* p) {
    const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(p);
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = q[i];
    return r;
}
// GATHER = false: result[0..3) <- field sum over the ranks of their result[0..3) (the sumcheck round message).
// GATHER = true : gathered[3 r .. 3 r + 3) <- rank r's result[0..3), nothing summed: an all-gather of 96 bytes per rank, and —
// whatever the payload — a barrier ON THE STREAM: kernels enqueued behind it start only after every rank of the group has
// reached the same point of ITS stream (the sharded commitment: a rank's rows are complete once every peer's last NTT pass,
// which stores into them over NVLink, is done).
template <bool GATHER>
__global__ void __launch_bounds__(32) k_shard_exchange(fr* result, ShardGroup G, uint32_t seq, uint32_t* status, fr* gathered) {
    const int lane = threadIdx.x;
    const size_t slot = seq % SHARD_SLOTS;
    bool ok = true;
    fr v[3] = {fr_zero(), fr_zero(), fr_zero()};
    if (lane < G.world) {
        // publish my partial sums into cell [slot][my rank] of rank `lane` (my own mailbox included)
        uint8_t* dst = G.mbox[lane] + (slot * SHARD_MAX_WORLD + G.rank) * SHARD_CELL_BYTES;
#pragma unroll
        for (int s = 0; s < 3; s++) st_sys(dst + 32 * s, fr_load(&result[s]));
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(dst + 96) = seq;
        // collect rank `lane`'s partial sums from my own mailbox
        const uint8_t* src = G.mbox[G.rank] + (slot * SHARD_MAX_WORLD + lane) * SHARD_CELL_BYTES;
        const volatile uint32_t* flag = reinterpret_cast<const volatile uint32_t*>(src + 96);
        uint32_t spins = 0;
        while (*flag != seq) {
            if (++spins > (1u << 27)) {  // ~10 s: a peer never arrived; report instead of hanging the GPU
                ok = false;
                break;
            }
            __nanosleep(40);
        }
        __threadfence_system();
        if (ok)
#pragma unroll
            for (int s = 0; s < 3; s++) v[s] = ld_sys(src + 32 * s);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (GATHER) {
        if (lane < G.world)
#pragma unroll
            for (int s = 0; s < 3; s++) fr_store(&gathered[3 * lane + s], v[s]);
    } else {
#pragma unroll
        for (int s = 0; s < 3; s++) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v[s] = fr_add(v[s], fr_shfl_down(v[s], d));
            if (lane == 0) fr_store(&result[s], v[s]);
        }
    }
    // sticky: a time-out on an earlier exchange of this stream must not be overwritten by a later success
    if (lane == 0 && !ok) *status = 1u;
}
int launch_shard_exchange(cudaStream_t st, void* result, ShardGroup g, uint32_t seq, uint32_t* status) {
    k_shard_exchange<false><<<1, 32, 0, st>>>((fr*)result, g, seq, status, nullptr);
    return 1;
}
int launch_shard_gather(cudaStream_t st, void* payload3, ShardGroup g, uint32_t seq, uint32_t* status, void* gathered) {
    k_shard_exchange<true><<<1, 32, 0, st>>>((fr*)payload3, g, seq, status, (fr*)gathered);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// R1CS sparse mat-vec (one thread per row; rows hold 1-3 entries in practice)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_spmv(const uint64_t* __restrict__ row_start, const uint32_t* __restrict__ col,
                                              const uint32_t* __restrict__ val, const fr* __restrict__ interned,
                                              const fr* __restrict__ x, fr* out, size_t num_rows, size_t nnz) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= num_rows) return;
    size_t s = row_start[r], e = r + 1 < num_rows ? row_start[r + 1] : nnz;
    if (e - s > (size_t)SPMV_LONG_ROW) return;  // handled by k_spmv_long_*
    fr acc = fr_zero();
    for (size_t k = s; k < e; k++) acc = fr_add(acc, fr_mul(fr_load_nc(&interned[val[k]]), fr_load_nc(&x[col[k]])));
    fr_store(&out[r], acc);
}
__global__ void __launch_bounds__(256) k_spmv_long_chunks(const uint32_t* __restrict__ col, const uint32_t* __restrict__ val,
                                                          const fr* __restrict__ interned, const fr* __restrict__ x,
                                                          const uint64_t* __restrict__ chunk_start,
                                                          const uint64_t* __restrict__ chunk_end, fr* partials) {
    size_t s = chunk_start[blockIdx.x], e = chunk_end[blockIdx.x];
    fr acc[1] = {fr_zero()};
    for (size_t k = s + threadIdx.x; k < e; k += blockDim.x)
        acc[0] = fr_add(acc[0], fr_mul(fr_load_nc(&interned[val[k]]), fr_load_nc(&x[col[k]])));
    block_reduce<1>(acc, &partials[blockIdx.x]);
}
__global__ void __launch_bounds__(32) k_spmv_long_rows(const fr* __restrict__ partials, const uint32_t* __restrict__ long_row,
                                                       const uint32_t* __restrict__ long_first,
                                                       const uint32_t* __restrict__ long_cnt, fr* out) {
    // one warp per long row
    uint32_t first = long_first[blockIdx.x], cnt = long_cnt[blockIdx.x];
    fr acc[1] = {fr_zero()};
    for (uint32_t c = threadIdx.x; c < cnt; c += 32) acc[0] = fr_add(acc[0], fr_load(&partials[first + c]));
    block_reduce<1>(acc, &out[long_row[blockIdx.x]]);
}
int launch_spmv_long(cudaStream_t st, const uint32_t* col, const uint32_t* val, const void* interned, const void* x,
                     void* out, const uint64_t* chunk_start, const uint64_t* chunk_end, size_t n_chunks,
                     const uint32_t* long_row, const uint32_t* long_first, const uint32_t* long_cnt, size_t n_long,
                     void* chunk_partials) {
    if (n_long == 0) return 0;
    k_spmv_long_chunks<<<(unsigned)n_chunks, 256, 0, st>>>(col, val, (const fr*)interned, (const fr*)x, chunk_start, chunk_end,
                                                         (fr*)chunk_partials);
    k_spmv_long_rows<<<(unsigned)n_long, 32, 0, st>>>((const fr*)chunk_partials, long_row, long_first, long_cnt, (fr*)out);
    return 2;
}
int launch_spmv(cudaStream_t st, const uint64_t* row_start, const uint32_t* col, const uint32_t* val,
                const void* interned, const void* x, void* out, size_t num_rows, size_t nnz) {
    if (num_rows == 0) return 0;
    k_spmv<<<(unsigned)((num_rows + 255) / 256), 256, 0, st>>>(row_start, col, val, (const fr*)interned, (const fr*)x,
                                                             (fr*)out, num_rows, nnz);
    return 1;
}
__global__ void __launch_bounds__(256) k_mul(const fr* __restrict__ a, const fr* __restrict__ b, fr* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_store(&out[i], fr_mul(fr_load_nc(&a[i]), fr_load_nc(&b[i])));
}
int launch_mul(cudaStream_t st, const void* a, const void* b, void* out, size_t n) {
    if (n == 0) return 0;
    k_mul<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const fr*)a, (const fr*)b, (fr*)out, n);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// K10 gathers for STIR answers / Merkle multipaths
// ------------------------------------------------------------------------------------------------
__global__ void k_gather_rows(const fr* __restrict__ leaves, size_t w, const uint64_t* __restrict__ idx, size_t n_idx,
                              fr* out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * w) return;
    size_t q = t / w, k = t % w;
    fr_store(&out[t], fr_load_nc(&leaves[idx[q] * w + k]));
}
int launch_gather_rows(cudaStream_t st, const void* leaves, size_t w, const uint64_t* idx_dev, size_t n_idx, void* out) {
    size_t n = n_idx * w;
    if (n == 0) return 0;
    k_gather_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const fr*)leaves, w, idx_dev, n_idx, (fr*)out);
    return 1;
}
__global__ void k_gather_paths(const fr* __restrict__ nodes, size_t L, const uint64_t* __restrict__ idx, size_t n_idx,
                               int depth, fr* out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * depth) return;
    size_t q = t / depth;
    int d = (int)(t % depth);
    size_t pos = (L + idx[q]) >> d;
    fr_store(&out[t], fr_load_nc(&nodes[pos ^ 1]));
}
int launch_gather_paths(cudaStream_t st, const void* nodes, size_t L, const uint64_t* idx_dev, size_t n_idx, int depth,
                        void* out) {
    size_t n = n_idx * depth;
    if (n == 0) return 0;
    k_gather_paths<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const fr*)nodes, L, idx_dev, n_idx, depth, (fr*)out);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// PoW grinding (skyscraper/core/src/pow.rs:24-41, generic.rs:42-71): accept iff
// compress(challenge, [nonce,0,0,0]) < threshold; the smallest accepted nonce wins (fetch_min).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pow_scan(fr_arg challenge, fr_arg threshold, uint64_t base, uint64_t count,
                                                  unsigned long long* best) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    uint64_t nonce = base + t;
    // a smaller accepted nonce already exists: nothing in this (higher) range can win (fetch_min semantics,
    // skyscraper/core/src/generic.rs:56-58)
    if (*reinterpret_cast<volatile unsigned long long*>(best) < nonce) return;
    fr l = sky_reduce(arg_fr(challenge));
    fr r = fr_zero();
    r.v[0] = (uint32_t)nonce;
    r.v[1] = (uint32_t)(nonce >> 32);
    // re-check between round pairs: the nonces in flight when the winner lands stop after at most two more rounds
    bool done;
    fr h = sky_compress_while(
        l, r, [&](int j) { return j == 0 || *reinterpret_cast<volatile unsigned long long*>(best) >= nonce; }, done);
    if (!done) return;
    fr thr = arg_fr(threshold);
    bool less = false;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (h.v[i] != thr.v[i]) {
            less = h.v[i] < thr.v[i];
            break;
        }
    }
    if (less) atomicMin(best, (unsigned long long)nonce);
}
int launch_pow_scan(cudaStream_t st, fr_arg challenge, fr_arg threshold, uint64_t base, uint64_t count,
                    unsigned long long* best) {
    k_pow_scan<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(challenge, threshold, base, count, best);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// microbenchmark: modmul/s ceiling of the integer pipe (two independent chains per thread)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_modmul_bench(fr* data, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    fr x = fr_load(&data[2 * i]), y = fr_load(&data[2 * i + 1]);
    for (int it = 0; it < iters; it++) {
        x = fr_mul(x, y);
        y = fr_mul(y, x);
    }
    fr_store(&data[2 * i], x);
    fr_store(&data[2 * i + 1], y);
}
__global__ void __launch_bounds__(256) k_modsqr_bench(fr* data, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    fr x = fr_load(&data[2 * i]), y = fr_load(&data[2 * i + 1]);
    for (int it = 0; it < iters; it++) {
        x = fr_sqr(x);
        y = fr_sqr(y);
    }
    fr_store(&data[2 * i], x);
    fr_store(&data[2 * i + 1], y);
}
int launch_modmul_bench(cudaStream_t st, void* data, size_t n_threads, int iters, bool square) {
    if (square)
        k_modsqr_bench<<<(unsigned)(n_threads / 256), 256, 0, st>>>((fr*)data, iters);
    else
        k_modmul_bench<<<(unsigned)(n_threads / 256), 256, 0, st>>>((fr*)data, iters);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// mask generator: the device twin of `F::rand(&mut thread_rng())` (provekit/common/src/utils/zk_utils.rs:13-22,
// provekit/prover/src/whir_r1cs.rs:211-225).  thread_rng is a ChaCha12 stream and Fp::rand rejection-samples
// 254-bit strings below p; here the stream is counter based, one thread per element:
//   block(i, a) = ChaCha12(key, words 12..15 = (i lo, i hi, stream, a)); candidates block(i,0)[0..8),
//   block(i,0)[8..16), block(i,1)[0..8), ... (top word masked to 30 bits); the first one < p becomes the element's
//   in-memory (Montgomery) representation.  ALU-pipe only (add / xor / funnel shift); 64 B of stream per element.
// ------------------------------------------------------------------------------------------------
struct rng_key {
    uint32_t k[8];
};
#define PK_QR(a, b, c, d)                                \
    a += b; d ^= a; d = __funnelshift_l(d, d, 16);       \
    c += d; b ^= c; b = __funnelshift_l(b, b, 12);       \
    a += b; d ^= a; d = __funnelshift_l(d, d, 8);        \
    c += d; b ^= c; b = __funnelshift_l(b, b, 7)
__global__ void __launch_bounds__(256) k_rng_fill(fr* out, size_t n, rng_key key, uint32_t stream) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
#pragma unroll
    for (int k = 0; k < 8; k++) in[4 + k] = key.k[k];
    in[12] = (uint32_t)i;
    in[13] = (uint32_t)((uint64_t)i >> 32);
    in[14] = stream;
    for (uint32_t attempt = 0;; attempt++) {
        in[15] = attempt;
        uint32_t x[16];
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = in[k];
#pragma unroll 1
        for (int r = 0; r < 12; r += 2) {
            PK_QR(x[0], x[4], x[8], x[12]);
            PK_QR(x[1], x[5], x[9], x[13]);
            PK_QR(x[2], x[6], x[10], x[14]);
            PK_QR(x[3], x[7], x[11], x[15]);
            PK_QR(x[0], x[5], x[10], x[15]);
            PK_QR(x[1], x[6], x[11], x[12]);
            PK_QR(x[2], x[7], x[8], x[13]);
            PK_QR(x[3], x[4], x[9], x[14]);
        }
        fr c0, c1, t;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            c0.v[k] = x[k] + in[k];
            c1.v[k] = x[8 + k] + in[8 + k];
        }
        c0.v[7] &= 0x3fffffffu;
        c1.v[7] &= 0x3fffffffu;
        if (sub_p(t, c0)) {  // borrow: candidate < p
            fr_store(&out[i], c0);
            return;
        }
        if (sub_p(t, c1)) {
            fr_store(&out[i], c1);
            return;
        }
    }
}
int launch_rng_fill(cudaStream_t st, void* out, size_t n, const uint32_t key[8], uint32_t stream) {
    if (n == 0) return 0;
    rng_key K;
    for (int i = 0; i < 8; i++) K.k[i] = key[i];
    k_rng_fill<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((fr*)out, n, K, stream);
    return 1;
}

cudaError_t init_kernel_attributes() {
    cudaError_t e;
    const int smem = 64 * 1024;
    if ((e = cudaFuncSetAttribute(k_wavelet_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(k_wavelet_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(k_wavelet_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, smem + 32768))) return e;
    return init_ntt_attributes();
}

}  // namespace pk
