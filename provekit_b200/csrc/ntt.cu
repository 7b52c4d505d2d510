// provekit_b200/csrc/ntt.cu — K1 Reed-Solomon encode, the TMA-staged radix-8 kernel (north_star: "radix-4/8 Cooley-Tukey
// butterflies with twiddles and tiles staged through shared memory via TMA").  Replaces [whir] ntt::{expand_from_coeff,
// ntt_batch, transpose} + restructure_evaluations as called from CommitmentWriter::commit_batch
// (provekit/prover/src/whir_r1cs.rs:200-206); same mathematics as k_ntt_pass in kernels.cu (which stays for tiny
// transforms and for the peer-store variant of the sharded commit), different machine mapping:
//
//   * TWO passes over HBM for a 2^17-point column (9 + 8 stage bits) instead of three (6 + 6 + 5): a tile is 2^S points x
//     4 columns (128-byte runs) = 64 KB, two CTAs per SM, so one CTA's loads overlap the other's butterflies.
//   * Tiles and the tile's twiddle slice arrive by TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) and leave by
//     TMA bulk stores (cp.async.bulk.global.shared::cta): no LDG->STS / LDS->STG staging through registers, no address
//     arithmetic on the LSU path.  Between passes the codeword lives in the NEXT pass's tile-major order, so every later
//     pass fetches its tile with ONE bulk copy; only the first pass (coefficient order) and the last (leaf order) move
//     128-byte runs.
//   * Radix-8 in registers: a thread owns the 8 points that differ in three consecutive stage bits of one column: 12
//     butterflies per 8 loads + 8 stores of shared memory and one barrier per three stages (radix-2: 36 accesses, three
//     barriers).  Twiddles of a tile are a precomputed contiguous slice (one bulk copy), 7 reads per radix-8 group.
//   * Warp shuffles are NOT used to move elements between butterfly steps: a 256-bit element is 8 SHFL (4 B per lane and
//     instruction) against 2 LDS.128 + 2 STS.128 through shared memory, i.e. 3x the instructions for the same crossbar
//     bandwidth; shuffles carry the field-sum reductions (kernels.cu block_reduce) where one
//     register per lane moves.
//   * Optionally the codeword is emitted as CANONICAL integers (what the Merkle leaf hash consumes,
//     provekit/common/src/skyscraper/whir.rs:21-24): the conversion rides on the coset twist of the first pass (a table
//     of canonical twiddles) instead of costing one reduction per codeword element in the leaf kernel.
#include "fr.cuh"
#include "kernels.cuh"

namespace pk {

constexpr int NTT8_NCL = 2;             // log2 columns per tile
constexpr int NTT8_NC = 1 << NTT8_NCL;  // 4 columns = 128-byte runs
constexpr int NTT8_RUN_BYTES = NTT8_NC * 32;

// ---- PTX wrappers: mbarrier + bulk async copies (TMA, 1-D) ----
static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
static __device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier
static __device__ __forceinline__ void tma_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global (bulk group)
static __device__ __forceinline__ void tma_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
static __device__ __forceinline__ void tma_store_commit_and_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
static __device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory tile in NATURAL layout (TMA writes plain bytes): element e at byte 32 e, i.e. [point][column] with a 128-byte
// row.  A quarter-warp (8 lanes) touches two rows with four columns each.  With swap = 1 the lanes of the odd row take the
// high 16 bytes first, so the eight 128-bit accesses of one instruction hit 32 distinct banks — at the price of 16 selects per
// element round trip to put the halves back in order.  The kernel is bound by its instruction count, not by shared-memory
// cycles, so it runs with swap = 0 (two-way conflicts, no selects; -DPK_NTT_SWAP restores the conflict-free form for A/B).
static __device__ __forceinline__ fr tile_get(const uint4* tile, int e, int swap) {
    const uint4* p = tile + 2 * e;
    const uint4 first = p[swap], second = p[swap ^ 1];
    const uint4 a = swap ? second : first, b = swap ? first : second;
    fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
static __device__ __forceinline__ void tile_put(uint4* tile, int e, int swap, const fr& x) {
    uint4* p = tile + 2 * e;
    const uint4 a = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]), b = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    p[swap] = swap ? b : a;
    p[swap ^ 1] = swap ? a : b;
}

// Lazy reduction (PK_NTT_LAZY, default on): tile values live in [0, 2p] between butterflies and passes.  a + b is reduced by
// one conditional subtraction of 2p, a - b + 2p in [0, 4p] goes unreduced into the twiddle product, whose Montgomery rows
// accept it and return a value below 2p without the final conditional subtraction (fr.cuh); only the last step of the
// last pass normalises to [0, p).  26 fewer ALU-pipe instructions per butterfly; the codeword is bit-identical.
#ifndef PK_NTT_LAZY
#define PK_NTT_LAZY 1
#endif
#if PK_NTT_LAZY
#define NTT_ADD fr_add_lazy
#define NTT_SUB fr_sub_lazy
#define NTT_MUL fr_mul_lazy
#define NTT_NEG fr_neg_lazy
#else
#define NTT_ADD fr_add
#define NTT_SUB fr_sub
#define NTT_MUL fr_mul
#define NTT_NEG fr_neg
#endif
// one radix-2^G step over stage bits [hb0, hb0 + G) of the tile (DIF: highest bit first); `final`: the last step of the last
// pass (values leave for the leaves: canonical representatives)
template <int G>
static __device__ __forceinline__ void ntt_step(uint4* tile, const fr* tw, const NttR8& P, int hb0, uint32_t Lo, bool twist,
                                                uint32_t s, uint32_t qbase, bool final) {
    constexpr int R = 1 << G;
    const int items = (1 << (P.S - G)) << NTT8_NCL;
    for (int item = threadIdx.x; item < items; item += blockDim.x) {
        const int k = item & (NTT8_NC - 1), grp = item >> NTT8_NCL;
#ifdef PK_NTT_SWAP
        const int swap = grp & 1;
#else
        const int swap = 0;  // see tile_get: the un-swapping selects cost more than the two-way bank conflict they avoid
#endif
        const int low = grp & ((1 << hb0) - 1);
        const int e0 = ((grp >> hb0) << (hb0 + G)) | low;
        fr x[R];
#pragma unroll
        for (int c = 0; c < R; c++) x[c] = tile_get(tile, ((e0 + (c << hb0)) << NTT8_NCL) + k, swap);
        if (twist) {
            // coset twist omega_M^(s q) on the coefficients (first pass, first step); with canonical output the table holds
            // canonical twiddles, so the Montgomery product strips the factor R, and untwisted elements are reduced once
            const uint32_t halfM = 1u << (P.logM - 1);
#pragma unroll
            for (int c = 0; c < R; c++) {
                const uint32_t q = qbase | ((uint32_t)(e0 + (c << hb0)) << P.l);
                uint32_t e = (s * q) & ((halfM << 1) - 1u);
                const bool neg = e >= halfM;
                e &= halfM - 1u;
                if (e)
                    x[c] = NTT_MUL(x[c], fr_load_nc(&P.Wtwist[(size_t)e << P.tbl_shift]));
                else if (P.canonical)
                    x[c] = fr_from_mont(x[c]);
                if (neg) x[c] = NTT_NEG(x[c]);
            }
        }
#pragma unroll
        for (int sbit = G - 1; sbit >= 0; sbit--) {
            const int hb = hb0 + sbit;
#pragma unroll
            for (int c = 0; c < R; c++) {
                if (c & (1 << sbit)) continue;
                const int c1 = c | (1 << sbit);
                const int j = low | ((c & ((1 << sbit) - 1)) << hb0);  // position inside the half-size-2^hb butterfly group
                const fr a = x[c], b = x[c1];
                x[c] = NTT_ADD(a, b);
                fr d = NTT_SUB(a, b);
                if ((uint32_t)j | Lo)
                    d = NTT_MUL(d, fr_load(&tw[(1 << hb) - 1 + j]));
#if PK_NTT_LAZY
                else
                    d = fr_reduce_2p_once(d);  // (j, Lo) = (0, 0): twiddle 1, only the range has to come back to [0, 2p]
#endif
                x[c1] = d;
            }
        }
#if PK_NTT_LAZY
        if (final) {
#pragma unroll
            for (int c = 0; c < R; c++) x[c] = fr_normalize_2p(x[c]);
        }
#endif
#pragma unroll
        for (int c = 0; c < R; c++) tile_put(tile, ((e0 + (c << hb0)) << NTT8_NCL) + k, swap, x[c]);
    }
}

// element offset (in field elements) of (coset s, column group cg, point q) in the tile-major order of a pass (l, S)
static __device__ __forceinline__ size_t tile_major(uint32_t s, uint32_t cg, uint32_t q, int L, int l, int S) {
    const uint32_t lo = q & ((1u << l) - 1u), mid = (q >> l) & ((1u << S) - 1u), hi = q >> (l + S);
    const size_t tile_id = ((size_t)hi << l) | lo;
    return ((((((size_t)s * (16 / NTT8_NC) + cg) << (L - S)) + tile_id) << S) + mid) << NTT8_NCL;
}

__global__ void __launch_bounds__(256, 2) k_ntt_r8(NttR8 P) {
    extern __shared__ __align__(128) uint8_t smem_bytes[];
    const int S = P.S, l = P.l, L = P.L;
    const int npts = 1 << S;
    uint4* tile = reinterpret_cast<uint4*>(smem_bytes);
    fr* tw = reinterpret_cast<fr*>(smem_bytes + (size_t)npts * NTT8_RUN_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_bytes + (size_t)npts * NTT8_RUN_BYTES + (size_t)npts * 32);
    // consecutive blocks = the 4 column groups x all cosets of the same rows: the 128-byte runs of the column groups are
    // neighbours in HBM, and every coset re-reads the same coefficients, so the first reader pulls them from HBM and the
    // others (scheduled right next to it) hit L2
    const uint32_t cg = blockIdx.x & (16 / NTT8_NC - 1);
    const uint32_t tile_lin = blockIdx.x / (16 / NTT8_NC);
    const uint32_t s = tile_lin & ((1u << P.logE) - 1u);
    const uint32_t t = tile_lin >> P.logE;
    const uint32_t Lo = t & ((1u << l) - 1u), H = t >> l;
    const uint32_t qbase = (H << (l + S)) | Lo;
    const uint32_t tile_bytes = (uint32_t)npts * NTT8_RUN_BYTES, tw_bytes = (uint32_t)npts * 32;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_arrive_expect_tx(bar, tile_bytes + tw_bytes);
    }
    __syncthreads();
    if (P.first) {
        // coefficient order: point q holds its 16 columns in one 512-byte row; this tile takes a 128-byte run per point
        for (int mid = threadIdx.x; mid < npts; mid += blockDim.x) {
            const uint32_t q = qbase | ((uint32_t)mid << l);
            tma_load(tile + 2 * (mid << NTT8_NCL), P.in + ((size_t)q * 16 + cg * NTT8_NC), NTT8_RUN_BYTES, bar);
        }
    } else if (threadIdx.x == 0) {
        tma_load(tile, P.in + tile_major(s, cg, qbase, L, l, S), tile_bytes, bar);  // the whole tile is one contiguous block
    }
    if (threadIdx.x == 32 || (blockDim.x <= 32 && threadIdx.x == 0)) tma_load(tw, P.tw + ((size_t)Lo << S), tw_bytes, bar);
    mbar_wait(bar, 0);

    bool twist = P.first != 0 && (s != 0 || P.canonical);
    for (int hb = S; hb > 0;) {  // stage bits [hb - g, hb)
        const int g = hb >= 3 ? 3 : hb;
        const bool final = P.last != 0 && hb == g;
        if (g == 3)
            ntt_step<3>(tile, tw, P, hb - 3, Lo, twist, s, qbase, final);
        else if (g == 2)
            ntt_step<2>(tile, tw, P, hb - 2, Lo, twist, s, qbase, final);
        else
            ntt_step<1>(tile, tw, P, hb - 1, Lo, twist, s, qbase, final);
        twist = false;
        hb -= g;
        if (hb > 0) __syncthreads();
    }
    // results leave by bulk stores: make the generic-proxy writes visible to the async proxy, then one run per point
    fence_async_smem();
    __syncthreads();
    for (int mid = threadIdx.x; mid < npts; mid += blockDim.x) {
        const uint32_t q = qbase | ((uint32_t)mid << l);
        fr* dst;
        if (P.last) {
            const uint32_t tq = __brev(q) >> (32 - L);
            const size_t row = (size_t)s + ((size_t)tq << P.logE);
            dst = P.out + row * P.leaf_stride + P.col_offset + cg * NTT8_NC;
        } else {
            dst = P.out + tile_major(s, cg, q, L, P.l_next, P.S_next);
        }
        tma_store(dst, tile + 2 * (mid << NTT8_NCL), NTT8_RUN_BYTES);
    }
    tma_store_commit_and_wait();  // shared memory must outlive the reads of the bulk stores
}

// per-pass twiddle slices: out[Lo * 2^S + t], t = 2^hb - 1 + j  ->  omega_(2^L)^( ((j << l) | Lo) << (L - 1 - hb - l) )
__global__ void __launch_bounds__(256) k_ntt_tile_twiddles(fr* out, const fr* __restrict__ W, int w_shift, int L, int l, int S) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << (l + S))) return;
    const uint32_t t = (uint32_t)i & ((1u << S) - 1u), Lo = (uint32_t)(i >> S);
    fr v = fr_one();
    if (t < (1u << S) - 1u) {
        const int hb = 31 - __clz(t + 1);
        const uint32_t j = t + 1 - (1u << hb);
        const uint32_t e = ((j << l) | Lo) << (L - 1 - hb - l);
        if (e) v = fr_load_nc(&W[(size_t)e << w_shift]);
    }
    fr_store(&out[i], v);
}
int launch_ntt_tile_twiddles(cudaStream_t st, void* out, const void* W, int table_log_m, int L, int l, int S) {
    const size_t n = (size_t)1 << (l + S);
    k_ntt_tile_twiddles<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((fr*)out, (const fr*)W, table_log_m - L, L, l, S);
    return 1;
}

size_t ntt_r8_smem_bytes(int S) { return ((size_t)NTT8_RUN_BYTES << S) + ((size_t)32 << S) + 16; }

int launch_ntt_r8_pass(cudaStream_t st, const NttR8& P) {
    const unsigned grid = (1u << (P.logE + P.L - P.S)) * (16 / NTT8_NC);
    const int items = (1 << (P.S - 3 > 0 ? P.S - 3 : 0)) << NTT8_NCL;
    int threads = items < 32 ? 32 : (items > 256 ? 256 : items);
    k_ntt_r8<<<grid, threads, ntt_r8_smem_bytes(P.S), st>>>(P);
    return 1;
}
cudaError_t init_ntt_attributes() {
    return cudaFuncSetAttribute(k_ntt_r8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_r8_smem_bytes(NTT8_MAX_S));
}

}  // namespace pk
