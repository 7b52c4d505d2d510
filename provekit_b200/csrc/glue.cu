// provekit_b200/csrc/glue.cu — the small kernels that keep the protocol flow of WhirR1CSProver::prove on the device:
// Fiat-Shamir exchanges on the device-resident transcript (devts.cuh), the scalar bookkeeping of the zk-sumcheck
// (provekit/prover/src/whir_r1cs.rs:103-171, 280-345), STIR index derivation, PoW grinding without a host loop, and the
// serialisation of the hints (STIR answers, ark MultiPath, claimed / deferred evaluations) straight into device memory.
// Everything here is launch-latency work (one thread, one warp or one block); the data-parallel kernels are in kernels.cu.
#include "devts.cuh"
#include "kernels.cuh"

namespace pk {

static __device__ __forceinline__ fr arg_fr(const fr_arg& a) {
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = a.v[i];
    return r;
}

// ------------------------------------------------------------------------------------------------
// transcript operations
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_ts_init(DevTs* t, fr_arg iv_canonical, uint32_t cap_words) {
    if (threadIdx.x) return;
    fr_store(&t->st[0], fr_zero());
    fr_store(&t->st[1], arg_fr(iv_canonical));
    t->absorb_pos = 0;
    t->squeeze_pos = 1;
    t->narg_words = 0;
    t->narg_cap_words = cap_words;
    t->error = 0;
}
int launch_ts_init(cudaStream_t st, void* ts, fr_arg iv_canonical, uint32_t cap_words) {
    k_ts_init<<<1, 32, 0, st>>>((DevTs*)ts, iv_canonical, cap_words);
    return 1;
}

// add_scalars(absorb[0..na)) then challenge_scalars -> squeeze[0..ns); `canonical`: the absorbed values are canonical
// integers already (a Merkle root), not Montgomery-form elements
__global__ void __launch_bounds__(32) k_ts_exchange(DevTs* t, const fr* absorb, int na, int canonical, fr* squeeze, int ns) {
    if (threadIdx.x) return;
    Ts ts;
    ts.load(t);
    for (int i = 0; i < na; i++) {
        fr x = fr_load(&absorb[i]);
        if (canonical) {
            ts.absorb_unit(x);
            ts.put_canon(x);
        } else {
            ts.add_scalar(x);
        }
    }
    for (int i = 0; i < ns; i++) fr_store(&squeeze[i], ts.challenge_scalar());
    ts.store();
}
int launch_ts_exchange(cudaStream_t st, void* ts, const void* absorb, int na, bool canonical, void* squeeze, int ns) {
    k_ts_exchange<<<1, 32, 0, st>>>((DevTs*)ts, (const fr*)absorb, na, canonical ? 1 : 0, (fr*)squeeze, ns);
    return 1;
}

__global__ void __launch_bounds__(32) k_ts_challenge_bytes(DevTs* t, uint32_t* out_words, int n_bytes) {
    if (threadIdx.x) return;
    Ts ts;
    ts.load(t);
    ts.challenge_bytes(out_words, n_bytes);
    ts.store();
}
int launch_ts_challenge_bytes(cudaStream_t st, void* ts, uint32_t* out_words, int n_bytes) {
    k_ts_challenge_bytes<<<1, 32, 0, st>>>((DevTs*)ts, out_words, n_bytes);
    return 1;
}

// ProverState::hint: u32-LE length + payload appended to the proof string (not absorbed).  len_ptr (device) overrides len.
__global__ void __launch_bounds__(256) k_ts_hint(DevTs* t, const uint32_t* payload, uint32_t len_words, const uint32_t* len_ptr) {
    __shared__ uint32_t base_s, ok_s;
    const uint32_t n = len_ptr ? *len_ptr : len_words;
    if (threadIdx.x == 0) {
        const uint32_t nw = t->narg_words;
        ok_s = nw + 1 + n <= t->narg_cap_words;
        base_s = nw;
        if (ok_s) {
            reinterpret_cast<uint32_t*>(t + 1)[nw] = 4u * n;  // byte length
            t->narg_words = nw + 1 + n;
        } else {
            t->error = TS_ERR_OVERFLOW;
        }
    }
    __syncthreads();
    if (!ok_s) return;
    uint32_t* dst = reinterpret_cast<uint32_t*>(t + 1) + base_s + 1;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = payload[i];
}
int launch_ts_hint(cudaStream_t st, void* ts, const uint32_t* payload, uint32_t len_words, const uint32_t* len_ptr) {
    k_ts_hint<<<1, 256, 0, st>>>((DevTs*)ts, payload, len_words, len_ptr);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// zk-sumcheck bookkeeping (run_zk_sumcheck_prover, whir_r1cs.rs:228-369).  compute_blinding_coefficients_for_round
// (:103-171) costs O(m_0) per round on the host; here the prefix is carried and the suffix sums are built once:
//   round idx: prefix = sum_{i<idx} g_i(alpha_i),  suffix[idx] = sum_{i>idx} (g_i(0) + g_i(1)),  pm = 2^(m_0-1-idx)
//   gp = pm * g_idx + (pm * prefix + pm/2 * suffix[idx])  on the constant coefficient
// Field arithmetic is exact, so the regrouping gives the reference's values bit for bit.
// ------------------------------------------------------------------------------------------------
static __device__ __forceinline__ fr eval_cubic(const fr c[4], const fr& x) {  // sumcheck.rs:174-176
    return fr_add(c[0], fr_mul(x, fr_add(c[1], fr_mul(x, fr_add(c[2], fr_mul(x, c[3]))))));
}
static __device__ __forceinline__ fr pow2_fr(int k) {  // 2^k, Montgomery form
    fr x = fr_one();
    for (int i = 0; i < k; i++) x = fr_dbl(x);
    return x;
}
// gp[0..4) of round idx
static __device__ __forceinline__ void blinding_round(const ZkGlue& G, int idx, const fr& prefix, const fr& half, fr gp[4]) {
    const fr pm = pow2_fr(G.m0 - 1 - idx);
    const fr sm = fr_mul(pm, half);
    const fr cst = fr_add(fr_mul(pm, prefix), fr_mul(sm, fr_load(&G.suffix[idx])));
#pragma unroll
    for (int k = 0; k < 4; k++) gp[k] = fr_mul(pm, fr_load(&G.blind[4 * idx + k]));
    gp[0] = fr_add(gp[0], cst);
}
__global__ void __launch_bounds__(32) k_zk_init(ZkGlue G, fr_arg half_arg) {
    if (threadIdx.x) return;
    fr acc = fr_zero();
    for (int i = G.m0 - 1; i >= 0; i--) {
        fr_store(&G.suffix[i], acc);
        fr c0 = fr_load(&G.blind[4 * i]);
        fr s = fr_add(fr_dbl(c0), fr_add(fr_load(&G.blind[4 * i + 1]), fr_add(fr_load(&G.blind[4 * i + 2]), fr_load(&G.blind[4 * i + 3]))));
        acc = fr_add(acc, s);
    }
    fr gp[4];
    blinding_round(G, 0, fr_zero(), arg_fr(half_arg), gp);
    // sum_over_hypercube (whir_r1cs.rs:173-180): g(0) + g(1) of the round-0 polynomial
    fr_store(&G.state[3], fr_add(fr_dbl(gp[0]), fr_add(gp[1], fr_add(gp[2], gp[3]))));
}
int launch_zk_init(cudaStream_t st, ZkGlue g, fr_arg half) {
    k_zk_init<<<1, 32, 0, st>>>(g, half);
    return 1;
}
// state: [0] rho, [1] saved, [2] prefix, [3] sum_g.  Round idx: consumes h3 (the kernel's f(0), f(-1), f(inf)) and, for
// idx > 0, alpha[idx-1]; produces cf[0..4).  With a device transcript it also sends cf and draws alpha[idx].
__global__ void __launch_bounds__(32) k_zk_glue(ZkGlue G, int idx, fr_arg half_arg, DevTs* t) {
    if (threadIdx.x) return;
    const fr half = arg_fr(half_arg);
    const fr rho = fr_load(&G.state[0]);
    fr saved, prefix;
    if (idx == 0) {
        saved = fr_mul(rho, fr_load(&G.state[3]));
        prefix = fr_zero();
    } else {
        fr cfp[4], gprev[4];
        const fr a = fr_load(&G.alpha[idx - 1]);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            cfp[k] = fr_load(&G.cf[k]);
            gprev[k] = fr_load(&G.blind[4 * (idx - 1) + k]);
        }
        saved = eval_cubic(cfp, a);
        prefix = fr_add(fr_load(&G.state[2]), eval_cubic(gprev, a));
    }
    fr gp[4], cf[4];
    blinding_round(G, idx, prefix, half, gp);
    const fr h0 = fr_load(&G.h3[0]), h1 = fr_load(&G.h3[1]), h2 = fr_load(&G.h3[2]);
    cf[0] = fr_add(h0, fr_mul(rho, gp[0]));
    const fr g_m1 = fr_sub(fr_add(fr_sub(gp[0], gp[1]), gp[2]), gp[3]);
    const fr c_m1 = fr_add(h1, fr_mul(rho, g_m1));
    cf[2] = fr_mul(half, fr_sub(fr_sub(fr_sub(fr_add(saved, c_m1), cf[0]), cf[0]), cf[0]));
    cf[3] = fr_add(h2, fr_mul(rho, gp[3]));
    cf[1] = fr_sub(fr_sub(fr_sub(fr_sub(saved, cf[0]), cf[0]), cf[3]), cf[2]);
    fr_store(&G.state[1], saved);
    fr_store(&G.state[2], prefix);
#pragma unroll
    for (int k = 0; k < 4; k++) fr_store(&G.cf[k], cf[k]);
    if (t) {
        Ts ts;
        ts.load(t);
#pragma unroll 1
        for (int k = 0; k < 4; k++) ts.add_scalar(cf[k]);
        fr_store(&G.alpha[idx], ts.challenge_scalar());
        ts.store();
    }
}
int launch_zk_glue(cudaStream_t st, ZkGlue g, int idx, fr_arg half, void* ts) {
    k_zk_glue<<<1, 32, 0, st>>>(g, idx, half, (DevTs*)ts);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// small vector helpers
// ------------------------------------------------------------------------------------------------
static __device__ __forceinline__ fr fr_pow_u32(const fr& b, uint32_t e) {
    fr r = fr_one(), x = b;
    while (e) {
        if (e & 1u) r = fr_mul(r, x);
        e >>= 1;
        if (e) x = fr_sqr(x);
    }
    return r;
}
// out[j] = base^(first_exp + j) for j < n; with count_ptr: entries j >= *count_ptr + (first_exp == 0) are zero (the
// combination scalars of a STIR batch whose size is only known on the device)
__global__ void __launch_bounds__(128) k_powers(fr* out, const fr* base, int n, int first_exp, const uint32_t* count_ptr) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int limit = count_ptr ? (int)*count_ptr + (first_exp == 0 ? 1 : 0) : n;
    fr_store(&out[j], j < limit ? fr_pow_u32(fr_load(base), (uint32_t)(first_exp + j)) : fr_zero());
}
int launch_powers(cudaStream_t st, void* out, const void* base, int n, int first_exp, const uint32_t* count_ptr) {
    if (n <= 0) return 0;
    k_powers<<<(n + 127) / 128, 128, 0, st>>>((fr*)out, (const fr*)base, n, first_exp, count_ptr);
    return 1;
}
// expand_powers (whir_r1cs.rs:347-377): out[4i + k] = alpha_i^k, k < 4, i < m0; out zero beyond (len elements in all)
__global__ void __launch_bounds__(128) k_expand_powers(fr* out, const fr* alpha, int m0, size_t len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    fr v = fr_zero();
    if (i < 4 * (size_t)m0) v = fr_pow_u32(fr_load(&alpha[i >> 2]), (uint32_t)(i & 3));
    fr_store(&out[i], v);
}
int launch_expand_powers(cudaStream_t st, void* out, const void* alpha, int m0, size_t len) {
    k_expand_powers<<<(unsigned)((len + 127) / 128), 128, 0, st>>>((fr*)out, (const fr*)alpha, m0, len);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// STIR queries [whir get_challenge_stir_queries; recursive-verifier/app/circuit/whir_utilities.go:48-77]: nq big-endian
// integers of nb bytes each, reduced mod 2^folded_log, sorted ascending, de-duplicated
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_stir_indices(const uint32_t* bytes_words, int nq, int nb, int folded_log, uint64_t* idx,
                                                     uint32_t* n_idx) {
    __shared__ uint64_t s[OPEN_MAX_QUERIES];  // the sort runs in shared memory: global read-modify-writes cost ~1 us each
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(bytes_words);
    const uint64_t mask = folded_log >= 64 ? ~0ull : (((uint64_t)1 << folded_log) - 1);
    for (int i = threadIdx.x; i < nq; i += 32) {
        uint64_t v = 0;
        for (int j = 0; j < nb; j++) v = (v << 8) | bytes[i * nb + j];
        s[i] = v & mask;
    }
    __syncwarp();
    // rank sort without duplicates: an element survives if no EARLIER element has the same value; its position is the
    // number of surviving smaller values
    __shared__ uint8_t first[OPEN_MAX_QUERIES];
    for (int i = threadIdx.x; i < nq; i += 32) {
        const uint64_t v = s[i];
        bool f = true;
        for (int j = 0; j < i; j++) f &= s[j] != v;
        first[i] = f;
    }
    __syncwarp();
    int mine = 0;
    for (int i = threadIdx.x; i < nq; i += 32) {
        if (!first[i]) continue;
        mine++;
        const uint64_t v = s[i];
        int rank = 0;
        for (int j = 0; j < nq; j++) rank += (first[j] && s[j] < v) ? 1 : 0;
        idx[rank] = v;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
    if (threadIdx.x == 0) *n_idx = (uint32_t)mine;
}
int launch_stir_indices(cudaStream_t st, const uint32_t* bytes_words, int nq, int nb, int folded_log, uint64_t* idx, uint32_t* n_idx) {
    k_stir_indices<<<1, 32, 0, st>>>(bytes_words, nq, nb, folded_log, idx, n_idx);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// STIR answers + ark MultiPath, serialised on the device exactly as the reference's hints are
// (ark-serialize compressed: u64-LE lengths, 32-byte canonical LE scalars; MultiPath layout
// recursive-verifier/app/circuit/mt.go:36-50, app/utilities/utilities.go:71-82):
//   hint1 "stir_answers": Vec<Vec<F>>      = n | n x ( w | w scalars )
//   hint2 "merkle_proof": MultiPath        = n | n leaf-sibling digests | n | n prefix lengths | n | n x ( len | suffix
//                                            digests, root -> leaf ) | n | n leaf indexes
// Prefix lengths come from comparing DIGESTS with the previous path, as ark's prefix_encode_path does.
// ------------------------------------------------------------------------------------------------
static __device__ __forceinline__ void put_u64(uint32_t* dst, uint64_t v) {
    dst[0] = (uint32_t)v;
    dst[1] = (uint32_t)(v >> 32);
}
__global__ void __launch_bounds__(256) k_open_hints(const fr* __restrict__ leaves, int leaves_canonical, int w,
                                                    const fr* __restrict__ nodes, size_t L, int depth, const uint64_t* idx,
                                                    const uint32_t* n_idx, uint32_t* hint1, uint32_t* hint2, uint32_t* lens) {
    __shared__ uint32_t pref[OPEN_MAX_QUERIES], soff[OPEN_MAX_QUERIES + 1];
    const int n = (int)*n_idx;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int plen = depth - 1;
    // ---- hint 1 ----
    if (tid == 0) put_u64(hint1, (uint64_t)n);
    const int per_q = 2 + 8 * w;
    for (int e = tid; e < n * w; e += nt) {
        const int q = e / w, k = e % w;
        fr x = fr_load_nc(&leaves[idx[q] * (size_t)w + k]);
        if (!leaves_canonical) x = fr_from_mont(x);
        uint32_t* dst = hint1 + 2 + (size_t)q * per_q;
        if (k == 0) put_u64(dst, (uint64_t)w);
        dst += 2 + 8 * k;
#pragma unroll
        for (int j = 0; j < 8; j++) dst[j] = x.v[j];
    }
    // ---- hint 2: prefix lengths ----
    for (int q = tid; q < n; q += nt) {
        int k = 0;
        if (q > 0) {
            const size_t a = L + idx[q - 1], b = L + idx[q];
            // root -> leaf element j is the sibling at level depth-1-j
            for (; k < plen; k++) {
                const int lv = depth - 1 - k;
                const uint4* x = reinterpret_cast<const uint4*>(&nodes[(a >> lv) ^ 1]);
                const uint4* y = reinterpret_cast<const uint4*>(&nodes[(b >> lv) ^ 1]);
                const uint4 x0 = x[0], x1 = x[1], y0 = y[0], y1 = y[1];
                const bool same = x0.x == y0.x && x0.y == y0.y && x0.z == y0.z && x0.w == y0.w && x1.x == y1.x && x1.y == y1.y &&
                                  x1.z == y1.z && x1.w == y1.w;
                if (!same) break;
            }
        }
        pref[q] = (uint32_t)k;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int q = 0; q < n; q++) {
            soff[q] = acc;
            acc += (uint32_t)plen - pref[q];
        }
        soff[n] = acc;
        lens[0] = 2u + (uint32_t)n * (uint32_t)per_q;
        lens[1] = (2u + 8u * n) + (2u + 2u * n) + (2u + 2u * n + 8u * acc) + (2u + 2u * n);
    }
    __syncthreads();
    uint32_t* sec1 = hint2;                  // n | siblings
    uint32_t* sec2 = sec1 + 2 + 8 * n;       // n | prefix lengths
    uint32_t* sec3 = sec2 + 2 + 2 * n;       // n | per path: len | suffix
    uint32_t* sec4 = sec3 + 2 + 2 * n + 8 * soff[n];  // n | indexes
    if (tid == 0) {
        put_u64(sec1, (uint64_t)n);
        put_u64(sec2, (uint64_t)n);
        put_u64(sec3, (uint64_t)n);
        put_u64(sec4, (uint64_t)n);
    }
    for (int q = tid; q < n; q += nt) {
        put_u64(sec2 + 2 + 2 * q, (uint64_t)pref[q]);
        put_u64(sec4 + 2 + 2 * q, idx[q]);
        put_u64(sec3 + 2 + 2 * q + 8 * soff[q], (uint64_t)((uint32_t)plen - pref[q]));
    }
    // digests: one (q, level) pair per thread; level d = 0 is the leaf sibling, j = 0..plen-1 the auth path root -> leaf
    for (int e = tid; e < n * depth; e += nt) {
        const int q = e / depth, d = e % depth;
        const size_t pos = ((L + idx[q]) >> d) ^ 1;
        uint32_t* dst;
        if (d == 0) {
            dst = sec1 + 2 + 8 * q;
        } else {
            const int j = depth - 1 - d;  // position in the root -> leaf path
            if (j < (int)pref[q]) continue;
            dst = sec3 + 2 + 2 * q + 8 * soff[q] + 2 + 8 * (j - (int)pref[q]);
        }
        const uint4* src = reinterpret_cast<const uint4*>(&nodes[pos]);
        const uint4 a = src[0], b = src[1];
        dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
        dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
    }
}
int launch_open_hints(cudaStream_t st, const void* leaves, bool leaves_canonical, int w, const void* nodes, size_t L, int depth,
                      const uint64_t* idx, const uint32_t* n_idx, uint32_t* hint1, uint32_t* hint2, uint32_t* lens) {
    k_open_hints<<<1, 256, 0, st>>>((const fr*)leaves, leaves_canonical ? 1 : 0, w, (const fr*)nodes, L, depth, idx, n_idx, hint1,
                                    hint2, lens);
    return 1;
}

// hint of scalars: n_groups x ( u64 group_len | group_len canonical scalars ); element e of group g = src[e*estride + g*gstride]
// (claimed_evaluations: (Vec<F>, Vec<F>) from the interleaved <w_j, f>, <w_j, g> sums; deferred_weight_evaluations: Vec<F>)
__global__ void __launch_bounds__(32) k_hint_scalars(uint32_t* out, const fr* src, int n_groups, int group_len, int estride, int gstride) {
    if (threadIdx.x) return;
    uint32_t* p = out;
    for (int g = 0; g < n_groups; g++) {
        put_u64(p, (uint64_t)group_len);
        p += 2;
        for (int e = 0; e < group_len; e++) {
            fr c = fr_from_mont(fr_load(&src[e * estride + g * gstride]));
#pragma unroll
            for (int j = 0; j < 8; j++) p[j] = c.v[j];
            p += 8;
        }
    }
}
int launch_hint_scalars(cudaStream_t st, uint32_t* out, const void* src, int n_groups, int group_len, int estride, int gstride) {
    k_hint_scalars<<<1, 32, 0, st>>>(out, (const fr*)src, n_groups, group_len, estride, gstride);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// PoW without a host loop (skyscraper/core/src/pow.rs:24-41, generic.rs:42-71): blocks draw batches of nonces from a
// ticket counter until a batch starts above the best accepted nonce; the smallest accepted nonce wins (fetch_min).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_pow_begin(DevTs* t, PowCtrl* c) {
    if (threadIdx.x) return;
    Ts ts;
    ts.load(t);
    ts.challenge_bytes(c->challenge, 32);  // challenge_pow: 32 challenge bytes
    ts.store();
    c->ticket = 0;
    c->best = ~0ull;
}
__global__ void __launch_bounds__(128) k_pow_grind(PowCtrl* c, fr_arg threshold) {
    __shared__ unsigned long long base_s, best_s;
    fr l;
#pragma unroll
    for (int i = 0; i < 8; i++) l.v[i] = c->challenge[i];
    l = sky_reduce(l);
    const fr thr = arg_fr(threshold);
    volatile unsigned long long* best = &c->best;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            base_s = atomicAdd(&c->ticket, (unsigned long long)blockDim.x);
            best_s = *best;
        }
        __syncthreads();
        const unsigned long long base = base_s;
        if (best_s < base) return;  // every nonce of this and all later batches is above an accepted one
        const unsigned long long nonce = base + threadIdx.x;
        if (*best < nonce) continue;
        fr r = fr_zero();
        r.v[0] = (uint32_t)nonce;
        r.v[1] = (uint32_t)(nonce >> 32);
        bool done;
        fr h = sky_compress_while(l, r, [&](int j) { return j == 0 || *best >= nonce; }, done);
        if (!done) continue;
        bool less = false;
#pragma unroll
        for (int i = 7; i >= 0; i--) {
            if (h.v[i] != thr.v[i]) {
                less = h.v[i] < thr.v[i];
                break;
            }
        }
        if (less) atomicMin(&c->best, nonce);
    }
}
__global__ void __launch_bounds__(32) k_pow_end(DevTs* t, PowCtrl* c) {
    if (threadIdx.x) return;
    Ts ts;
    ts.load(t);
    ts.add_nonce_be(c->best);
    ts.store();
}
int launch_pow_begin(cudaStream_t st, void* ts, PowCtrl* c) {
    k_pow_begin<<<1, 32, 0, st>>>((DevTs*)ts, c);
    return 1;
}
int launch_pow_grind(cudaStream_t st, PowCtrl* c, fr_arg threshold, int blocks) {
    k_pow_grind<<<blocks, 128, 0, st>>>(c, threshold);
    return 1;
}
int launch_pow_end(cudaStream_t st, void* ts, PowCtrl* c) {
    k_pow_end<<<1, 32, 0, st>>>((DevTs*)ts, c);
    return 1;
}

}  // namespace pk
