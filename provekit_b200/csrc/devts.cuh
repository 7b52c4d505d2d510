// provekit_b200/csrc/devts.cuh — the Fiat-Shamir transcript as device-resident state.
//
// Restates, for ONE device thread, what the host class pkh::ProverState (csrc/host/transcript.cpp) does:
//   * the sponge of provekit/common/src/skyscraper/sponge.rs:24-58 (2 cells, rate 1, permutation =
//     skyscraper::reference::permute, skyscraper/core/src/reference.rs:49-98) under spongefish's duplex rules
//     (overwrite-mode absorb, permute on squeeze; recalled, see DESIGN.md "parity unpinned"),
//   * the codecs: scalars = 32 B canonical LE, challenge bytes = low 15 bytes of one squeezed unit each,
//     add_bytes = one unit per byte, hints = u32-LE length + payload,
//   * the proof string (NARG), appended in HBM right behind the state.
// With the state in HBM the prover's round loops need no host round trip: the kernel that produces a prover message
// (or a one-warp kernel behind it) absorbs it, squeezes the challenge and leaves it in device memory for the next kernel.
// The sponge state is kept CANONICAL (the permutation works on canonical integers; the host version converts on every
// permute): absorb takes a Montgomery element and stores its canonical value, squeeze returns Montgomery form.
#pragma once
#include "skyscraper.cuh"

namespace pk {

struct alignas(16) DevTs {
    fr st[2];                        // canonical integers < p
    uint32_t absorb_pos, squeeze_pos;
    uint32_t narg_words;             // proof string length so far, in 32-bit words (every append is a multiple of 4 B)
    uint32_t narg_cap_words;
    uint32_t error;                  // 0 ok; 1 proof string overflow; 2 zk-sumcheck identity; 3 PoW failure
    uint32_t pad[3];
    // the proof string follows this header
};
static_assert(sizeof(DevTs) == 96, "DevTs layout");
enum { TS_ERR_OVERFLOW = 1, TS_ERR_IDENTITY = 2, TS_ERR_POW = 3 };

// the 18-round Skyscraper-v2 permutation on canonical (l, r) -> canonical (l', r') (reference.rs:49-60); same paired
// Feistel rounds and lazy reduction as sky_compress (register-only round sums: one thread, latency-bound)
static __device__ __noinline__ void sky_permute(fr& l, fr& r) {
#pragma unroll 1
    for (int j = 0; j < 9; j++) {
        const bool is_bar = (j == 3) | (j == 5);
        if (is_bar) {
            r = sky_round_sum<0>(r, sky_bar(sky_canon(l)), 2 * j);
            l = sky_round_sum<0>(l, sky_bar(sky_canon(r)), 2 * j + 1);
        } else {
            r = sky_round_sum<0>(r, fr_sqr_lazy(l), 2 * j);
            l = sky_round_sum<0>(l, fr_sqr_lazy(r), 2 * j + 1);
        }
    }
    l = sky_canon(l);
    r = sky_canon(r);
}

// register-resident working copy of the transcript for the one thread that drives it
struct Ts {
    fr s0, s1;
    uint32_t ap, sp, nw, cap, err;
    DevTs* g;
    uint32_t* narg;
    __device__ __forceinline__ void load(DevTs* t) {
        g = t;
        s0 = fr_load(&t->st[0]);
        s1 = fr_load(&t->st[1]);
        ap = t->absorb_pos;
        sp = t->squeeze_pos;
        nw = t->narg_words;
        cap = t->narg_cap_words;
        err = t->error;
        narg = reinterpret_cast<uint32_t*>(t + 1);
    }
    __device__ __forceinline__ void store() {
        fr_store(&g->st[0], s0);
        fr_store(&g->st[1], s1);
        g->absorb_pos = ap;
        g->squeeze_pos = sp;
        g->narg_words = nw;
        g->error = err;
    }
    __device__ __forceinline__ void put_word(uint32_t w) {
        if (nw < cap)
            narg[nw++] = w;
        else
            err = TS_ERR_OVERFLOW;
    }
    __device__ __forceinline__ void put_canon(const fr& c) {
#pragma unroll
        for (int k = 0; k < 8; k++) put_word(c.v[k]);
    }
    // DuplexSponge::absorb of one unit whose canonical value is c (overwrite mode)
    __device__ __forceinline__ void absorb_unit(const fr& c) {
        if (ap == 1) {
            sky_permute(s0, s1);
            ap = 0;
        }
        s0 = c;
        ap = 1;
        sp = 1;
    }
    // ProverState::add_scalars of one Montgomery-form scalar
    __device__ __forceinline__ void add_scalar(const fr& mont) {
        fr c = fr_from_mont(mont);
        absorb_unit(c);
        put_canon(c);
    }
    // one squeezed unit, canonical
    __device__ __forceinline__ fr squeeze_unit() {
        if (sp == 1) {
            sp = 0;
            ap = 0;
            sky_permute(s0, s1);
        }
        sp = 1;
        return s0;
    }
    // ProverState::challenge_scalars, one scalar, Montgomery form
    __device__ __forceinline__ fr challenge_scalar() { return fr_to_mont(squeeze_unit()); }
    // ProverState::challenge_bytes into 32-bit words (n_bytes need not be a multiple of 4; the tail of the last word is 0)
    __device__ __forceinline__ void challenge_bytes(uint32_t* out_words, int n_bytes) {
        int pos = 0;
        uint32_t cur = 0;  // the word being assembled: stored once per word, not once per byte
        while (pos < n_bytes) {
            fr u = squeeze_unit();
            const int take = n_bytes - pos < 15 ? n_bytes - pos : 15;
            for (int b = 0; b < take; b++, pos++) {
                const uint32_t byte = (u.v[b >> 2] >> (8 * (b & 3))) & 0xffu;
                cur |= byte << (8 * (pos & 3));
                if ((pos & 3) == 3) {
                    out_words[pos >> 2] = cur;
                    cur = 0;
                }
            }
        }
        if (n_bytes & 3) out_words[n_bytes >> 2] = cur;
    }
    // ProverState::add_bytes of 8 bytes (the PoW nonce, big-endian): one unit per byte; the proof string gets the raw bytes
    __device__ __forceinline__ void add_nonce_be(unsigned long long nonce) {
        uint32_t w[2] = {0, 0};
        for (int i = 0; i < 8; i++) {
            const uint32_t byte = (uint32_t)(nonce >> (56 - 8 * i)) & 0xffu;
            fr c = fr_zero();
            c.v[0] = byte;
            absorb_unit(c);
            w[i >> 2] |= byte << (8 * (i & 3));
        }
        put_word(w[0]);
        put_word(w[1]);
    }
};

}  // namespace pk
