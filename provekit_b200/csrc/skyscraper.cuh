// provekit_b200/csrc/skyscraper.cuh — Skyscraper-v2 two-to-one compression on the B200 integer pipe.
//
// Replaces skyscraper::simple::compress (skyscraper/core/src/simple.rs:12-14 -> generic.rs:77-102) as
// called by the Merkle plug-in (provekit/common/src/skyscraper/whir.rs:20-25) and the PoW solver
// (skyscraper/core/src/generic.rs:42-71).  Spec: skyscraper/core/src/reference.rs:41-98.
// 18 Feistel rounds (l, r) <- (r + F_i(l) + rc_i, l); F = x^2 * 2^-256 (a raw Montgomery square, the
// block_multiplier::scalar_sqr semantics) except rounds 6,7,10,11 where F = bar.  All values are raw
// canonical integers in [0, p); the reference keeps lazily reduced representatives, which only `bar`
// could observe and it canonicalises first (bar.rs:17), so outputs are bit-identical.
#pragma once
#include "fr.cuh"

namespace pk {

// skyscraper/core/src/constants.rs:32-51 as 32-bit little-endian limbs
__constant__ uint32_t SKY_RC[18][8] = {
    {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0x270bd744u, 0x903c4324u, 0x08a7d269u, 0x873125f7u, 0x06c83855u, 0x081dd279u, 0xea6d7667u, 0x276b1823u},
    {0x4b378d71u, 0x7ac8edbbu, 0xd99e2cb7u, 0xe29d79f3u, 0x4c1a5a18u, 0x75141791u, 0x58a484a6u, 0x0cf02bd7u},
    {0x69e5bc36u, 0xfa7adc67u, 0x7cca387du, 0x1c3f8e29u, 0x63481db0u, 0x0eb7730du, 0x18ede544u, 0x25b0e03fu},
    {0x2f03cfb7u, 0x57847e65u, 0x68873404u, 0x33440b96u, 0x49af80bcu, 0x955a32e8u, 0xbe14ae70u, 0x002882fcu},
    {0x6257d4d7u, 0x97923139u, 0x1b37d3c1u, 0x29989c3eu, 0x7f1277bau, 0x12ef02b4u, 0x1e2b7a9cu, 0x039ad857u},
    {0xabbb7887u, 0xb5b48465u, 0xe6ba2d2bu, 0xa72a6bc5u, 0x712f7b29u, 0x4cd48043u, 0x0fc1fc1au, 0x1142d541u},
    {0x059075d3u, 0x7ab2c156u, 0x047999b2u, 0x17cb3594u, 0x98f289f7u, 0x44f2c935u, 0x69bc0becu, 0x1d78439fu},
    {0x138b8edbu, 0x05d7a965u, 0xd55c48b1u, 0x36ef35a3u, 0xac6f1628u, 0x8ddfb8a1u, 0x08f4ff82u, 0x258588a5u},
    {0xfccb49e9u, 0x1596fb9au, 0x9a09a95bu, 0x9a7367d6u, 0x84e4c157u, 0x9bc43f69u, 0xd2f514feu, 0x13087879u},
    {0x3b4109fau, 0x295ccd23u, 0xed868012u, 0xe1d72f89u, 0x4bc88a8eu, 0x2e9e1eeau, 0x98c45232u, 0x17dadee8u},
    {0xaa1f486fu, 0x9a8590b4u, 0x30e9130eu, 0xb75834b4u, 0x34d5de31u, 0xb8e90b10u, 0x46e7f4a6u, 0x295c6d15u},
    {0x4c6eb892u, 0x850adcb7u, 0x05b92fc3u, 0x07699ef3u, 0xa1720f2du, 0x4ef96a2bu, 0x1d3ed446u, 0x1288ca0eu},
    {0x49d1b5eeu, 0x01960f93u, 0x69371c69u, 0x8ccad307u, 0x91c98662u, 0xe5c81e89u, 0x1ae023f3u, 0x17563b4du},
    {0x76b32917u, 0x6ba01e94u, 0xdd977bc9u, 0xa1cb0a3au, 0x5815f030u, 0x86815a94u, 0xe91a1eeau, 0x2869043bu},
    {0x5511d976u, 0x81776c88u, 0x47f414e7u, 0x7475d34fu, 0x095d96cfu, 0x5d090056u, 0xff59e79au, 0x14941f0au},
    {0x8fc8c034u, 0xbc40b4fdu, 0xcce4fd48u, 0xbb7142c3u, 0x8a39005au, 0x31835675u, 0x90f4379fu, 0x1ce337a1u},
    {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
};

// per-byte sbox on 4 packed bytes (bar.rs:40-65): rotl1(v ^ (~rotl1(v) & rotl2(v) & rotl3(v)))
// On the ALU pipe only.  A byte-wise rotate-left by k is the 32-bit rotate-left by k with the k low bits of
// every byte replaced by those of the 32-bit rotate-right by 8-k (the bits that wrapped inside the byte): two funnel shifts
// and ONE three-input LOP3 (bit select under a constant mask).  The plain shift-and-mask form costs the multiplier pipe four
// instructions per word (ptxas lowers the left shifts to IMAD.SHL / IMAD.IADD), and the multiplier pipe is what bounds the
// hash; this form is 8 SHF + 6 LOP3 and leaves that pipe alone.
template <int K>
__device__ __forceinline__ uint32_t rotl_bytes(uint32_t v) {
    const uint32_t low = (0xffu >> (8 - K)) * 0x01010101u;
    const uint32_t a = __funnelshift_l(v, v, K), b = __funnelshift_r(v, v, 8 - K);
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(d) : "r"(a), "r"(b), "r"(low));  // (a & ~low) | (b & low)
    return d;
}
__device__ __forceinline__ uint32_t sbox4(uint32_t v) {
    const uint32_t t1 = rotl_bytes<1>(v), t2 = rotl_bytes<2>(v), t3 = rotl_bytes<3>(v);
    return rotl_bytes<1>((~t1 & t2 & t3) ^ v);
}

// q * p for q = 0..5 (skyscraper/core/src/constants.rs:9-16 MODULUS[0..5]) as 32-bit limbs
__constant__ uint32_t SKY_QP[6][8] = {
    {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u},
    {0xe0000002u, 0x87c3eb27u, 0xf372e122u, 0x5067d090u, 0x0302b0bau, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u},
    {0xd0000003u, 0xcba5e0bbu, 0x6d2c51b3u, 0x789bb8d9u, 0x84840917u, 0x28f0d123u, 0xa394e07du, 0x912ceb58u},
    {0xc0000004u, 0x0f87d64fu, 0xe6e5c245u, 0xa0cfa121u, 0x06056174u, 0xe14116dau, 0x84c680a6u, 0xc19139cbu},
    {0xb0000005u, 0x5369cbe3u, 0x609f32d6u, 0xc903896au, 0x8786b9d1u, 0x99915c90u, 0x65f820d0u, 0xf1f5883eu},
};
// value < 2^256 -> [0, p): q = floor(top limb / (P7+1)) in 0..5, subtract q*p, one conditional subtract
// (the reduce_partial idea of skyscraper/core/src/reduce.rs:33-40 followed by reduce_1 :21-29)
__device__ __forceinline__ fr sky_reduce(const fr& x) {
    uint32_t q = x.v[7] / (PK_P7 + 1u);
    uint32_t qp[8];
    // the multiple comes from constant memory (8 LDC) instead of 8 wide multiplies: the multiplier pipe bounds the hash
#pragma unroll
    for (int k = 0; k < 8; k++) qp[k] = SKY_QP[q][k];
    fr d;
    asm("sub.cc.u32 %0, %8, %16;\n\t"
        "subc.cc.u32 %1, %9, %17;\n\t"
        "subc.cc.u32 %2, %10, %18;\n\t"
        "subc.cc.u32 %3, %11, %19;\n\t"
        "subc.cc.u32 %4, %12, %20;\n\t"
        "subc.cc.u32 %5, %13, %21;\n\t"
        "subc.cc.u32 %6, %14, %22;\n\t"
        "subc.u32 %7, %15, %23;"
        : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]),
          "=r"(d.v[7])
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]),
          "r"(qp[0]), "r"(qp[1]), "r"(qp[2]), "r"(qp[3]), "r"(qp[4]), "r"(qp[5]), "r"(qp[6]), "r"(qp[7]));
    return fr_reduce_once(d);
}

// bar on a canonical value: swap the 16-byte halves, sbox every byte, reduce (reference.rs:80-94)
__device__ __forceinline__ fr sky_bar(const fr& x) {
    fr y;
    y.v[0] = sbox4(x.v[4]); y.v[1] = sbox4(x.v[5]); y.v[2] = sbox4(x.v[6]); y.v[3] = sbox4(x.v[7]);
    y.v[4] = sbox4(x.v[0]); y.v[5] = sbox4(x.v[1]); y.v[6] = sbox4(x.v[2]); y.v[7] = sbox4(x.v[3]);
    return sky_reduce(y);
}

__device__ __forceinline__ fr sky_rc(int i) {
    fr c;
#pragma unroll
    for (int k = 0; k < 8; k++) c.v[k] = SKY_RC[i][k];
    return c;
}

// raw 256-bit a + b + c (caller guarantees no overflow)
__device__ __forceinline__ fr add3_raw(const fr& a, const fr& b, const fr& c) {
    fr s;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]), "=r"(s.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(s.v[0]), "+r"(s.v[1]), "+r"(s.v[2]), "+r"(s.v[3]), "+r"(s.v[4]), "+r"(s.v[5]), "+r"(s.v[6]), "+r"(s.v[7])
        : "r"(c.v[0]), "r"(c.v[1]), "r"(c.v[2]), "r"(c.v[3]), "r"(c.v[4]), "r"(c.v[5]), "r"(c.v[6]), "r"(c.v[7]));
    return s;
}

// Lazy reduction (the reference's reduce_partial idea, skyscraper/core/src/reduce.rs:33-55, with 2p granularity):
// any s < 2^256 -> s - q*2p in [0, 2p + 3*2^224), q = floor(s7 / (top limb of 2p + 1)) in {0, 1, 2}.
__device__ __forceinline__ fr sky_reduce_2p(const fr& s) {
    const uint32_t D = 0x60c89ce6u;
    const bool q1 = s.v[7] >= D, q2 = s.v[7] >= 2u * D;
    // k = 0, 2p or 4p (skyscraper/core/src/constants.rs:9-16 MODULUS[2], MODULUS[4])
    uint32_t k0 = q2 ? 0xc0000004u : (q1 ? 0xe0000002u : 0u), k1 = q2 ? 0x0f87d64fu : (q1 ? 0x87c3eb27u : 0u);
    uint32_t k2 = q2 ? 0xe6e5c245u : (q1 ? 0xf372e122u : 0u), k3 = q2 ? 0xa0cfa121u : (q1 ? 0x5067d090u : 0u);
    uint32_t k4 = q2 ? 0x06056174u : (q1 ? 0x0302b0bau : 0u), k5 = q2 ? 0xe14116dau : (q1 ? 0x70a08b6du : 0u);
    uint32_t k6 = q2 ? 0x84c680a6u : (q1 ? 0xc2634053u : 0u), k7 = q2 ? 0xc19139cbu : (q1 ? 0x60c89ce5u : 0u);
    fr d;
    asm("sub.cc.u32 %0, %8, %16;\n\t"
        "subc.cc.u32 %1, %9, %17;\n\t"
        "subc.cc.u32 %2, %10, %18;\n\t"
        "subc.cc.u32 %3, %11, %19;\n\t"
        "subc.cc.u32 %4, %12, %20;\n\t"
        "subc.cc.u32 %5, %13, %21;\n\t"
        "subc.cc.u32 %6, %14, %22;\n\t"
        "subc.u32 %7, %15, %23;"
        : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7])
        : "r"(s.v[0]), "r"(s.v[1]), "r"(s.v[2]), "r"(s.v[3]), "r"(s.v[4]), "r"(s.v[5]), "r"(s.v[6]), "r"(s.v[7]),
          "r"(k0), "r"(k1), "r"(k2), "r"(k3), "r"(k4), "r"(k5), "r"(k6), "r"(k7));
    return d;
}
// [0, 2p + eps) -> canonical [0, p)
__device__ __forceinline__ fr sky_canon(const fr& x) { return fr_reduce_once(fr_reduce_once(x)); }

// l, r canonical (< p).  Returns compress(l, r) canonical.
// Two Feistel rounds per iteration so that (l, r) swap roles without register moves:
//   round 2j  : r <- r + F(l) + rc[2j]      (the new left half now lives in r)
//   round 2j+1: l <- l + F(r) + rc[2j+1]    (roles restored)
// Both rounds of a pair use the same F (bar for pairs 3 and 5 = rounds 6,7,10,11; reference.rs:49-60).
// The state is kept lazily reduced in [0, 2p + eps) like the reference does (generic.rs:81-101); only bar
// canonicalises its input (bar.rs:17) and only the output is fully reduced.  Bounds: fr_sqr_lazy(x) < 1.84p for
// x < 2.1p, so r + F + rc < 4.9p < 2^256.
// `keep_going(j)` is polled before every pair of rounds; when it returns false the compression is abandoned and `done` is
// cleared (the PoW scan stops hashing nonces that can no longer win).
template <class KeepGoing>
__device__ __forceinline__ fr sky_compress_while(const fr& l_in, const fr& r_in, KeepGoing keep_going, bool& done) {
    fr l = l_in, r = r_in;
    done = true;
#pragma unroll 1
    for (int j = 0; j < 9; j++) {
        if (!keep_going(j)) {
            done = false;
            return l;
        }
        const bool is_bar = (j == 3) | (j == 5);
        if (is_bar) {
            r = sky_reduce_2p(add3_raw(r, sky_bar(sky_canon(l)), sky_rc(2 * j)));
            l = sky_reduce_2p(add3_raw(l, sky_bar(sky_canon(r)), sky_rc(2 * j + 1)));
        } else {
            r = sky_reduce_2p(add3_raw(r, fr_sqr_lazy(l), sky_rc(2 * j)));
            l = sky_reduce_2p(add3_raw(l, fr_sqr_lazy(r), sky_rc(2 * j + 1)));
        }
    }
    return sky_reduce(add3_raw(l, l_in, fr_zero()));
}
__device__ __forceinline__ fr sky_compress(const fr& l_in, const fr& r_in) {
    bool done;
    return sky_compress_while(l_in, r_in, [](int) { return true; }, done);
}

// provekit/common/src/skyscraper/whir.rs:20-25 on Montgomery-form field elements
__device__ __forceinline__ fr sky_compress_mont(const fr& l, const fr& r) {
    return fr_to_mont(sky_compress(fr_from_mont(l), fr_from_mont(r)));
}

}  // namespace pk
