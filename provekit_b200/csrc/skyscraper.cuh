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
#ifdef PK_BAR_SLOW
__device__ __forceinline__ uint32_t sbox4(uint32_t v) {
    const uint32_t t1 = rotl_bytes<1>(v), t2 = rotl_bytes<2>(v), t3 = rotl_bytes<3>(v);
    return rotl_bytes<1>((~t1 & t2 & t3) ^ v);
}
#else
// rotl2(v) & rotl3(v) = rotl2(v & rotl1(v)): three byte-wise rotations instead of four, 6 SHF + 5 LOP3 per word
__device__ __forceinline__ uint32_t sbox4(uint32_t v) {
    const uint32_t t1 = rotl_bytes<1>(v);
    const uint32_t t23 = rotl_bytes<2>(v & t1);
    return rotl_bytes<1>((~t1 & t23) ^ v);
}
#endif

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
// LAZY: stop after the subtraction of q p, i.e. return a representative in [0, p + 6 * 2^224) — enough for a bar output that
// goes straight into a table-driven round sum (which only needs r + F < 5 p; see sky_round_sum)
template <bool LAZY = false>
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
    return LAZY ? d : fr_reduce_once(d);
}

// bar on a canonical value: swap the 16-byte halves, sbox every byte, reduce (reference.rs:80-94)
template <bool LAZY = false>
__device__ __forceinline__ fr sky_bar(const fr& x) {
    fr y;
    y.v[0] = sbox4(x.v[4]); y.v[1] = sbox4(x.v[5]); y.v[2] = sbox4(x.v[6]); y.v[3] = sbox4(x.v[7]);
    y.v[4] = sbox4(x.v[0]); y.v[5] = sbox4(x.v[1]); y.v[6] = sbox4(x.v[2]); y.v[7] = sbox4(x.v[3]);
    return sky_reduce<LAZY>(y);
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

// rc[i] - q p (mod 2^256), i = 0..17, q = 0..4: one Feistel round's constant and the reduction of r + F(l) folded into ONE
// table look-up.  s0 = r + F < 2.11 p + 1.84 p, q = floor(top limb of s0 / (top limb of p + 1)) <= floor(s0 / p), so
// s0 - q p lies in [0, p + 5 * 2^224) and s0 + T[i][q] = s0 - q p + rc[i] in [0, 2 p + 5 * 2^224) < 2.11 p: the lazy range the
// squaring accepts (skyscraper/core/src/generic.rs:81-101 keeps the same kind of representative; reduce.rs:33-55).
// 16 add-with-carry + the constant loads per round instead of 24 add/sub-with-carry + 16 selects.  Two copies: constant memory
// (8 LDC per round; lanes with different q serialise) and global memory (2 LDG.128 per round through L1).
#define PK_RCQ_ROWS \
    {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u}, \
    {0x0fffffffu, 0xbc1e0a6cu, 0x86468f6eu, 0xd7cc17b7u, 0x7e7ea7a2u, 0x47afba49u, 0x1ece5fd6u, 0xcf9bb18du}, \
    {0x1ffffffeu, 0x783c14d8u, 0x0c8d1eddu, 0xaf982f6fu, 0xfcfd4f45u, 0x8f5f7492u, 0x3d9cbfacu, 0x9f37631au}, \
    {0x2ffffffdu, 0x345a1f44u, 0x92d3ae4cu, 0x87644726u, 0x7b7bf6e8u, 0xd70f2edcu, 0x5c6b1f82u, 0x6ed314a7u}, \
    {0x3ffffffcu, 0xf07829b0u, 0x191a3dbau, 0x5f305edeu, 0xf9fa9e8bu, 0x1ebee925u, 0x7b397f59u, 0x3e6ec634u}, \
    {0x270bd744u, 0x903c4324u, 0x08a7d269u, 0x873125f7u, 0x06c83855u, 0x081dd279u, 0xea6d7667u, 0x276b1823u}, \
    {0x370bd743u, 0x4c5a4d90u, 0x8eee61d8u, 0x5efd3daeu, 0x8546dff8u, 0x4fcd8cc2u, 0x093bd63du, 0xf706c9b1u}, \
    {0x470bd742u, 0x087857fcu, 0x1534f147u, 0x36c95566u, 0x03c5879bu, 0x977d470cu, 0x280a3613u, 0xc6a27b3eu}, \
    {0x570bd741u, 0xc4966268u, 0x9b7b80b5u, 0x0e956d1du, 0x82442f3eu, 0xdf2d0155u, 0x46d895e9u, 0x963e2ccbu}, \
    {0x670bd740u, 0x80b46cd4u, 0x21c21024u, 0xe66184d5u, 0x00c2d6e0u, 0x26dcbb9fu, 0x65a6f5c0u, 0x65d9de58u}, \
    {0x4b378d71u, 0x7ac8edbbu, 0xd99e2cb7u, 0xe29d79f3u, 0x4c1a5a18u, 0x75141791u, 0x58a484a6u, 0x0cf02bd7u}, \
    {0x5b378d70u, 0x36e6f827u, 0x5fe4bc26u, 0xba6991abu, 0xca9901bbu, 0xbcc3d1dau, 0x7772e47cu, 0xdc8bdd64u}, \
    {0x6b378d6fu, 0xf3050293u, 0xe62b4b94u, 0x9235a962u, 0x4917a95eu, 0x04738c24u, 0x96414453u, 0xac278ef1u}, \
    {0x7b378d6eu, 0xaf230cffu, 0x6c71db03u, 0x6a01c11au, 0xc7965101u, 0x4c23466du, 0xb50fa429u, 0x7bc3407eu}, \
    {0x8b378d6du, 0x6b41176bu, 0xf2b86a72u, 0x41cdd8d1u, 0x4614f8a4u, 0x93d300b7u, 0xd3de03ffu, 0x4b5ef20bu}, \
    {0x69e5bc36u, 0xfa7adc67u, 0x7cca387du, 0x1c3f8e29u, 0x63481db0u, 0x0eb7730du, 0x18ede544u, 0x25b0e03fu}, \
    {0x79e5bc35u, 0xb698e6d3u, 0x0310c7ecu, 0xf40ba5e1u, 0xe1c6c552u, 0x56672d56u, 0x37bc451au, 0xf54c91ccu}, \
    {0x89e5bc34u, 0x72b6f13fu, 0x8957575bu, 0xcbd7bd98u, 0x60456cf5u, 0x9e16e7a0u, 0x568aa4f0u, 0xc4e84359u}, \
    {0x99e5bc33u, 0x2ed4fbabu, 0x0f9de6cau, 0xa3a3d550u, 0xdec41498u, 0xe5c6a1e9u, 0x755904c6u, 0x9483f4e6u}, \
    {0xa9e5bc32u, 0xeaf30617u, 0x95e47638u, 0x7b6fed07u, 0x5d42bc3bu, 0x2d765c33u, 0x9427649du, 0x641fa673u}, \
    {0x2f03cfb7u, 0x57847e65u, 0x68873404u, 0x33440b96u, 0x49af80bcu, 0x955a32e8u, 0xbe14ae70u, 0x002882fcu}, \
    {0x3f03cfb6u, 0x13a288d1u, 0xeecdc373u, 0x0b10234du, 0xc82e285fu, 0xdd09ed31u, 0xdce30e46u, 0xcfc43489u}, \
    {0x4f03cfb5u, 0xcfc0933du, 0x751452e1u, 0xe2dc3b05u, 0x46acd001u, 0x24b9a77bu, 0xfbb16e1du, 0x9f5fe616u}, \
    {0x5f03cfb4u, 0x8bde9da9u, 0xfb5ae250u, 0xbaa852bcu, 0xc52b77a4u, 0x6c6961c4u, 0x1a7fcdf3u, 0x6efb97a4u}, \
    {0x6f03cfb3u, 0x47fca815u, 0x81a171bfu, 0x92746a74u, 0x43aa1f47u, 0xb4191c0eu, 0x394e2dc9u, 0x3e974931u}, \
    {0x6257d4d7u, 0x97923139u, 0x1b37d3c1u, 0x29989c3eu, 0x7f1277bau, 0x12ef02b4u, 0x1e2b7a9cu, 0x039ad857u}, \
    {0x7257d4d6u, 0x53b03ba5u, 0xa17e6330u, 0x0164b3f5u, 0xfd911f5du, 0x5a9ebcfdu, 0x3cf9da72u, 0xd33689e4u}, \
    {0x8257d4d5u, 0x0fce4611u, 0x27c4f29fu, 0xd930cbadu, 0x7c0fc6ffu, 0xa24e7747u, 0x5bc83a48u, 0xa2d23b71u}, \
    {0x9257d4d4u, 0xcbec507du, 0xae0b820du, 0xb0fce364u, 0xfa8e6ea2u, 0xe9fe3190u, 0x7a969a1eu, 0x726decfeu}, \
    {0xa257d4d3u, 0x880a5ae9u, 0x3452117cu, 0x88c8fb1cu, 0x790d1645u, 0x31adebdau, 0x9964f9f5u, 0x42099e8bu}, \
    {0xabbb7887u, 0xb5b48465u, 0xe6ba2d2bu, 0xa72a6bc5u, 0x712f7b29u, 0x4cd48043u, 0x0fc1fc1au, 0x1142d541u}, \
    {0xbbbb7886u, 0x71d28ed1u, 0x6d00bc9au, 0x7ef6837du, 0xefae22ccu, 0x94843a8cu, 0x2e905bf0u, 0xe0de86ceu}, \
    {0xcbbb7885u, 0x2df0993du, 0xf3474c09u, 0x56c29b34u, 0x6e2cca6fu, 0xdc33f4d6u, 0x4d5ebbc6u, 0xb07a385bu}, \
    {0xdbbb7884u, 0xea0ea3a9u, 0x798ddb77u, 0x2e8eb2ecu, 0xecab7212u, 0x23e3af1fu, 0x6c2d1b9du, 0x8015e9e8u}, \
    {0xebbb7883u, 0xa62cae15u, 0xffd46ae6u, 0x065acaa3u, 0x6b2a19b5u, 0x6b936969u, 0x8afb7b73u, 0x4fb19b75u}, \
    {0x059075d3u, 0x7ab2c156u, 0x047999b2u, 0x17cb3594u, 0x98f289f7u, 0x44f2c935u, 0x69bc0becu, 0x1d78439fu}, \
    {0x159075d2u, 0x36d0cbc2u, 0x8ac02921u, 0xef974d4bu, 0x17713199u, 0x8ca2837fu, 0x888a6bc2u, 0xed13f52cu}, \
    {0x259075d1u, 0xf2eed62eu, 0x1106b88fu, 0xc7636503u, 0x95efd93cu, 0xd4523dc8u, 0xa758cb98u, 0xbcafa6b9u}, \
    {0x359075d0u, 0xaf0ce09au, 0x974d47feu, 0x9f2f7cbau, 0x146e80dfu, 0x1c01f812u, 0xc6272b6fu, 0x8c4b5846u}, \
    {0x459075cfu, 0x6b2aeb06u, 0x1d93d76du, 0x76fb9472u, 0x92ed2882u, 0x63b1b25bu, 0xe4f58b45u, 0x5be709d3u}, \
    {0x138b8edbu, 0x05d7a965u, 0xd55c48b1u, 0x36ef35a3u, 0xac6f1628u, 0x8ddfb8a1u, 0x08f4ff82u, 0x258588a5u}, \
    {0x238b8edau, 0xc1f5b3d1u, 0x5ba2d81fu, 0x0ebb4d5bu, 0x2aedbdcbu, 0xd58f72ebu, 0x27c35f58u, 0xf5213a32u}, \
    {0x338b8ed9u, 0x7e13be3du, 0xe1e9678eu, 0xe6876512u, 0xa96c656du, 0x1d3f2d34u, 0x4691bf2fu, 0xc4bcebbfu}, \
    {0x438b8ed8u, 0x3a31c8a9u, 0x682ff6fdu, 0xbe537ccau, 0x27eb0d10u, 0x64eee77eu, 0x65601f05u, 0x94589d4cu}, \
    {0x538b8ed7u, 0xf64fd315u, 0xee76866bu, 0x961f9481u, 0xa669b4b3u, 0xac9ea1c7u, 0x842e7edbu, 0x63f44ed9u}, \
    {0xfccb49e9u, 0x1596fb9au, 0x9a09a95bu, 0x9a7367d6u, 0x84e4c157u, 0x9bc43f69u, 0xd2f514feu, 0x13087879u}, \
    {0x0ccb49e8u, 0xd1b50607u, 0x205038c9u, 0x723f7f8eu, 0x036368fau, 0xe373f9b3u, 0xf1c374d4u, 0xe2a42a06u}, \
    {0x1ccb49e7u, 0x8dd31073u, 0xa696c838u, 0x4a0b9745u, 0x81e2109du, 0x2b23b3fcu, 0x1091d4abu, 0xb23fdb94u}, \
    {0x2ccb49e6u, 0x49f11adfu, 0x2cdd57a7u, 0x21d7aefdu, 0x0060b840u, 0x72d36e46u, 0x2f603481u, 0x81db8d21u}, \
    {0x3ccb49e5u, 0x060f254bu, 0xb323e716u, 0xf9a3c6b4u, 0x7edf5fe2u, 0xba83288fu, 0x4e2e9457u, 0x51773eaeu}, \
    {0x3b4109fau, 0x295ccd23u, 0xed868012u, 0xe1d72f89u, 0x4bc88a8eu, 0x2e9e1eeau, 0x98c45232u, 0x17dadee8u}, \
    {0x4b4109f9u, 0xe57ad78fu, 0x73cd0f80u, 0xb9a34741u, 0xca473231u, 0x764dd933u, 0xb792b208u, 0xe7769075u}, \
    {0x5b4109f8u, 0xa198e1fbu, 0xfa139eefu, 0x916f5ef8u, 0x48c5d9d4u, 0xbdfd937du, 0xd66111deu, 0xb7124202u}, \
    {0x6b4109f7u, 0x5db6ec67u, 0x805a2e5eu, 0x693b76b0u, 0xc7448177u, 0x05ad4dc6u, 0xf52f71b5u, 0x86adf38fu}, \
    {0x7b4109f6u, 0x19d4f6d3u, 0x06a0bdcdu, 0x41078e68u, 0x45c3291au, 0x4d5d0810u, 0x13fdd18bu, 0x5649a51du}, \
    {0xaa1f486fu, 0x9a8590b4u, 0x30e9130eu, 0xb75834b4u, 0x34d5de31u, 0xb8e90b10u, 0x46e7f4a6u, 0x295c6d15u}, \
    {0xba1f486eu, 0x56a39b20u, 0xb72fa27du, 0x8f244c6bu, 0xb35485d4u, 0x0098c559u, 0x65b6547du, 0xf8f81ea2u}, \
    {0xca1f486du, 0x12c1a58cu, 0x3d7631ecu, 0x66f06423u, 0x31d32d77u, 0x48487fa3u, 0x8484b453u, 0xc893d02fu}, \
    {0xda1f486cu, 0xcedfaff8u, 0xc3bcc15au, 0x3ebc7bdau, 0xb051d51au, 0x8ff839ecu, 0xa3531429u, 0x982f81bcu}, \
    {0xea1f486bu, 0x8afdba64u, 0x4a0350c9u, 0x16889392u, 0x2ed07cbdu, 0xd7a7f436u, 0xc22173ffu, 0x67cb3349u}, \
    {0x4c6eb892u, 0x850adcb7u, 0x05b92fc3u, 0x07699ef3u, 0xa1720f2du, 0x4ef96a2bu, 0x1d3ed446u, 0x1288ca0eu}, \
    {0x5c6eb891u, 0x4128e723u, 0x8bffbf32u, 0xdf35b6aau, 0x1ff0b6cfu, 0x96a92475u, 0x3c0d341cu, 0xe2247b9bu}, \
    {0x6c6eb890u, 0xfd46f18fu, 0x12464ea0u, 0xb701ce62u, 0x9e6f5e72u, 0xde58debeu, 0x5adb93f2u, 0xb1c02d28u}, \
    {0x7c6eb88fu, 0xb964fbfbu, 0x988cde0fu, 0x8ecde619u, 0x1cee0615u, 0x26089908u, 0x79a9f3c9u, 0x815bdeb5u}, \
    {0x8c6eb88eu, 0x75830667u, 0x1ed36d7eu, 0x6699fdd1u, 0x9b6cadb8u, 0x6db85351u, 0x9878539fu, 0x50f79042u}, \
    {0x49d1b5eeu, 0x01960f93u, 0x69371c69u, 0x8ccad307u, 0x91c98662u, 0xe5c81e89u, 0x1ae023f3u, 0x17563b4du}, \
    {0x59d1b5edu, 0xbdb419ffu, 0xef7dabd7u, 0x6496eabeu, 0x10482e05u, 0x2d77d8d3u, 0x39ae83cau, 0xe6f1ecdau}, \
    {0x69d1b5ecu, 0x79d2246bu, 0x75c43b46u, 0x3c630276u, 0x8ec6d5a8u, 0x7527931cu, 0x587ce3a0u, 0xb68d9e67u}, \
    {0x79d1b5ebu, 0x35f02ed7u, 0xfc0acab5u, 0x142f1a2du, 0x0d457d4bu, 0xbcd74d66u, 0x774b4376u, 0x86294ff4u}, \
    {0x89d1b5eau, 0xf20e3943u, 0x82515a23u, 0xebfb31e5u, 0x8bc424edu, 0x048707afu, 0x9619a34du, 0x55c50181u}, \
    {0x76b32917u, 0x6ba01e94u, 0xdd977bc9u, 0xa1cb0a3au, 0x5815f030u, 0x86815a94u, 0xe91a1eeau, 0x2869043bu}, \
    {0x86b32916u, 0x27be2900u, 0x63de0b38u, 0x799721f2u, 0xd69497d3u, 0xce3114ddu, 0x07e87ec0u, 0xf804b5c9u}, \
    {0x96b32915u, 0xe3dc336cu, 0xea249aa6u, 0x516339a9u, 0x55133f76u, 0x15e0cf27u, 0x26b6de97u, 0xc7a06756u}, \
    {0xa6b32914u, 0x9ffa3dd8u, 0x706b2a15u, 0x292f5161u, 0xd391e719u, 0x5d908970u, 0x45853e6du, 0x973c18e3u}, \
    {0xb6b32913u, 0x5c184844u, 0xf6b1b984u, 0x00fb6918u, 0x52108ebcu, 0xa54043bau, 0x64539e43u, 0x66d7ca70u}, \
    {0x5511d976u, 0x81776c88u, 0x47f414e7u, 0x7475d34fu, 0x095d96cfu, 0x5d090056u, 0xff59e79au, 0x14941f0au}, \
    {0x6511d975u, 0x3d9576f4u, 0xce3aa456u, 0x4c41eb06u, 0x87dc3e72u, 0xa4b8ba9fu, 0x1e284770u, 0xe42fd098u}, \
    {0x7511d974u, 0xf9b38160u, 0x548133c4u, 0x240e02beu, 0x065ae615u, 0xec6874e9u, 0x3cf6a746u, 0xb3cb8225u}, \
    {0x8511d973u, 0xb5d18bccu, 0xdac7c333u, 0xfbda1a75u, 0x84d98db7u, 0x34182f32u, 0x5bc5071du, 0x836733b2u}, \
    {0x9511d972u, 0x71ef9638u, 0x610e52a2u, 0xd3a6322du, 0x0358355au, 0x7bc7e97cu, 0x7a9366f3u, 0x5302e53fu}, \
    {0x8fc8c034u, 0xbc40b4fdu, 0xcce4fd48u, 0xbb7142c3u, 0x8a39005au, 0x31835675u, 0x90f4379fu, 0x1ce337a1u}, \
    {0x9fc8c033u, 0x785ebf69u, 0x532b8cb7u, 0x933d5a7bu, 0x08b7a7fdu, 0x793310bfu, 0xafc29775u, 0xec7ee92eu}, \
    {0xafc8c032u, 0x347cc9d5u, 0xd9721c26u, 0x6b097232u, 0x87364fa0u, 0xc0e2cb08u, 0xce90f74bu, 0xbc1a9abbu}, \
    {0xbfc8c031u, 0xf09ad441u, 0x5fb8ab94u, 0x42d589eau, 0x05b4f743u, 0x08928552u, 0xed5f5722u, 0x8bb64c48u}, \
    {0xcfc8c030u, 0xacb8deadu, 0xe5ff3b03u, 0x1aa1a1a1u, 0x84339ee6u, 0x50423f9bu, 0x0c2db6f8u, 0x5b51fdd6u}, \
    {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u}, \
    {0x0fffffffu, 0xbc1e0a6cu, 0x86468f6eu, 0xd7cc17b7u, 0x7e7ea7a2u, 0x47afba49u, 0x1ece5fd6u, 0xcf9bb18du}, \
    {0x1ffffffeu, 0x783c14d8u, 0x0c8d1eddu, 0xaf982f6fu, 0xfcfd4f45u, 0x8f5f7492u, 0x3d9cbfacu, 0x9f37631au}, \
    {0x2ffffffdu, 0x345a1f44u, 0x92d3ae4cu, 0x87644726u, 0x7b7bf6e8u, 0xd70f2edcu, 0x5c6b1f82u, 0x6ed314a7u}, \
    {0x3ffffffcu, 0xf07829b0u, 0x191a3dbau, 0x5f305edeu, 0xf9fa9e8bu, 0x1ebee925u, 0x7b397f59u, 0x3e6ec634u}, \

__constant__ uint32_t SKY_RCQ[18 * 5][8] = {PK_RCQ_ROWS};
__device__ const uint4 SKY_RCQ_G[18 * 5][2] = {PK_RCQ_ROWS};
#undef PK_RCQ_ROWS
// How a round adds its constant and reduces: 0 = add rc, subtract a selected multiple of 2p (registers only: the form for
// latency-bound one-thread / one-block kernels), 1 = constant-memory table, 2 = global-memory table (throughput kernels).
#ifndef PK_ROUND_MODE
#define PK_ROUND_MODE 2
#endif
template <int MODE>
__device__ __forceinline__ fr sky_round_sum(const fr& r, const fr& F, int i) {
    if (MODE == 0) return sky_reduce_2p(add3_raw(r, F, sky_rc(i)));
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
        : "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]),
          "r"(F.v[0]), "r"(F.v[1]), "r"(F.v[2]), "r"(F.v[3]), "r"(F.v[4]), "r"(F.v[5]), "r"(F.v[6]), "r"(F.v[7]));
    const uint32_t q = s.v[7] / (PK_P7 + 1u);
    uint32_t t[8];
    if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = SKY_RCQ[5 * i + q][k];
    } else {
        const uint4 a = __ldg(&SKY_RCQ_G[5 * i + q][0]), b = __ldg(&SKY_RCQ_G[5 * i + q][1]);
        t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w;
        t[4] = b.x; t[5] = b.y; t[6] = b.z; t[7] = b.w;
    }
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(s.v[0]), "+r"(s.v[1]), "+r"(s.v[2]), "+r"(s.v[3]), "+r"(s.v[4]), "+r"(s.v[5]), "+r"(s.v[6]), "+r"(s.v[7])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]));
    return s;
}

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
template <int MODE = PK_ROUND_MODE, class KeepGoing>
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
#ifdef PK_BAR_SLOW
            constexpr bool LAZY_BAR = false;
#else
            constexpr bool LAZY_BAR = MODE != 0;  // the table-driven sum absorbs a bar output below p + 6 * 2^224
#endif
            r = sky_round_sum<MODE>(r, sky_bar<LAZY_BAR>(sky_canon(l)), 2 * j);
            l = sky_round_sum<MODE>(l, sky_bar<LAZY_BAR>(sky_canon(r)), 2 * j + 1);
        } else {
            r = sky_round_sum<MODE>(r, fr_sqr_lazy(l), 2 * j);
            l = sky_round_sum<MODE>(l, fr_sqr_lazy(r), 2 * j + 1);
        }
    }
    return sky_reduce(add3_raw(l, l_in, fr_zero()));
}
template <int MODE = PK_ROUND_MODE>
__device__ __forceinline__ fr sky_compress(const fr& l_in, const fr& r_in) {
    bool done;
    return sky_compress_while<MODE>(l_in, r_in, [](int) { return true; }, done);
}

// provekit/common/src/skyscraper/whir.rs:20-25 on Montgomery-form field elements
__device__ __forceinline__ fr sky_compress_mont(const fr& l, const fr& r) {
    return fr_to_mont(sky_compress<0>(fr_from_mont(l), fr_from_mont(r)));
}

}  // namespace pk
