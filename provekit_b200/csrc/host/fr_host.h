// provekit_b200/csrc/host/fr_host.h — host-side BN254-Fr arithmetic for the product's own host code
// (Fiat-Shamir sponge, challenge bookkeeping, twiddle seeds).  The Rust host of the reference does the
// same work with ark-ff [EXT]; this is the C++ stand-in the C-ABI harness needs (no Rust toolchain in
// the build image).  Independent of oracle/ by design: the product never links the oracle.
#pragma once
#include <cstdint>
#include <cstring>

namespace pkh {

typedef unsigned __int128 u128;

struct Fr {
    uint64_t l[4];
    bool operator==(const Fr& o) const { return std::memcmp(l, o.l, 32) == 0; }
    bool operator!=(const Fr& o) const { return !(*this == o); }
};

// skyscraper/block-multiplier/src/constants.rs:1-39
static const uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL,
                              0x30644e72e131a029ULL};
static const Fr ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const Fr R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const Fr ZERO = {{0, 0, 0, 0}};
static const uint64_t NP0 = 0xc2e1f593efffffffULL;

inline bool geq_p(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] != P[i]) return a[i] > P[i];
    }
    return true;
}
inline void sub_p(uint64_t a[4]) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - P[i] - b;
        a[i] = (uint64_t)d;
        b = (d >> 64) & 1;
    }
}
inline Fr add(const Fr& a, const Fr& b) {
    Fr r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_p(r.l)) sub_p(r.l);
    return r;
}
inline Fr sub(const Fr& a, const Fr& b) {
    Fr r;
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.l[i] - b.l[i] - bw;
        r.l[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    if (bw) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)r.l[i] + P[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
// Montgomery product a*b*2^-256 mod p (operands < p)
inline Fr mul(const Fr& a, const Fr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * NP0;
        c = ((u128)m * P[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_p(r.l)) sub_p(r.l);
    return r;
}
inline Fr sqr(const Fr& a) { return mul(a, a); }
inline Fr dbl(const Fr& a) { return add(a, a); }
inline Fr neg(const Fr& a) { return sub(ZERO, a); }
// any 256-bit integer -> Montgomery form of (value mod p): Fp::new(BigInt)
inline Fr from_canonical(const uint64_t c[4]) {
    Fr x = {{c[0], c[1], c[2], c[3]}};
    while (geq_p(x.l)) sub_p(x.l);
    return mul(x, R2);
}
inline void to_canonical(const Fr& a, uint64_t c[4]) {
    Fr one = {{1, 0, 0, 0}};
    Fr r = mul(a, one);
    std::memcpy(c, r.l, 32);
}
inline Fr from_u64(uint64_t v) {
    uint64_t c[4] = {v, 0, 0, 0};
    return from_canonical(c);
}
inline Fr pow_u64(Fr b, uint64_t e) {
    Fr r = ONE;
    while (e) {
        if (e & 1) r = mul(r, b);
        b = sqr(b);
        e >>= 1;
    }
    return r;
}
inline Fr inv(const Fr& a) {
    uint64_t e[4] = {P[0] - 2, P[1], P[2], P[3]};
    Fr r = ONE;
    for (int i = 255; i >= 0; i--) {
        r = sqr(r);
        if ((e[i / 64] >> (i % 64)) & 1) r = mul(r, a);
    }
    return r;
}
// arkworks BN254 Fr TWO_ADIC_ROOT_OF_UNITY (order 2^28), canonical [EXT ark-bn254]
static const uint64_t ROOT28[4] = {0x9bd61b6e725b19f0ULL, 0x402d111e41112ed4ULL, 0x00e0a7eb8ef62abcULL,
                                   0x2a3c09f0a58a7e85ULL};
// generator of the Radix2EvaluationDomain of size 2^log_n
inline Fr root_of_unity(int log_n) {
    Fr g = from_canonical(ROOT28);
    for (int i = log_n; i < 28; i++) g = sqr(g);
    return g;
}
// 1/2, provekit/common/src/utils/mod.rs:23-25
inline Fr half() { return inv(from_u64(2)); }

}  // namespace pkh
