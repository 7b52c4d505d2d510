/* oracle/fr.h — BN254 scalar field on 4x64-bit limbs (CPU).
 *
 * TEST INFRASTRUCTURE ONLY (CPU oracle).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from this directory.
 *
 * Restates the arithmetic the reference gets from ark-ff 0.5 [EXT] (`Fp<MontBackend<BN254Config,4>,4>`,
 * provekit/prover/src/whir_r1cs.rs:136-139): elements are 4 little-endian u64 limbs in Montgomery
 * form, R = 2^256.  Constants: skyscraper/block-multiplier/src/constants.rs:1-39.
 * Any correct Montgomery implementation yields the same canonical values, so a
 * plain CIOS multiplier on unsigned __int128 is used.
 */
#ifndef PK_ORACLE_FR_H
#define PK_ORACLE_FR_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fr_t;

/* constants.rs:3-8 */
static const uint64_t FR_P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL,
                                 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
/* constants.rs:18-23  R mod p  (Montgomery form of 1) */
static const fr_t FR_ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL,
                             0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
/* constants.rs:26-31  R^2 mod p */
static const fr_t FR_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL,
                            0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const fr_t FR_ZERO = {{0, 0, 0, 0}};
/* constants.rs:1 */
#define FR_NP0 0xc2e1f593efffffffULL

static inline int fr_raw_geq_p(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > FR_P[i]) return 1;
        if (a[i] < FR_P[i]) return 0;
    }
    return 1;
}
static inline void fr_raw_sub_p(uint64_t a[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - FR_P[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}
static inline int fr_eq(const fr_t *a, const fr_t *b) { return memcmp(a, b, sizeof(fr_t)) == 0; }
static inline int fr_is_zero(const fr_t *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }

static inline fr_t fr_add(fr_t a, fr_t b) {
    fr_t r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    /* p < 2^254 so a+b < 2^255: no carry out */
    if (fr_raw_geq_p(r.l)) fr_raw_sub_p(r.l);
    return r;
}
static inline fr_t fr_sub(fr_t a, fr_t b) {
    fr_t r;
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.l[i] - b.l[i] - borrow;
        r.l[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)r.l[i] + FR_P[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
static inline fr_t fr_neg(fr_t a) { return fr_sub(FR_ZERO, a); }
static inline fr_t fr_dbl(fr_t a) { return fr_add(a, a); }

/* Montgomery product a*b*2^-256 mod p on raw limbs (inputs < p, output < p).
 * Same function as block_multiplier::scalar_mul (skyscraper/block-multiplier/src/scalar.rs:72-132)
 * up to the representative: the reference returns a lazily reduced value, we return the canonical one. */
static inline fr_t fr_mul(fr_t a, fr_t b) {
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
        uint64_t m = t[0] * FR_NP0;
        c = (u128)m * FR_P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * FR_P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fr_t r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fr_raw_geq_p(r.l)) fr_raw_sub_p(r.l);
    return r;
}
static inline fr_t fr_sqr(fr_t a) { return fr_mul(a, a); }

/* canonical integer (any value < 2^256 is reduced mod p) -> Montgomery form; Fp::new(BigInt) */
static inline fr_t fr_from_canonical(const uint64_t c[4]) {
    fr_t x = {{c[0], c[1], c[2], c[3]}};
    while (fr_raw_geq_p(x.l)) fr_raw_sub_p(x.l);
    return fr_mul(x, FR_R2);
}
/* Montgomery form -> canonical integer; Fp::into_bigint */
static inline void fr_to_canonical(fr_t a, uint64_t c[4]) {
    fr_t one = {{1, 0, 0, 0}};
    fr_t r = fr_mul(a, one);
    memcpy(c, r.l, 32);
}
static inline fr_t fr_from_u64(uint64_t v) {
    uint64_t c[4] = {v, 0, 0, 0};
    return fr_from_canonical(c);
}
static inline fr_t fr_pow_u64(fr_t b, uint64_t e) {
    fr_t r = FR_ONE;
    while (e) {
        if (e & 1) r = fr_mul(r, b);
        b = fr_sqr(b);
        e >>= 1;
    }
    return r;
}
/* a^(p-2) */
static inline fr_t fr_inv(fr_t a) {
    uint64_t e[4] = {FR_P[0] - 2, FR_P[1], FR_P[2], FR_P[3]};
    fr_t r = FR_ONE;
    for (int i = 255; i >= 0; i--) {
        r = fr_sqr(r);
        if ((e[i / 64] >> (i % 64)) & 1) r = fr_mul(r, a);
    }
    return r;
}
/* arkworks BN254 Fr TWO_ADIC_ROOT_OF_UNITY (2^28-th root), canonical limbs [EXT ark-bn254];
 * pinned by the domain generator stored in the reference fixture. */
static const uint64_t FR_ROOT28_CANON[4] = {0x9bd61b6e725b19f0ULL, 0x402d111e41112ed4ULL,
                                            0x00e0a7eb8ef62abcULL, 0x2a3c09f0a58a7e85ULL};
static inline fr_t fr_root_of_unity(int log_n) {
    fr_t g = fr_from_canonical(FR_ROOT28_CANON);
    for (int i = log_n; i < 28; i++) g = fr_sqr(g);
    return g;
}
#endif
