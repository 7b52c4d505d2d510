/* oracle/skyscraper.c — Skyscraper v2 (and the stale v1) over BN254-Fr, t = 1 (CPU oracle).
 * TEST INFRASTRUCTURE ONLY — see oracle/fr.h.
 *
 * Follows skyscraper/core/src/reference.rs:41-98 (spec) with the structure of
 * generic.rs:77-102; squaring = block_multiplier::scalar_sqr semantics x^2 * 2^-256 mod p
 * (skyscraper/block-multiplier/src/scalar.rs:11-70); bar = bar.rs:15-42; constants.rs:32-51.
 * Intermediate representatives are kept canonical here (the reference keeps them lazily reduced);
 * only `bar` sees the representative and it canonicalises first (bar.rs:17), and the output is
 * fully reduced (generic.rs:101), so the results coincide bit for bit.
 */
#include "pk_oracle.h"
#include "fr.h"

static const uint64_t SKY_RC[18][4] = {
    {0x0000000000000000ULL, 0x0000000000000000ULL, 0x0000000000000000ULL, 0x0000000000000000ULL},
    {0x903c4324270bd744ULL, 0x873125f708a7d269ULL, 0x081dd27906c83855ULL, 0x276b1823ea6d7667ULL},
    {0x7ac8edbb4b378d71ULL, 0xe29d79f3d99e2cb7ULL, 0x751417914c1a5a18ULL, 0x0cf02bd758a484a6ULL},
    {0xfa7adc6769e5bc36ULL, 0x1c3f8e297cca387dULL, 0x0eb7730d63481db0ULL, 0x25b0e03f18ede544ULL},
    {0x57847e652f03cfb7ULL, 0x33440b9668873404ULL, 0x955a32e849af80bcULL, 0x002882fcbe14ae70ULL},
    {0x979231396257d4d7ULL, 0x29989c3e1b37d3c1ULL, 0x12ef02b47f1277baULL, 0x039ad8571e2b7a9cULL},
    {0xb5b48465abbb7887ULL, 0xa72a6bc5e6ba2d2bULL, 0x4cd48043712f7b29ULL, 0x1142d5410fc1fc1aULL},
    {0x7ab2c156059075d3ULL, 0x17cb3594047999b2ULL, 0x44f2c93598f289f7ULL, 0x1d78439f69bc0becULL},
    {0x05d7a965138b8edbULL, 0x36ef35a3d55c48b1ULL, 0x8ddfb8a1ac6f1628ULL, 0x258588a508f4ff82ULL},
    {0x1596fb9afccb49e9ULL, 0x9a7367d69a09a95bULL, 0x9bc43f6984e4c157ULL, 0x13087879d2f514feULL},
    {0x295ccd233b4109faULL, 0xe1d72f89ed868012ULL, 0x2e9e1eea4bc88a8eULL, 0x17dadee898c45232ULL},
    {0x9a8590b4aa1f486fULL, 0xb75834b430e9130eULL, 0xb8e90b1034d5de31ULL, 0x295c6d1546e7f4a6ULL},
    {0x850adcb74c6eb892ULL, 0x07699ef305b92fc3ULL, 0x4ef96a2ba1720f2dULL, 0x1288ca0e1d3ed446ULL},
    {0x01960f9349d1b5eeULL, 0x8ccad30769371c69ULL, 0xe5c81e8991c98662ULL, 0x17563b4d1ae023f3ULL},
    {0x6ba01e9476b32917ULL, 0xa1cb0a3add977bc9ULL, 0x86815a945815f030ULL, 0x2869043be91a1eeaULL},
    {0x81776c885511d976ULL, 0x7475d34f47f414e7ULL, 0x5d090056095d96cfULL, 0x14941f0aff59e79aULL},
    {0xbc40b4fd8fc8c034ULL, 0xbb7142c3cce4fd48ULL, 0x318356758a39005aULL, 0x1ce337a190f4379fULL},
    {0x0000000000000000ULL, 0x0000000000000000ULL, 0x0000000000000000ULL, 0x0000000000000000ULL},
};

/* bar.rs:58-65 (sbox_8): the per-byte sbox applied to 8 bytes at once */
static inline uint64_t sbox8(uint64_t v) {
    uint64_t t1 = ((v & 0x8080808080808080ULL) >> 7) | ((v & 0x7f7f7f7f7f7f7f7fULL) << 1);
    uint64_t t2 = ((v & 0xc0c0c0c0c0c0c0c0ULL) >> 6) | ((v & 0x3f3f3f3f3f3f3f3fULL) << 2);
    uint64_t t3 = ((v & 0xe0e0e0e0e0e0e0e0ULL) >> 5) | ((v & 0x1f1f1f1f1f1f1f1fULL) << 3);
    uint64_t tmp = (~t1 & t2 & t3) ^ v;
    return ((tmp & 0x8080808080808080ULL) >> 7) | ((tmp & 0x7f7f7f7f7f7f7f7fULL) << 1);
}

/* bar on a canonical value (reference.rs:80-94 / bar.rs:15-31): swap 128-bit halves, sbox every
 * byte, reduce mod p. */
static inline fr_t sky_bar(fr_t x) {
    fr_t y = {{sbox8(x.l[2]), sbox8(x.l[3]), sbox8(x.l[0]), sbox8(x.l[1])}};
    while (fr_raw_geq_p(y.l)) fr_raw_sub_p(y.l);
    return y;
}

/* All values are raw canonical integers < p held in fr_t; fr_mul(x,x) on raw integers IS
 * x^2 * 2^-256 mod p, the spec's x^2 * sigma^-1 (reference.rs:22-26,63-69). */
static inline void sky_reduce_in(const uint64_t in[4], fr_t *o) {
    memcpy(o->l, in, 32);
    while (fr_raw_geq_p(o->l)) fr_raw_sub_p(o->l);
}

void orc_sky_permute(const uint64_t l_in[4], const uint64_t r_in[4], uint64_t l_out[4], uint64_t r_out[4]) {
    fr_t l, r;
    sky_reduce_in(l_in, &l);
    sky_reduce_in(r_in, &r);
    for (int i = 0; i < 18; i++) {
        fr_t f = (i == 6 || i == 7 || i == 10 || i == 11) ? sky_bar(l) : fr_mul(l, l);
        fr_t rc;
        memcpy(rc.l, SKY_RC[i], 32);
        fr_t nl = fr_add(fr_add(r, f), rc);
        r = l;
        l = nl;
    }
    memcpy(l_out, l.l, 32);
    memcpy(r_out, r.l, 32);
}

void orc_sky_compress(const uint64_t l_in[4], const uint64_t r_in[4], uint64_t out[4]) {
    fr_t t, l, r;
    sky_reduce_in(l_in, &t);
    orc_sky_permute(l_in, r_in, l.l, r.l);
    l = fr_add(l, t);
    memcpy(out, l.l, 32);
}

/* skyscraper/core/src/v1.rs:19-32 — fixture tests only. */
void orc_sky_compress_v1(const uint64_t l_in[4], const uint64_t r_in[4], uint64_t out[4]) {
    fr_t t, l, r;
    sky_reduce_in(l_in, &l);
    sky_reduce_in(r_in, &r);
    t = l;
    for (int i = 0; i < 10; i++) {
        fr_t f = (i == 2 || i == 3 || i == 6 || i == 7) ? sky_bar(l) : fr_mul(l, l);
        fr_t nl = fr_add(r, f);
        if (i < 9) {
            fr_t rc;
            memcpy(rc.l, SKY_RC[i], 32);
            nl = fr_add(nl, rc);
        }
        r = l;
        l = nl;
    }
    l = fr_add(l, t);
    memcpy(out, l.l, 32);
}

/* CompressManyFn contract: skyscraper/core/src/lib.rs:26, generic.rs:14-37. */
int orc_sky_compress_many(const uint8_t *messages, uint8_t *hashes, size_t n, int version) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        uint64_t m[8], h[4];
        memcpy(m, messages + 64 * i, 64);
        if (version == 1)
            orc_sky_compress_v1(m, m + 4, h);
        else
            orc_sky_compress(m, m + 4, h);
        memcpy(hashes + 32 * i, h, 32);
    }
    return 0;
}

/* provekit/common/src/skyscraper/whir.rs:20-25: Montgomery in, canonical compress, Montgomery out. */
fr_t orc_compress_fr(fr_t l, fr_t r) {
    uint64_t a[4], b[4], h[4];
    fr_to_canonical(l, a);
    fr_to_canonical(r, b);
    orc_sky_compress(a, b, h);
    return fr_from_canonical(h);
}

/* ---- PoW: skyscraper/core/src/pow.rs ------------------------------------------------------ */
#include <math.h>
/* pow.rs:61-82 */
static void f64_to_u256(double f, uint64_t out[4]) {
    uint64_t bits;
    memcpy(&bits, &f, 8);
    int sign = (int)(bits >> 63);
    int exp_bits = (int)((bits >> 52) & 0x7ff);
    uint64_t frac = bits & ((1ULL << 52) - 1);
    int exp;
    uint64_t sig;
    if (exp_bits == 0) { exp = -1022; sig = frac; } else { exp = exp_bits - 1023; sig = frac + (1ULL << 52); }
    memset(out, 0, 32);
    if (sign) return;
    if (exp > 256) { memset(out, 0xff, 32); return; }
    int shift = exp - 52;
    if (shift < 0) {
        out[0] = (uint64_t)round(f);
    } else {
        int limb = shift / 64, sh = shift % 64;
        out[limb] = sig << sh;
        if (sh != 0 && limb < 3) out[limb + 1] = sig >> (64 - sh);
    }
}
/* pow.rs:14-22 */
void orc_pow_threshold(double difficulty, uint64_t out[4]) {
    double modulus = (double)FR_P[3] * ldexp(1.0, 192);
    double prob = exp2(-difficulty);
    f64_to_u256(prob * modulus, out);
}
static int raw_less_than(const uint64_t a[4], const uint64_t b[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] < b[i]) return 1;
        if (a[i] > b[i]) return 0;
    }
    return 0;
}
/* pow.rs:24-26 */
int orc_pow_verify(const uint64_t challenge[4], double difficulty, uint64_t nonce) {
    if (difficulty == 0.0) return 1;
    uint64_t thr[4], n[4] = {nonce, 0, 0, 0}, h[4];
    orc_pow_threshold(difficulty, thr);
    orc_sky_compress(challenge, n, h);
    return raw_less_than(h, thr);
}
/* pow.rs:33-41 + generic.rs:42-71: the smallest accepted nonce (fetch_min semantics). */
uint64_t orc_pow_solve(const uint64_t challenge[4], double difficulty) {
    if (difficulty == 0.0) return 0;
    uint64_t thr[4];
    orc_pow_threshold(difficulty + 0.01, thr);
    uint64_t best = UINT64_MAX;
    const uint64_t BLOCK = 4096;
    for (uint64_t base = 0; best == UINT64_MAX; base += BLOCK) {
#pragma omp parallel for schedule(static)
        for (uint64_t k = 0; k < BLOCK; k++) {
            uint64_t n[4] = {base + k, 0, 0, 0}, h[4];
            orc_sky_compress(challenge, n, h);
            if (raw_less_than(h, thr)) {
#pragma omp critical
                if (base + k < best) best = base + k;
            }
        }
    }
    return best;
}
