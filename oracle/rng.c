/* oracle/rng.c — CPU restatement of the mask generator `pk_rng_fill` (ChaCha12 counter stream -> uniform Fr).
 * TEST INFRASTRUCTURE ONLY — see oracle/pk_oracle.h.
 *
 * The reference draws its masks with `F::rand(&mut thread_rng())` (provekit/common/src/utils/zk_utils.rs:13-22,
 * provekit/prover/src/whir_r1cs.rs:211-225): thread_rng is a ChaCha12 stream, Fp::rand rejection-samples
 * 254-bit strings below p.  The device path keeps that construction but makes it counter based so that every
 * element is independent (one thread each):
 *     block(i, a) = ChaCha12(key = seed, words 12..15 = (i & 0xffffffff, i >> 32, stream, a))
 *     candidates of element i, in order: block(i,0)[0..8), block(i,0)[8..16), block(i,1)[0..8), ...
 *     each candidate = 8 LE words with the top word masked to 30 bits; the first one < p is the element's
 *     32-byte in-memory (Montgomery) representation, which is uniform over Fr because x -> xR is a bijection.
 * The block function is RFC 8439's (pinned by its section 2.3.2 vector with rounds = 20).
 */
#include <string.h>

#include "fr.h"
#include "pk_oracle.h"

#define ROTL(x, n) (((x) << (n)) | ((x) >> (32 - (n))))
#define QR(a, b, c, d)                 \
    a += b; d ^= a; d = ROTL(d, 16);   \
    c += d; b ^= c; b = ROTL(b, 12);   \
    a += b; d ^= a; d = ROTL(d, 8);    \
    c += d; b ^= c; b = ROTL(b, 7)

void orc_chacha_block(const uint32_t in[16], int rounds, uint32_t out[16]) {
    uint32_t x[16];
    memcpy(x, in, 64);
    for (int r = 0; r < rounds; r += 2) {
        QR(x[0], x[4], x[8], x[12]);
        QR(x[1], x[5], x[9], x[13]);
        QR(x[2], x[6], x[10], x[14]);
        QR(x[3], x[7], x[11], x[15]);
        QR(x[0], x[5], x[10], x[15]);
        QR(x[1], x[6], x[11], x[12]);
        QR(x[2], x[7], x[8], x[13]);
        QR(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

static const uint32_t P32[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};

static int below_p(const uint32_t c[8]) {
    for (int i = 7; i >= 0; i--) {
        if (c[i] < P32[i]) return 1;
        if (c[i] > P32[i]) return 0;
    }
    return 0;
}

void orc_rng_fill(uint64_t *out, size_t n, const uint8_t seed[32], uint32_t stream) {
    uint32_t key[8];
    for (int i = 0; i < 8; i++)
        key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
                 ((uint32_t)seed[4 * i + 3] << 24);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        memcpy(in + 4, key, 32);
        in[12] = (uint32_t)i;
        in[13] = (uint32_t)((uint64_t)i >> 32);
        in[14] = stream;
        for (uint32_t attempt = 0;; attempt++) {
            uint32_t blk[16];
            in[15] = attempt;
            orc_chacha_block(in, 12, blk);
            int done = 0;
            for (int c = 0; c < 2 && !done; c++) {
                uint32_t *cand = blk + 8 * c;
                cand[7] &= 0x3fffffffu;
                if (below_p(cand)) {
                    for (int k = 0; k < 4; k++) out[4 * i + k] = (uint64_t)cand[2 * k] | ((uint64_t)cand[2 * k + 1] << 32);
                    done = 1;
                }
            }
            if (done) break;
        }
    }
}

/* the five mask arrays of one proof, drawn from one seed (streams 0..4 in the order of orc_rand) */
void orc_rng_masks(const uint8_t seed[32], int m, int m0, int mh, uint64_t *mask_w, uint64_t *g_w, uint64_t *blind,
                   uint64_t *mask_h, uint64_t *g_h) {
    orc_rng_fill(mask_w, (size_t)1 << (m - 1), seed, 0);
    orc_rng_fill(g_w, (size_t)1 << m, seed, 1);
    orc_rng_fill(blind, 4 * (size_t)m0, seed, 2);
    orc_rng_fill(mask_h, (size_t)1 << (mh - 1), seed, 3);
    orc_rng_fill(g_h, (size_t)1 << mh, seed, 4);
}
