/* oracle/transcript.h — CPU restatement of the Fiat-Shamir layer (spongefish [EXT] + the in-tree
 * Skyscraper sponge plug-in).  TEST INFRASTRUCTURE ONLY — see oracle/pk_oracle.h.
 *
 * In-tree, certain:
 *   - permutation: state = 2 field elements, rate 1 (cell 0), capacity 1; new(iv) = [0, Fr(iv as LE int)];
 *     permute = skyscraper::reference::permute (provekit/common/src/skyscraper/sponge.rs:24-58)
 *   - duplex discipline in overwrite mode, mirrored for Keccak by
 *     recursive-verifier/app/keccakSponge/keccakSponge.go:40-75
 *   - framing of the proof string: scalars 32 B canonical LE; hints u32-LE length + payload;
 *     "pow-nonce" 8 raw bytes big-endian (recursive-verifier/app/circuit/common.go:30-105)
 *   - a digest is one scalar (provekit/common/src/skyscraper/whir.rs:88-111)
 * [EXT], PARITY UNPINNED (recalled from upstream spongefish / whir; challenge VALUES depend on them):
 *   - IV = first 32 bytes squeezed from an (unpadded) Keccak-f1600 duplex that absorbed the
 *     domain-separator string; the string's per-op encoding "\0" + {A,S,H} + count + label and whir's labels
 *   - challenge_bytes squeezes one unit per 15 bytes and takes the low bytes of its canonical LE form;
 *     add_bytes absorbs one unit per byte.
 */
#ifndef PK_ORACLE_TRANSCRIPT_H
#define PK_ORACLE_TRANSCRIPT_H
#include <stdio.h>
#include <stdlib.h>

#include "fr.h"
#include "pk_oracle.h"

typedef struct {
    uint8_t *p;
    size_t len, cap;
} bytebuf;
static inline void bb_push(bytebuf *b, const void *src, size_t n) {
    if (b->len + n > b->cap) {
        b->cap = (b->len + n) * 2 + 64;
        b->p = (uint8_t *)realloc(b->p, b->cap);
    }
    memcpy(b->p + b->len, src, n);
    b->len += n;
}
static inline void bb_u64(bytebuf *b, uint64_t v) { bb_push(b, &v, 8); }
static inline void bb_fr(bytebuf *b, fr_t x) {
    uint64_t c[4];
    fr_to_canonical(x, c);
    bb_push(b, c, 32);
}
static inline void bb_str(bytebuf *b, const char *s) { bb_push(b, s, strlen(s)); }

/* ---- Keccak-f[1600] (only for the IV) ---- */
void orc_keccak_f1600(uint64_t st[25]);
void orc_domsep_tag(const uint8_t *io, size_t len, uint8_t tag[32]);

/* ---- duplex sponge over Fr, N = 2, R = 1 ---- */
typedef struct {
    fr_t st[2];
    int absorb_pos, squeeze_pos;
} sponge_t;
void sponge_init(sponge_t *s, const uint8_t iv[32]);
void sponge_absorb(sponge_t *s, const fr_t *x, size_t n);
void sponge_squeeze(sponge_t *s, fr_t *out, size_t n);

/* ---- prover / verifier state ---- */
typedef struct {
    sponge_t sp;
    bytebuf narg;          /* prover: grows; verifier: wraps the proof */
    size_t rd;             /* verifier read cursor */
    int is_verifier, failed;
    /* foreign transcript (orc_prove_with_transcript / orc_verify_with_transcript): every operation is forwarded to the
     * caller's ProverState / VerifierState instead of the in-tree sponge; a non-zero return marks the run failed */
    const orc_transcript_vtbl *vt;
    void *user;
} fs_state;
void fs_init_foreign(fs_state *fs, const orc_transcript_vtbl *vt, void *user, int is_verifier);
void fs_init(fs_state *fs, const uint8_t *domsep, size_t domsep_len, const uint8_t *proof, size_t proof_len);
void fs_add_scalars(fs_state *fs, const fr_t *x, size_t n);         /* prover */
void fs_next_scalars(fs_state *fs, fr_t *x, size_t n);              /* verifier */
void fs_challenge_scalars(fs_state *fs, fr_t *out, size_t n);
void fs_challenge_bytes(fs_state *fs, uint8_t *out, size_t n);
void fs_add_bytes(fs_state *fs, const uint8_t *b, size_t n);        /* prover */
void fs_next_bytes(fs_state *fs, uint8_t *b, size_t n);             /* verifier */
void fs_hint(fs_state *fs, const uint8_t *b, size_t n);             /* prover */
const uint8_t *fs_next_hint(fs_state *fs, size_t *n);               /* verifier */

/* ---- domain separator string builder ---- */
static inline void ds_op(bytebuf *b, char kind, size_t count, const char *label) {
    char tmp[32];
    uint8_t z = 0;
    bb_push(b, &z, 1);
    if (kind == 'H')
        snprintf(tmp, sizeof tmp, "H");
    else
        snprintf(tmp, sizeof tmp, "%c%zu", kind, count);
    bb_str(b, tmp);
    bb_str(b, label);
}
#endif
