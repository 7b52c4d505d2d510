/* oracle/transcript.c — see transcript.h.  TEST INFRASTRUCTURE ONLY. */
#include "transcript.h"

#include <stdio.h>

void orc_sky_permute(const uint64_t l[4], const uint64_t r[4], uint64_t lo[4], uint64_t ro[4]);

/* ---- Keccak-f[1600] ---- */
static const uint64_t KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
void orc_keccak_f1600(uint64_t s[25]) {
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rol64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) s[i] ^= D[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(s[x + 5 * y], KROT[x][y]);
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) s[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        s[0] ^= KRC[rnd];
    }
}
/* [EXT] spongefish generate_tag: unpadded Keccak duplex (rate 136) absorbs the domain separator in
 * overwrite mode, then squeezes 32 bytes (duplex discipline as keccakSponge.go:40-75). */
void orc_domsep_tag(const uint8_t *io, size_t len, uint8_t tag[32]) {
    uint64_t st[25];
    uint8_t *b = (uint8_t *)st;
    memset(st, 0, sizeof st);
    size_t ap = 0;
    for (size_t i = 0; i < len; i++) {
        if (ap == 136) {
            orc_keccak_f1600(st);
            ap = 0;
        }
        b[ap++] = io[i];
    }
    orc_keccak_f1600(st);
    memcpy(tag, b, 32);
}

/* ---- duplex sponge over Fr ---- */
static void sponge_permute(sponge_t *s) {
    uint64_t l[4], r[4], lo[4], ro[4];
    fr_to_canonical(s->st[0], l);
    fr_to_canonical(s->st[1], r);
    orc_sky_permute(l, r, lo, ro);
    s->st[0] = fr_from_canonical(lo);
    s->st[1] = fr_from_canonical(ro);
}
void sponge_init(sponge_t *s, const uint8_t iv[32]) {
    uint64_t c[4];
    memcpy(c, iv, 32);
    s->st[0] = FR_ZERO;
    s->st[1] = fr_from_canonical(c); /* sponge.rs:47-52 */
    s->absorb_pos = 0;
    s->squeeze_pos = 1;
}
void sponge_absorb(sponge_t *s, const fr_t *x, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (s->absorb_pos == 1) {
            sponge_permute(s);
            s->absorb_pos = 0;
        }
        s->st[0] = x[i];
        s->absorb_pos = 1;
    }
    if (n) s->squeeze_pos = 1;
}
void sponge_squeeze(sponge_t *s, fr_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (s->squeeze_pos == 1) {
            s->squeeze_pos = 0;
            s->absorb_pos = 0;
            sponge_permute(s);
        }
        out[i] = s->st[0];
        s->squeeze_pos = 1;
    }
}

/* ---- prover / verifier state ---- */
void fs_init(fs_state *fs, const uint8_t *domsep, size_t domsep_len, const uint8_t *proof, size_t proof_len) {
    uint8_t tag[32];
    memset(fs, 0, sizeof *fs);
    orc_domsep_tag(domsep, domsep_len, tag);
    sponge_init(&fs->sp, tag);
    if (proof) {
        fs->is_verifier = 1;
        fs->narg.p = (uint8_t *)proof;
        fs->narg.len = proof_len;
    }
}
void fs_init_foreign(fs_state *fs, const orc_transcript_vtbl *vt, void *user, int is_verifier) {
    memset(fs, 0, sizeof *fs);
    fs->vt = vt;
    fs->user = user;
    fs->is_verifier = is_verifier;
}
#define FS_FOREIGN(call)                 \
    do {                                 \
        if ((call) != 0) fs->failed = 1; \
        return;                          \
    } while (0)
void fs_add_scalars(fs_state *fs, const fr_t *x, size_t n) {
    if (fs->vt) FS_FOREIGN(fs->vt->add_scalars(fs->user, (const uint64_t *)x, n));
    sponge_absorb(&fs->sp, x, n);
    for (size_t i = 0; i < n; i++) bb_fr(&fs->narg, x[i]);
}
void fs_next_scalars(fs_state *fs, fr_t *x, size_t n) {
    if (fs->vt) {
        memset(x, 0, n * sizeof(fr_t));
        FS_FOREIGN(fs->vt->next_scalars(fs->user, (uint64_t *)x, n));
    }
    for (size_t i = 0; i < n; i++) {
        if (fs->rd + 32 > fs->narg.len) {
            fs->failed = 1;
            x[i] = FR_ZERO;
            continue;
        }
        uint64_t c[4];
        memcpy(c, fs->narg.p + fs->rd, 32);
        fs->rd += 32;
        if (fr_raw_geq_p(c)) fs->failed = 1; /* ark deserialisation rejects non-canonical scalars */
        x[i] = fr_from_canonical(c);
    }
    sponge_absorb(&fs->sp, x, n);
}
void fs_challenge_scalars(fs_state *fs, fr_t *out, size_t n) {
    if (fs->vt) {
        memset(out, 0, n * sizeof(fr_t));
        FS_FOREIGN(fs->vt->challenge_scalars(fs->user, (uint64_t *)out, n));
    }
    sponge_squeeze(&fs->sp, out, n);
}
void fs_challenge_bytes(fs_state *fs, uint8_t *out, size_t n) {
    if (fs->vt) {
        memset(out, 0, n);
        FS_FOREIGN(fs->vt->challenge_bytes(fs->user, out, n));
    }
    while (n) {
        fr_t u;
        uint64_t c[4];
        sponge_squeeze(&fs->sp, &u, 1);
        fr_to_canonical(u, c);
        size_t take = n < 15 ? n : 15; /* bytes_uniform_modp(254) = (254-128)/8 */
        memcpy(out, c, take);
        out += take;
        n -= take;
    }
}
void fs_add_bytes(fs_state *fs, const uint8_t *b, size_t n) {
    if (fs->vt) FS_FOREIGN(fs->vt->add_bytes(fs->user, b, n));
    for (size_t i = 0; i < n; i++) {
        fr_t u = fr_from_u64(b[i]);
        sponge_absorb(&fs->sp, &u, 1);
    }
    bb_push(&fs->narg, b, n);
}
void fs_next_bytes(fs_state *fs, uint8_t *b, size_t n) {
    if (fs->vt) {
        memset(b, 0, n);
        FS_FOREIGN(fs->vt->next_bytes(fs->user, b, n));
    }
    if (fs->rd + n > fs->narg.len) {
        fs->failed = 1;
        memset(b, 0, n);
        return;
    }
    memcpy(b, fs->narg.p + fs->rd, n);
    fs->rd += n;
    for (size_t i = 0; i < n; i++) {
        fr_t u = fr_from_u64(b[i]);
        sponge_absorb(&fs->sp, &u, 1);
    }
}
void fs_hint(fs_state *fs, const uint8_t *b, size_t n) {
    if (fs->vt) FS_FOREIGN(fs->vt->hint(fs->user, b, n));
    uint32_t len = (uint32_t)n;
    bb_push(&fs->narg, &len, 4);
    bb_push(&fs->narg, b, n);
}
const uint8_t *fs_next_hint(fs_state *fs, size_t *n) {
    if (fs->vt) {
        const uint8_t *p = NULL;
        *n = 0;
        if (fs->vt->next_hint(fs->user, &p, n) != 0 || (!p && *n)) {
            fs->failed = 1;
            *n = 0;
            return NULL;
        }
        return p;
    }
    if (fs->rd + 4 > fs->narg.len) {
        fs->failed = 1;
        *n = 0;
        return NULL;
    }
    uint32_t len;
    memcpy(&len, fs->narg.p + fs->rd, 4);
    if (fs->rd + 4 + len > fs->narg.len) {
        fs->failed = 1;
        *n = 0;
        return NULL;
    }
    const uint8_t *p = fs->narg.p + fs->rd + 4;
    fs->rd += 4 + (size_t)len;
    *n = len;
    return p;
}
