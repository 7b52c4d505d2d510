// provekit_b200/csrc/host/transcript.cpp — see transcript.hpp
#include "transcript.hpp"

namespace pkh {

// skyscraper/core/src/constants.rs:32-51
const uint64_t SKY_RC[18][4] = {
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

static inline uint64_t sbox8(uint64_t v) {  // bar.rs:58-65
    uint64_t t1 = ((v & 0x8080808080808080ULL) >> 7) | ((v & 0x7f7f7f7f7f7f7f7fULL) << 1);
    uint64_t t2 = ((v & 0xc0c0c0c0c0c0c0c0ULL) >> 6) | ((v & 0x3f3f3f3f3f3f3f3fULL) << 2);
    uint64_t t3 = ((v & 0xe0e0e0e0e0e0e0e0ULL) >> 5) | ((v & 0x1f1f1f1f1f1f1f1fULL) << 3);
    uint64_t tmp = (~t1 & t2 & t3) ^ v;
    return ((tmp & 0x8080808080808080ULL) >> 7) | ((tmp & 0x7f7f7f7f7f7f7f7fULL) << 1);
}

// reference.rs:49-98 on raw canonical integers held in Fr (mul(x,x) on raw ints = x^2 * 2^-256)
void sky_permute(uint64_t l_io[4], uint64_t r_io[4]) {
    Fr l, r;
    std::memcpy(l.l, l_io, 32);
    std::memcpy(r.l, r_io, 32);
    while (geq_p(l.l)) sub_p(l.l);
    while (geq_p(r.l)) sub_p(r.l);
    for (int i = 0; i < 18; i++) {
        Fr f;
        if (i == 6 || i == 7 || i == 10 || i == 11) {
            f = Fr{{sbox8(l.l[2]), sbox8(l.l[3]), sbox8(l.l[0]), sbox8(l.l[1])}};
            while (geq_p(f.l)) sub_p(f.l);
        } else {
            f = mul(l, l);
        }
        Fr rc;
        std::memcpy(rc.l, SKY_RC[i], 32);
        Fr nl = add(add(r, f), rc);
        r = l;
        l = nl;
    }
    std::memcpy(l_io, l.l, 32);
    std::memcpy(r_io, r.l, 32);
}

static const uint64_t KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
void keccak_f1600(uint64_t s[25]) {
    auto rol = [](uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; };
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t C[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) {
            uint64_t d = C[(x + 4) % 5] ^ rol(C[(x + 1) % 5], 1);
            for (int y = 0; y < 5; y++) s[x + 5 * y] ^= d;
        }
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rol(s[x + 5 * y], KROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) s[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        s[0] ^= KRC[rnd];
    }
}
void domsep_tag(const std::string& io, uint8_t tag[32]) {
    uint64_t st[25] = {0};
    uint8_t* b = reinterpret_cast<uint8_t*>(st);
    size_t ap = 0;
    for (unsigned char ch : io) {
        if (ap == 136) {
            keccak_f1600(st);
            ap = 0;
        }
        b[ap++] = ch;
    }
    keccak_f1600(st);
    std::memcpy(tag, b, 32);
}

void Sponge::init(const uint8_t iv[32]) {
    uint64_t c[4];
    std::memcpy(c, iv, 32);
    st_[0] = ZERO;
    st_[1] = from_canonical(c);
    absorb_pos_ = 0;
    squeeze_pos_ = 1;
}
void Sponge::permute() {
    uint64_t l[4], r[4];
    to_canonical(st_[0], l);
    to_canonical(st_[1], r);
    sky_permute(l, r);
    st_[0] = from_canonical(l);
    st_[1] = from_canonical(r);
}
void Sponge::absorb(const Fr* x, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (absorb_pos_ == 1) {
            permute();
            absorb_pos_ = 0;
        }
        st_[0] = x[i];
        absorb_pos_ = 1;
    }
    if (n) squeeze_pos_ = 1;
}
void Sponge::squeeze(Fr* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (squeeze_pos_ == 1) {
            squeeze_pos_ = 0;
            absorb_pos_ = 0;
            permute();
        }
        out[i] = st_[0];
        squeeze_pos_ = 1;
    }
}

ProverState::ProverState(const std::string& domsep) {
    uint8_t tag[32];
    domsep_tag(domsep, tag);
    sp_.init(tag);
}
void ProverState::add_scalars(const Fr* x, size_t n) {
    sp_.absorb(x, n);
    for (size_t i = 0; i < n; i++) put_fr(narg_, x[i]);
}
void ProverState::challenge_scalars(Fr* out, size_t n) { sp_.squeeze(out, n); }
void ProverState::challenge_bytes(uint8_t* out, size_t n) {
    while (n) {
        Fr u;
        uint64_t c[4];
        sp_.squeeze(&u, 1);
        to_canonical(u, c);
        size_t take = n < 15 ? n : 15;
        std::memcpy(out, c, take);
        out += take;
        n -= take;
    }
}
void ProverState::add_bytes(const uint8_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fr u = from_u64(b[i]);
        sp_.absorb(&u, 1);
    }
    narg_.insert(narg_.end(), b, b + n);
}
void ProverState::hint(const std::vector<uint8_t>& payload) {
    uint32_t len = (uint32_t)payload.size();
    for (int i = 0; i < 4; i++) narg_.push_back((uint8_t)(len >> (8 * i)));
    narg_.insert(narg_.end(), payload.begin(), payload.end());
}

}  // namespace pkh
