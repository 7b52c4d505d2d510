// provekit_b200/csrc/host/transcript.hpp — host-side Fiat-Shamir transcript of the C++ harness.
//
// In a ProveKit deployment the Rust host keeps `ProverState<SkyscraperSponge, FieldElement>`
// (provekit/prover/src/whir_r1cs.rs:57-59) and calls the fine-grained C-ABI; this file is the C++
// stand-in used by pk_prove so that the whole path can be driven and measured without a Rust toolchain.
// Restates: provekit/common/src/skyscraper/sponge.rs:24-58 (permutation plug-in: 2 cells, rate 1,
// state = [0, Fr(iv)]), skyscraper/core/src/reference.rs:49-98 (permute), and spongefish's duplex /
// codec rules [EXT, recalled; see DESIGN.md "parity unpinned"]: overwrite-mode absorb, permute on squeeze,
// IV = Keccak duplex of the domain separator, scalars = 32 B canonical LE, hints = u32-LE length + payload,
// challenge bytes = low 15 bytes of one squeezed unit each, add_bytes = one unit per byte.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/pkwhir.h"
#include "fr_host.h"

namespace pkh {

// ---- Skyscraper-v2 permutation on the host (sequential; a few hundred calls per proof) ----
extern const uint64_t SKY_RC[18][4];
void sky_permute(uint64_t l[4], uint64_t r[4]);  // canonical in/out (reduced mod p first)

// ---- Keccak-f[1600], only for the 32-byte IV ----
void keccak_f1600(uint64_t st[25]);
void domsep_tag(const std::string& io, uint8_t tag[32]);

class Sponge {
   public:
    void init(const uint8_t iv[32]);
    void absorb(const Fr* x, size_t n);
    void squeeze(Fr* out, size_t n);

   private:
    Fr st_[2];
    int absorb_pos_ = 0, squeeze_pos_ = 1;
    void permute();
};

// The spongefish ProverState surface the path uses (provekit/prover/src/whir_r1cs.rs:240-242,268-272,335-337;
// pattern provekit/common/src/whir_r1cs.rs:28-39).  Every operation reports failure through ok().
class Transcript {
   public:
    virtual ~Transcript() {}
    virtual void add_scalars(const Fr* x, size_t n) = 0;
    virtual void challenge_scalars(Fr* out, size_t n) = 0;
    virtual void challenge_bytes(uint8_t* out, size_t n) = 0;
    virtual void add_bytes(const uint8_t* b, size_t n) = 0;
    virtual void hint(const std::vector<uint8_t>& payload) = 0;
    virtual bool ok() const { return true; }
};

// the in-tree default: Skyscraper duplex sponge + proof string kept here
class ProverState : public Transcript {
   public:
    explicit ProverState(const std::string& domsep);
    void add_scalars(const Fr* x, size_t n) override;
    void challenge_scalars(Fr* out, size_t n) override;
    void challenge_bytes(uint8_t* out, size_t n) override;
    void add_bytes(const uint8_t* b, size_t n) override;
    void hint(const std::vector<uint8_t>& payload) override;
    std::vector<uint8_t>& narg() { return narg_; }

   private:
    Sponge sp_;
    std::vector<uint8_t> narg_;
};

// the host's own ProverState behind pk_transcript_vtbl (include/pkwhir.h): sponge, codecs and the proof string all
// live on the caller's side
class CallbackTranscript : public Transcript {
   public:
    CallbackTranscript(const pk_transcript_vtbl* vt, void* user) : vt_(vt), user_(user) {}
    void add_scalars(const Fr* x, size_t n) override { note(vt_->add_scalars(user_, x[0].l, n)); }
    void challenge_scalars(Fr* out, size_t n) override {
        std::memset(out, 0, n * sizeof(Fr));
        note(vt_->challenge_scalars(user_, out[0].l, n));
    }
    void challenge_bytes(uint8_t* out, size_t n) override {
        std::memset(out, 0, n);
        note(vt_->challenge_bytes(user_, out, n));
    }
    void add_bytes(const uint8_t* b, size_t n) override { note(vt_->add_bytes(user_, b, n)); }
    void hint(const std::vector<uint8_t>& payload) override { note(vt_->hint(user_, payload.data(), payload.size())); }
    bool ok() const override { return ok_; }

   private:
    const pk_transcript_vtbl* vt_;
    void* user_;
    bool ok_ = true;
    void note(int rc) {
        if (rc != 0) ok_ = false;
    }
};

// domain-separator builder: "\0" + {A,S,H} + count + label per op
class DomSep {
   public:
    explicit DomSep(const std::string& session) : io_(session) {}
    DomSep& absorb(size_t count, const char* label) { return op('A', count, label); }
    DomSep& squeeze(size_t count, const char* label) { return op('S', count, label); }
    DomSep& hint(const char* label) {
        io_.push_back('\0');
        io_ += "H";
        io_ += label;
        return *this;
    }
    const std::string& str() const { return io_; }

   private:
    std::string io_;
    DomSep& op(char k, size_t count, const char* label) {
        io_.push_back('\0');
        io_.push_back(k);
        io_ += std::to_string(count);
        io_ += label;
        return *this;
    }
};

inline void put_u64(std::vector<uint8_t>& v, uint64_t x) {
    for (int i = 0; i < 8; i++) v.push_back((uint8_t)(x >> (8 * i)));
}
inline void put_canonical(std::vector<uint8_t>& v, const uint64_t c[4]) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(c);
    v.insert(v.end(), p, p + 32);
}
inline void put_fr(std::vector<uint8_t>& v, const Fr& x) {
    uint64_t c[4];
    to_canonical(x, c);
    put_canonical(v, c);
}

}  // namespace pkh
