// provekit_b200/csrc/host/np_file.cpp — the `.np` proof container.
//
// Layout restated from provekit/common/src/file/bin.rs:16-18,22-60 and file/mod.rs:33-37:
//   8 B magic "\xDC\xDFOZkp\x01\x00" | 8 B format "NPSProof" | u16-LE major 0 | u16-LE minor 0 | zstd( postcard(NoirProof) )
// and NoirProof { whir_r1cs_proof: WhirR1CSProof { transcript: Vec<u8> } } is postcard varint(len) + raw bytes
// (checked against the reference fixture tooling/provekit-bench/benches/poseidon-1000.np, tests/test_np_file.py).
// zstd is taken from the system's libzstd.so.1 at run time (the image ships the runtime library but no headers).
#include <dlfcn.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../../include/pkwhir.h"

namespace {

struct ZBuf {
    void* p;
    size_t size, pos;
};
struct Zstd {
    size_t (*compressBound)(size_t) = nullptr;
    size_t (*compress)(void*, size_t, const void*, size_t, int) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    void* (*createDStream)() = nullptr;
    size_t (*initDStream)(void*) = nullptr;
    size_t (*decompressStream)(void*, ZBuf*, ZBuf*) = nullptr;
    size_t (*freeDStream)(void*) = nullptr;
    bool ok = false;
};
Zstd& zstd() {
    static Zstd z;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        z.compressBound = (size_t(*)(size_t))dlsym(h, "ZSTD_compressBound");
        z.compress = (size_t(*)(void*, size_t, const void*, size_t, int))dlsym(h, "ZSTD_compress");
        z.isError = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
        z.createDStream = (void* (*)())dlsym(h, "ZSTD_createDStream");
        z.initDStream = (size_t(*)(void*))dlsym(h, "ZSTD_initDStream");
        z.decompressStream = (size_t(*)(void*, ZBuf*, ZBuf*))dlsym(h, "ZSTD_decompressStream");
        z.freeDStream = (size_t(*)(void*))dlsym(h, "ZSTD_freeDStream");
        z.ok = z.compressBound && z.compress && z.isError && z.createDStream && z.initDStream && z.decompressStream && z.freeDStream;
    });
    return z;
}
const uint8_t MAGIC[8] = {0xDC, 0xDF, 'O', 'Z', 'k', 'p', 0x01, 0x00};
const char FORMAT_NP[9] = "NPSProof";

}  // namespace

extern "C" {

int pk_np_encode(const uint8_t* transcript, size_t len, uint8_t** out, size_t* out_len) {
    if ((!transcript && len) || !out || !out_len) return PK_ERR_INVALID_ARG;
    Zstd& z = zstd();
    if (!z.ok) return PK_ERR_INTERNAL;
    std::vector<uint8_t> payload;
    for (size_t v = len;;) {  // postcard varint
        uint8_t b = v & 0x7f;
        v >>= 7;
        payload.push_back(v ? (b | 0x80) : b);
        if (!v) break;
    }
    payload.insert(payload.end(), transcript, transcript + len);
    size_t bound = z.compressBound(payload.size());
    uint8_t* buf = (uint8_t*)std::malloc(20 + bound);
    if (!buf) return PK_ERR_OOM;
    std::memcpy(buf, MAGIC, 8);
    std::memcpy(buf + 8, FORMAT_NP, 8);
    std::memset(buf + 16, 0, 4);  // version (0, 0)
    size_t n = z.compress(buf + 20, bound, payload.data(), payload.size(), 3);  // zstd::DEFAULT_COMPRESSION_LEVEL
    if (z.isError(n)) {
        std::free(buf);
        return PK_ERR_INTERNAL;
    }
    *out = buf;
    *out_len = 20 + n;
    return PK_OK;
}

static const size_t NP_MAX_INFLATED = (size_t)2 << 30;  // proofs are a few hundred KB

int pk_np_decode(const uint8_t* file, size_t len, uint8_t** transcript, size_t* transcript_len) {
    if (!file || !transcript || !transcript_len) return PK_ERR_INVALID_ARG;
    // bin.rs:86-99: magic, format, major == 0, minor >= 0
    if (len < 20 || std::memcmp(file, MAGIC, 8) != 0 || std::memcmp(file + 8, FORMAT_NP, 8) != 0 || file[16] != 0 || file[17] != 0)
        return PK_ERR_INVALID_ARG;
    Zstd& z = zstd();
    if (!z.ok) return PK_ERR_INTERNAL;
    void* ds = z.createDStream();
    if (!ds) return PK_ERR_OOM;
    z.initDStream(ds);
    std::vector<uint8_t> payload;
    std::vector<uint8_t> chunk(1 << 20);
    ZBuf in = {(void*)(file + 20), len - 20, 0};
    int rc = PK_OK;
    for (;;) {
        ZBuf o = {chunk.data(), chunk.size(), 0};
        size_t r = z.decompressStream(ds, &o, &in);
        if (z.isError(r)) {
            rc = PK_ERR_INVALID_ARG;
            break;
        }
        payload.insert(payload.end(), chunk.begin(), chunk.begin() + o.pos);
        if (payload.size() > NP_MAX_INFLATED) {  // a crafted frame must not exhaust host memory
            rc = PK_ERR_INVALID_ARG;
            break;
        }
        if (r == 0) break;  // frame complete
        if (in.pos == in.size && o.pos < o.size) {
            rc = PK_ERR_INVALID_ARG;  // input exhausted in the middle of a frame: truncated file
            break;
        }
    }
    z.freeDStream(ds);
    if (rc != PK_OK) return rc;
    size_t n = 0, pos = 0;
    for (int shift = 0;; shift += 7) {
        if (pos >= payload.size() || shift > 63) return PK_ERR_INVALID_ARG;
        uint8_t b = payload[pos++];
        n |= (size_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
    }
    if (pos + n != payload.size()) return PK_ERR_INVALID_ARG;
    uint8_t* buf = (uint8_t*)std::malloc(n ? n : 1);
    if (!buf) return PK_ERR_OOM;
    std::memcpy(buf, payload.data() + pos, n);
    *transcript = buf;
    *transcript_len = n;
    return PK_OK;
}

}  // extern "C"
