// provekit_b200/csrc/host/nps_file.cpp — R1CS out of the `.nps` scheme container.
//
// Container (provekit/common/src/file/bin.rs:16-18,22-60, file/mod.rs:27-29): 8 B magic "\xDC\xDFOZkp\x01\x00" |
// 8 B format "NrProScm" | u16-LE major | u16-LE minor | zstd( postcard(NoirProofScheme) ) with
//   NoirProofScheme { program: acir Program, r1cs: R1CS, witness_builders, witness_generator, whir_for_witness }
//   (provekit/common/src/noir_proof_scheme.rs:16-23)
//   R1CS { num_public_inputs: usize, interner: Interner, a, b, c: SparseMatrix }            (r1cs.rs:7-13)
//   Interner { values: Vec<Fr> via serde_ark } = postcard bytes( u64-LE count | count x 32 B canonical LE )
//                                                                        (interner.rs:6-10, utils/serde_ark.rs:11-30)
//   SparseMatrix { num_rows, num_cols: usize, new_row_indices: Vec<u32>, col_indices: Vec<u32>,
//                  values: Vec<InternedFieldElement(usize)> }           (sparse_matrix.rs:10-27), postcard varints.
// The ACIR `Program` in front of the R1CS is an [EXT] schema (acir crate) this library does not restate, so the R1CS is
// LOCATED rather than reached by parsing: the interner is the first byte string whose postcard length 8 + 32c is
// followed by the u64 c, whose c elements are all canonical (< p) and which is followed by three well-formed sparse
// matrices of identical shape (row starts monotone, columns < num_cols, values < c).  This is the recipe of SURVEY A.2,
// which finds the fixture's R1CS (729 560 x 860 637, nnz 740 508 / 609 440 / 1 915 568, 366 constants).
// Everything behind the R1CS (witness builders, ABI, WhirConfigs) is not needed: pk_prover_create derives m, m_0 and
// both WhirConfigs from the R1CS shape exactly as provekit/r1cs-compiler/src/whir_r1cs.rs:15-54 does.
#include <dlfcn.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../../include/pkwhir.h"
#include "fr_host.h"

struct pk_nps {
    pk_r1cs r1cs;
    int64_t num_public_inputs = -1;
    std::vector<uint64_t> interned;
    std::vector<uint64_t> row_start[3];
    std::vector<uint32_t> col[3], val[3];
};

namespace {

struct ZBuf {
    void* p;
    size_t size, pos;
};
const uint8_t MAGIC[8] = {0xDC, 0xDF, 'O', 'Z', 'k', 'p', 0x01, 0x00};
const char FORMAT_NPS[9] = "NrProScm";
const size_t NPS_MAX_INFLATED = (size_t)8 << 30;  // the reference fixture inflates to ~70 MB

bool zstd_inflate(const uint8_t* src, size_t len, std::vector<uint8_t>& out) {
    static void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);  // once per process (thread-safe static init)
    if (!h) return false;
    auto createDStream = (void* (*)())dlsym(h, "ZSTD_createDStream");
    auto initDStream = (size_t(*)(void*))dlsym(h, "ZSTD_initDStream");
    auto decompressStream = (size_t(*)(void*, ZBuf*, ZBuf*))dlsym(h, "ZSTD_decompressStream");
    auto freeDStream = (size_t(*)(void*))dlsym(h, "ZSTD_freeDStream");
    auto isError = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
    if (!createDStream || !initDStream || !decompressStream || !freeDStream || !isError) return false;
    void* ds = createDStream();
    if (!ds) return false;
    initDStream(ds);
    std::vector<uint8_t> chunk(4 << 20);
    ZBuf in = {(void*)src, len, 0};
    bool ok = true;
    for (;;) {
        ZBuf o = {chunk.data(), chunk.size(), 0};
        size_t r = decompressStream(ds, &o, &in);
        if (isError(r)) {
            ok = false;
            break;
        }
        out.insert(out.end(), chunk.begin(), chunk.begin() + o.pos);
        if (out.size() > NPS_MAX_INFLATED) {  // a crafted frame must not exhaust host memory
            ok = false;
            break;
        }
        if (r == 0) break;  // frame complete
        if (in.pos == in.size && o.pos < o.size) {
            ok = false;  // input exhausted in the middle of a frame: truncated file
            break;
        }
    }
    freeDStream(ds);
    return ok;
}

// postcard varint; false on overrun / more than 10 bytes
inline bool varint(const uint8_t* b, size_t n, size_t& pos, uint64_t& v) {
    v = 0;
    for (int shift = 0; shift < 70; shift += 7) {
        if (pos >= n) return false;
        uint8_t c = b[pos++];
        v |= (uint64_t)(c & 0x7f) << shift;
        if (!(c & 0x80)) return true;
    }
    return false;
}
inline uint64_t le64(const uint8_t* p) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    return v;
}

// one SparseMatrix at pos; fills row_start / col / val, advances pos
bool parse_matrix(const uint8_t* b, size_t n, size_t& pos, uint64_t n_interned, uint64_t& rows, uint64_t& cols,
                  std::vector<uint64_t>& row_start, std::vector<uint32_t>& col, std::vector<uint32_t>& val) {
    uint64_t cnt, v;
    if (!varint(b, n, pos, rows) || !varint(b, n, pos, cols) || rows == 0 || cols == 0 || rows > (1ull << 32) || cols > (1ull << 32))
        return false;
    if (!varint(b, n, pos, cnt) || cnt != rows || pos + cnt > n) return false;
    row_start.resize(rows);
    uint64_t prev = 0;
    for (uint64_t i = 0; i < rows; i++) {
        if (!varint(b, n, pos, v) || v < prev || v > 0xffffffffull) return false;
        row_start[i] = prev = v;
    }
    if (!varint(b, n, pos, cnt) || cnt < prev || cnt > 0xffffffffull || pos + cnt > n) return false;
    const uint64_t nnz = cnt;
    col.resize(nnz);
    for (uint64_t i = 0; i < nnz; i++) {
        if (!varint(b, n, pos, v) || v >= cols) return false;
        col[i] = (uint32_t)v;
    }
    if (!varint(b, n, pos, cnt) || cnt != nnz || pos + cnt > n) return false;
    val.resize(nnz);
    for (uint64_t i = 0; i < nnz; i++) {
        if (!varint(b, n, pos, v) || v >= n_interned) return false;
        val[i] = (uint32_t)v;
    }
    return true;
}

}  // namespace

extern "C" {

int pk_nps_read_r1cs(const uint8_t* file, size_t len, pk_nps** out) {
    if (!file || !out) return PK_ERR_INVALID_ARG;
    *out = nullptr;
    // bin.rs:86-99: magic, format tag, major version must match (0), any minor
    if (len < 20 || std::memcmp(file, MAGIC, 8) != 0 || std::memcmp(file + 8, FORMAT_NPS, 8) != 0 || file[16] != 0 || file[17] != 0)
        return PK_ERR_INVALID_ARG;
    std::vector<uint8_t> raw;
    if (!zstd_inflate(file + 20, len - 20, raw)) return PK_ERR_INTERNAL;
    const uint8_t* b = raw.data();
    const size_t n = raw.size();
    pk_nps* s = new pk_nps();
    pk_nps* found = nullptr;
    for (size_t at = 1; at + 48 < n; at++) {
        // cheap reject first: a length varint of 2..5 bytes whose value is 8 + 32c, followed by u64 c
        if (!(b[at] & 0x80)) continue;  // 8 + 32c >= 40 < 128 only for c = 1..3: a scheme has more constants
        size_t pos = at;
        uint64_t L;
        if (!varint(b, n, pos, L) || L < 40 || ((L - 8) & 31) != 0 || pos + L > n) continue;
        const uint64_t c = (L - 8) / 32;
        if (le64(b + pos) != c) continue;
        const uint8_t* elems = b + pos + 8;
        bool canonical = true;
        for (uint64_t i = 0; i < c && canonical; i++) {
            uint64_t limbs[4];
            std::memcpy(limbs, elems + 32 * i, 32);
            canonical = !pkh::geq_p(limbs);
        }
        if (!canonical) continue;
        size_t p2 = pos + L;
        uint64_t rows[3], cols[3];
        bool ok = true;
        for (int k = 0; k < 3 && ok; k++)
            ok = parse_matrix(b, n, p2, c, rows[k], cols[k], s->row_start[k], s->col[k], s->val[k]) &&
                 rows[k] == rows[0] && cols[k] == cols[0];
        if (!ok) continue;
        s->interned.resize(4 * c);
        for (uint64_t i = 0; i < c; i++) {
            uint64_t limbs[4];
            std::memcpy(limbs, elems + 32 * i, 32);
            pkh::Fr m = pkh::from_canonical(limbs);
            std::memcpy(&s->interned[4 * i], m.l, 32);
        }
        // num_public_inputs is the varint right in front of the interner; decodable backwards only when it is one byte
        // and the byte before it ends a varint too
        if (at >= 2 && !(b[at - 1] & 0x80) && !(b[at - 2] & 0x80)) s->num_public_inputs = b[at - 1];
        s->r1cs.num_constraints = rows[0];
        s->r1cs.num_witnesses = cols[0];
        s->r1cs.num_interned = c;
        s->r1cs.interned = s->interned.data();
        pk_csr* m[3] = {&s->r1cs.a, &s->r1cs.b, &s->r1cs.c};
        for (int k = 0; k < 3; k++) {
            m[k]->num_rows = rows[k];
            m[k]->num_cols = cols[k];
            m[k]->nnz = s->col[k].size();
            m[k]->row_start = s->row_start[k].data();
            m[k]->col = s->col[k].data();
            m[k]->val = s->val[k].data();
        }
        if (found) {  // a second well-formed candidate: the heuristic is ambiguous on this stream, refuse to guess
            delete found;
            delete s;
            return PK_ERR_INVALID_ARG;
        }
        found = s;
        s = new pk_nps();
        at = p2 - 1;  // keep scanning behind the matched R1CS
    }
    delete s;
    if (!found) return PK_ERR_INVALID_ARG;
    *out = found;
    return PK_OK;
}
const pk_r1cs* pk_nps_r1cs(const pk_nps* s) { return s ? &s->r1cs : nullptr; }
int64_t pk_nps_num_public_inputs(const pk_nps* s) { return s ? s->num_public_inputs : -1; }
void pk_nps_free(pk_nps* s) { delete s; }

}  // extern "C"
