"""CPU parity of the C++ host transcript (provekit_b200/csrc/host/transcript.cpp, what pk_prove drives) with the oracle's
C restatement (oracle/transcript.c): both are compiled into tiny host-only drivers that run the same scripted sequence of
transcript operations (scalars, challenge scalars, challenge bytes, byte absorption as for the PoW nonce, hints) over the
same domain separator, and must print the same challenges and the same proof string.  No GPU involved: this is the
host-logic half of the end-to-end byte-identity that the GPU tests check."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "provekit_b200", "csrc", "host")
ORC = os.path.join(ROOT, "oracle")

SCRIPT = """
    // op codes: 0 add_scalars(n), 1 challenge_scalars(n), 2 challenge_bytes(n), 3 add_bytes(n), 4 hint(n bytes)
    static const int OPS[][2] = {{0,1},{1,1},{0,2},{1,1},{1,20},{0,1},{0,1},{1,1},{0,4},{1,1},{2,32},{3,8},{2,45},{4,0},{4,37},
                                 {0,3},{1,1},{2,1},{2,15},{2,16},{0,5},{1,3},{3,1},{1,2},{4,1000},{0,1},{1,1}};
"""

CPP = r"""
#include <cstdio>
#include "transcript.hpp"
using namespace pkh;
%s
static uint64_t lcg = 88172645463325252ULL;
static uint64_t next() { lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17; return lcg; }
static void hex(const uint8_t* p, size_t n) { for (size_t i = 0; i < n; i++) printf("%%02x", p[i]); printf("\n"); }
int main() {
    DomSep d("\xF0\x9F\x8C\xAA\xEF\xB8\x8F");
    d.absorb(1, "merkle_digest").squeeze(1, "ood_query").absorb(2, "ood_ans").squeeze(20, "rand").hint("stir_answers").absorb(8, "pow-nonce");
    hex((const uint8_t*)d.str().data(), d.str().size());
    ProverState ps(d.str());
    for (auto& op : OPS) {
        int n = op[1];
        if (op[0] == 0) { std::vector<Fr> x(n); for (auto& e : x) { uint64_t c[4] = {next(), next(), next(), next() >> 3}; e = from_canonical(c); } ps.add_scalars(x.data(), n); }
        if (op[0] == 1) { std::vector<Fr> x(n); ps.challenge_scalars(x.data(), n); for (auto& e : x) { uint64_t c[4]; to_canonical(e, c); hex((const uint8_t*)c, 32); } }
        if (op[0] == 2) { std::vector<uint8_t> b(n); ps.challenge_bytes(b.data(), n); hex(b.data(), n); }
        if (op[0] == 3) { std::vector<uint8_t> b(n); for (auto& e : b) e = (uint8_t)next(); ps.add_bytes(b.data(), n); }
        if (op[0] == 4) { std::vector<uint8_t> b(n); for (auto& e : b) e = (uint8_t)next(); ps.hint(b); }
    }
    hex(ps.narg().data(), ps.narg().size());
    return 0;
}
"""

C = r"""
#include <stdio.h>
#include <stdlib.h>
#include "transcript.h"
%s
static uint64_t lcg = 88172645463325252ULL;
static uint64_t next(void) { lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17; return lcg; }
static void hex(const uint8_t* p, size_t n) { for (size_t i = 0; i < n; i++) printf("%%02x", p[i]); printf("\n"); }
int main(void) {
    bytebuf d = {0};
    bb_str(&d, "\xF0\x9F\x8C\xAA\xEF\xB8\x8F");
    ds_op(&d, 'A', 1, "merkle_digest"); ds_op(&d, 'S', 1, "ood_query"); ds_op(&d, 'A', 2, "ood_ans"); ds_op(&d, 'S', 20, "rand");
    ds_op(&d, 'H', 0, "stir_answers"); ds_op(&d, 'A', 8, "pow-nonce");
    hex(d.p, d.len);
    fs_state fs;
    fs_init(&fs, d.p, d.len, NULL, 0);
    for (size_t k = 0; k < sizeof OPS / sizeof OPS[0]; k++) {
        int n = OPS[k][1];
        if (OPS[k][0] == 0) { fr_t* x = malloc(sizeof(fr_t) * (n ? n : 1)); for (int i = 0; i < n; i++) { uint64_t c[4]; c[0] = next(); c[1] = next(); c[2] = next(); c[3] = next() >> 3; x[i] = fr_from_canonical(c); } fs_add_scalars(&fs, x, n); free(x); }
        if (OPS[k][0] == 1) { fr_t* x = malloc(sizeof(fr_t) * (n ? n : 1)); fs_challenge_scalars(&fs, x, n); for (int i = 0; i < n; i++) { uint64_t c[4]; fr_to_canonical(x[i], c); hex((const uint8_t*)c, 32); } free(x); }
        if (OPS[k][0] == 2) { uint8_t* b = malloc(n ? n : 1); fs_challenge_bytes(&fs, b, n); hex(b, n); free(b); }
        if (OPS[k][0] == 3) { uint8_t* b = malloc(n ? n : 1); for (int i = 0; i < n; i++) b[i] = (uint8_t)next(); fs_add_bytes(&fs, b, n); free(b); }
        if (OPS[k][0] == 4) { uint8_t* b = malloc(n ? n : 1); for (int i = 0; i < n; i++) b[i] = (uint8_t)next(); fs_hint(&fs, b, n); free(b); }
    }
    hex(fs.narg.p, fs.narg.len);
    return 0;
}
"""


def test_host_transcript_equals_oracle_transcript(tmp_path):
    cpp, c = tmp_path / "host.cpp", tmp_path / "orc.c"
    cpp.write_text(CPP % SCRIPT)
    c.write_text(C % SCRIPT)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, str(cpp), os.path.join(HOST, "transcript.cpp"), "-o", str(tmp_path / "host")])
    subprocess.check_call(["gcc", "-O1", "-std=gnu11", "-march=x86-64-v3", "-I", ORC, str(c), os.path.join(ORC, "transcript.c"),
                           os.path.join(ORC, "skyscraper.c"), "-lm", "-o", str(tmp_path / "orc")])
    a = subprocess.run([str(tmp_path / "host")], capture_output=True, text=True, check=True).stdout.splitlines()
    b = subprocess.run([str(tmp_path / "orc")], capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(a) == len(b) and len(a) > 30
    assert a[0] == b[0], "domain separator strings differ"
    for i, (x, y) in enumerate(zip(a, b)):
        assert x == y, f"output line {i} differs"
    assert len(a[-1]) // 2 == 32 * (1 + 2 + 1 + 1 + 4 + 3 + 5 + 1) + 8 + 1 + (4 + 0) + (4 + 37) + (4 + 1000)   # scalars + bytes + framed hints
