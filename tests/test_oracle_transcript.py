"""Fiat-Shamir layer of the oracle (oracle/transcript.c): Keccak-f[1600] against hashlib's SHA3, duplex
discipline against a straight Python restatement, hint / scalar framing against the reference fixture layout.
The challenge VALUES of the real spongefish remain parity-unpinned; what is pinned here is that
the oracle's building blocks are what they claim to be."""
import ctypes
import hashlib

import numpy as np

from helpers import ptr
from oracle import pyref as o

P = o.P


def sha3_256_via_permutation(orc, msg: bytes) -> bytes:
    """SHA3-256 built around orc_keccak_f1600 (rate 136, pad 0x06..0x80)."""
    st = np.zeros(25, np.uint64)
    b = st.view(np.uint8)
    m = bytearray(msg) + b"\x06"
    m += b"\x00" * ((-len(m)) % 136)
    m[-1] |= 0x80
    for i in range(0, len(m), 136):
        b[:136] ^= np.frombuffer(bytes(m[i:i + 136]), np.uint8)
        orc.orc_keccak_f1600(ptr(st))
    return bytes(b[:32])


def test_keccak_permutation_matches_sha3(orc):
    for msg in (b"", b"abc", b"x" * 135, b"y" * 136, b"z" * 500):
        assert sha3_256_via_permutation(orc, msg) == hashlib.sha3_256(msg).digest()


def test_domsep_tag_is_unpadded_overwrite_duplex(orc):
    """tag = first 32 bytes after absorbing the domain separator in overwrite mode, no padding
    (duplex discipline mirrored by recursive-verifier/app/keccakSponge/keccakSponge.go:40-75)."""
    for io in (b"", b"\xF0\x9F\x8C\xAA\xEF\xB8\x8F\0A1merkle_digest", bytes(range(256)) * 2):
        st = np.zeros(25, np.uint64)
        b = st.view(np.uint8)
        ap = 0
        for ch in io:
            if ap == 136:
                orc.orc_keccak_f1600(ptr(st))
                ap = 0
            b[ap] = ch
            ap += 1
        orc.orc_keccak_f1600(ptr(st))
        exp = bytes(b[:32])
        tag = np.zeros(32, np.uint8)
        buf = np.frombuffer(io, np.uint8) if io else np.zeros(1, np.uint8)
        orc.orc_domsep_tag(ptr(buf), ctypes.c_size_t(len(io)), ptr(tag))
        assert tag.tobytes() == exp


def test_field_sponge_matches_python_restatement():
    """Skyscraper duplex (provekit/common/src/skyscraper/sponge.rs:24-58: state [0, Fr(iv)], rate 1) restated in
    Python with o.permute: absorb overwrites cell 0 (permuting first when full), squeeze permutes then reads."""
    iv = bytes(range(32))
    st = [0, int.from_bytes(iv, "little") % P]
    ap, sp = 0, 1
    out = []

    def absorb(x):
        nonlocal st, ap, sp
        if ap == 1:
            st = list(o.permute(*st))
            ap = 0
        st[0] = x % P
        ap, sp = 1, 1

    def squeeze():
        nonlocal st, ap, sp
        if sp == 1:
            sp, ap = 0, 0
            st = list(o.permute(*st))
        sp = 1
        return st[0]

    absorb(5)
    out.append(squeeze())
    out.append(squeeze())
    absorb(7)
    absorb(9)
    out.append(squeeze())
    # the same sequence through the oracle prover's transcript is exercised end to end by test_oracle_prover;
    # here we pin the discipline itself: two consecutive squeezes differ, absorb after squeeze does not permute first
    assert len(set(out)) == 3
    st2 = [0, int.from_bytes(iv, "little") % P]
    st2[0] = 5
    assert out[0] == o.permute(*st2)[0]
