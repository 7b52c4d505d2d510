"""Mask generator (ChaCha12 counter stream -> uniform Fr): block function pinned by the RFC 8439 known answer,
C oracle against the pure-Python root, range / determinism / stream separation.  CPU only.

Reference construction being mirrored: F::rand(&mut thread_rng()) — provekit/common/src/utils/zk_utils.rs:13-22,
provekit/prover/src/whir_r1cs.rs:211-225 (thread_rng = ChaCha12 stream; Fp::rand = rejection sampling)."""
import ctypes

import numpy as np

import oracle
from oracle import pyref

RFC_KEY = bytes(range(32))
# RFC 8439 section 2.3.2: counter = 1, nonce = 00:00:00:09:00:00:00:4a:00:00:00:00
RFC_STATE = pyref.CHACHA_CONST + [int.from_bytes(RFC_KEY[4 * i:4 * i + 4], "little") for i in range(8)] + \
    [1, 0x09000000, 0x4A000000, 0]
RFC_OUT = [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
           0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]


def c_block(state, rounds):
    L = oracle.lib()
    a = (ctypes.c_uint32 * 16)(*state)
    o = (ctypes.c_uint32 * 16)()
    L.orc_chacha_block(a, rounds, o)
    return list(o)


def c_fill(n, seed, stream):
    L = oracle.lib()
    out = np.zeros((n, 4), np.uint64)
    L.orc_rng_fill(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), seed, ctypes.c_uint32(stream))
    return out


def test_chacha20_rfc8439_block():
    assert pyref.chacha_block(RFC_STATE, 20) == RFC_OUT
    assert c_block(RFC_STATE, 20) == RFC_OUT


def test_chacha12_c_matches_python():
    rng = np.random.default_rng(3)
    for _ in range(20):
        st = [int(x) for x in rng.integers(0, 2**32, 16)]
        assert c_block(st, 12) == pyref.chacha_block(st, 12)


def test_rng_fill_matches_python_and_is_in_range():
    seed = bytes((7 * i + 1) & 0xFF for i in range(32))
    for stream in (0, 3):
        got = c_fill(300, seed, stream)
        for i in range(300):
            v = sum(int(got[i, k]) << (64 * k) for k in range(4))
            assert v < pyref.P
            assert v == pyref.rng_element(seed, stream, i)


def test_rng_streams_and_seeds_differ_and_repeat():
    s1, s2 = b"\x01" * 32, b"\x02" * 32
    a = c_fill(4096, s1, 0)
    assert np.array_equal(a, c_fill(4096, s1, 0))
    assert not np.array_equal(a, c_fill(4096, s1, 1))
    assert not np.array_equal(a, c_fill(4096, s2, 0))
    # prefix property: element i does not depend on n (counter based)
    assert np.array_equal(a[:100], c_fill(100, s1, 0))
    # rejection sampling really happens (p / 2^254 = 0.756) and the top limb stays below p's
    assert int(a[:, 3].max()) <= 0x30644E72E131A029
    assert len({tuple(r) for r in a.tolist()}) == 4096
