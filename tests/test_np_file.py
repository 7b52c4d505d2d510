"""`.np` proof container (provekit/common/src/file/bin.rs): decode the file the reference wrote, and round-trip ours."""
import hashlib
import os

import provekit_b200 as pk

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_decode_reference_np_file():
    raw = open(os.path.join(GOLD, "poseidon-1000.np"), "rb").read()
    assert hashlib.sha256(raw).hexdigest() == "ea5f39db9733ad03353a424d19fc1014751a882675e1fad1cc90d451507ac028"
    transcript = pk.np_decode(raw)
    assert transcript == open(os.path.join(GOLD, "poseidon-1000.transcript.bin"), "rb").read()


def test_encode_roundtrip_and_header():
    t = open(os.path.join(GOLD, "poseidon-1000.transcript.bin"), "rb").read()
    for payload in (t, b"", b"\x01", bytes(range(256)) * 3):
        f = pk.np_encode(payload)
        assert f[:8] == b"\xDC\xDFOZkp\x01\x00" and f[8:16] == b"NPSProof" and f[16:20] == b"\0\0\0\0"  # bin.rs:16-18, mod.rs:33-37
        assert pk.np_decode(f) == payload
    assert len(pk.np_encode(t)) < len(t)  # actually compressed


def test_decode_rejects_bad_containers():
    import pytest
    good = pk.np_encode(b"hello")
    for bad in (b"", good[:19], b"X" + good[1:], good[:8] + b"NrProScm" + good[16:], good[:20] + b"garbage"):
        with pytest.raises(pk.PkError):
            pk.np_decode(bad)


def test_decode_rejects_truncated_and_oversized_frames():
    """a zstd frame cut short must be an error, not a silently shorter payload (the input running out with the decoder
    still expecting data used to be accepted)"""
    import pytest
    t = open(os.path.join(GOLD, "poseidon-1000.transcript.bin"), "rb").read()
    good = pk.np_encode(t)
    for cut in (len(good) - 1, len(good) - 100, len(good) // 2, 40):
        with pytest.raises(pk.PkError):
            pk.np_decode(good[:cut])
