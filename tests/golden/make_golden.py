"""Regenerates tests/golden/* from the reference's checked-in fixtures.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Source: tooling/provekit-bench/benches/poseidon-1000.np (a reference-produced proof,
container format provekit/common/src/file/bin.rs:16-38: 8 B magic, 8 B format tag, u16 major,
u16 minor, then zstd(postcard(NoirProof))).  The postcard payload is varint(len) + raw
spongefish transcript.  The transcript is stored verbatim; tests walk it.
"""
import ctypes
import hashlib
import os
import sys

REF = "/root/reference/tooling/provekit-bench/benches"
HERE = os.path.dirname(os.path.abspath(__file__))


def zstd_decompress(data: bytes, cap: int = 1 << 28) -> bytes:
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_createDStream.restype = ctypes.c_void_p
    z.ZSTD_initDStream.argtypes = [ctypes.c_void_p]
    z.ZSTD_initDStream.restype = ctypes.c_size_t

    class Buf(ctypes.Structure):
        _fields_ = [("p", ctypes.c_void_p), ("size", ctypes.c_size_t), ("pos", ctypes.c_size_t)]

    z.ZSTD_decompressStream.argtypes = [ctypes.c_void_p, ctypes.POINTER(Buf), ctypes.POINTER(Buf)]
    z.ZSTD_decompressStream.restype = ctypes.c_size_t
    z.ZSTD_isError.argtypes = [ctypes.c_size_t]
    ds = z.ZSTD_createDStream()
    z.ZSTD_initDStream(ds)
    src = ctypes.create_string_buffer(data, len(data))
    dst = ctypes.create_string_buffer(cap)
    ib = Buf(ctypes.cast(src, ctypes.c_void_p), len(data), 0)
    ob = Buf(ctypes.cast(dst, ctypes.c_void_p), cap, 0)
    while ib.pos < ib.size:
        r = z.ZSTD_decompressStream(ds, ctypes.byref(ob), ctypes.byref(ib))
        if z.ZSTD_isError(r):
            raise RuntimeError("zstd error")
        if r == 0:
            break
    return dst.raw[:ob.pos]


def varint(b: bytes, pos: int):
    v = s = 0
    while True:
        c = b[pos]
        pos += 1
        v |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return v, pos


def main():
    raw = open(os.path.join(REF, "poseidon-1000.np"), "rb").read()
    assert raw[:8] == b"\xDC\xDFOZkp\x01\x00", raw[:8]
    assert raw[8:16] == b"NPSProof", raw[8:16]
    payload = zstd_decompress(raw[20:])
    n, pos = varint(payload, 0)
    transcript = payload[pos:pos + n]
    assert len(transcript) == n == 268756, (n, len(payload))
    out = os.path.join(HERE, "poseidon-1000.transcript.bin")
    open(out, "wb").write(transcript)
    print("wrote", out, len(transcript), hashlib.sha256(transcript).hexdigest())
    # the container itself, verbatim (258 KB): exercises pk_np_decode on bytes the reference wrote
    out2 = os.path.join(HERE, "poseidon-1000.np")
    open(out2, "wb").write(raw)
    print("wrote", out2, len(raw), hashlib.sha256(raw).hexdigest())


if __name__ == "__main__":
    sys.exit(main())
