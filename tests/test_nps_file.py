"""`.nps` scheme container (row f3): pk_nps_read_r1cs must recover the R1CS of a NoirProofScheme file.

* a synthetic container written here in the reference's wire format (provekit/common/src/file/bin.rs:16-60,
  r1cs.rs:7-13, interner.rs:6-10 + utils/serde_ark.rs, sparse_matrix.rs:10-27; postcard varints), with opaque bytes
  standing in for the ACIR program in front and the witness builders behind;
* the reference's own fixture tooling/provekit-bench/benches/poseidon-1000.nps when /root/reference is mounted
  (this container only; 4.6 MB compressed, 130 MB inflated): shape and nnz of SURVEY fact 9."""
import ctypes
import os

import numpy as np
import pytest

import provekit_b200 as pk
from helpers import P, arr_to_ints, from_mont
from r1cs_util import SyntheticR1CS

MAGIC = b"\xDC\xDFOZkp\x01\x00"
REF_NPS = "/root/reference/tooling/provekit-bench/benches/poseidon-1000.nps"


def zstd_compress(data: bytes) -> bytes:
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_compressBound.restype = ctypes.c_size_t
    z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    z.ZSTD_compress.restype = ctypes.c_size_t
    z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    cap = z.ZSTD_compressBound(len(data))
    dst = ctypes.create_string_buffer(cap)
    n = z.ZSTD_compress(dst, cap, data, len(data), 3)
    return dst.raw[:n]


def varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | 0x80 if v else b)
        if not v:
            return bytes(out)


def vec(xs) -> bytes:
    return varint(len(xs)) + b"".join(varint(int(x)) for x in xs)


def postcard_r1cs(r: SyntheticR1CS, num_public_inputs=0) -> bytes:
    canon = from_mont(r.interned)
    blob = len(canon).to_bytes(8, "little") + b"".join(v.to_bytes(32, "little") for v in canon)  # ark Vec<Fr>
    out = varint(num_public_inputs) + varint(len(blob)) + blob
    for rs, col, val in (r.A, r.B, r.C):
        out += varint(r.nc) + varint(r.nw) + vec(rs) + vec(col) + vec(val)
    return out


def nps_file(payload: bytes, fmt=b"NrProScm") -> bytes:
    return MAGIC + fmt + b"\0\0\0\0" + zstd_compress(payload)


def check_equal(got: dict, r: SyntheticR1CS):
    assert got["num_constraints"] == r.nc and got["num_witnesses"] == r.nw
    assert np.array_equal(got["interned"], r.interned)  # Montgomery form, as Prover takes it
    for k, t in (("a", r.A), ("b", r.B), ("c", r.C)):
        for x, y in zip(got[k], t):
            assert np.array_equal(x, y), k


@pytest.mark.parametrize("nc,nfree,n_interned", [(200, 150, 17), (3000, 2000, 300)])
def test_synthetic_scheme_round_trip(nc, nfree, n_interned):
    r = SyntheticR1CS(nc, nfree, seed=nc, n_interned=n_interned)
    rng = np.random.default_rng(nc)
    program = rng.integers(0, 256, size=50_000, dtype=np.uint8).tobytes()  # opaque stand-in for the ACIR Program
    builders = rng.integers(0, 256, size=10_000, dtype=np.uint8).tobytes()
    got = pk.nps_read_r1cs(nps_file(program + postcard_r1cs(r) + builders))
    check_equal(got, r)
    assert got["num_public_inputs"] in (0, -1)


def test_decoy_interner_is_skipped():
    """a byte string that looks like an interner (length 8 + 32c, count c, canonical elements) but is not followed by
    three well-formed matrices must not stop the scan"""
    r = SyntheticR1CS(100, 80, seed=3, n_interned=40)
    decoy = (40).to_bytes(8, "little") + b"".join((7 + i).to_bytes(32, "little") for i in range(40))
    payload = b"\x01" * 100 + varint(len(decoy)) + decoy + b"\xff" * 64 + b"\x00" + postcard_r1cs(r)
    check_equal(pk.nps_read_r1cs(nps_file(payload)), r)


def test_rejects_bad_containers():
    r = SyntheticR1CS(100, 80, seed=3, n_interned=40)
    good = nps_file(b"\x00" * 10 + postcard_r1cs(r))
    pk.nps_read_r1cs(good)
    bad_matrix = postcard_r1cs(r)[:-5]  # truncated value vector
    non_canonical = (P + 1).to_bytes(32, "little")
    for bad in (b"", good[:19], b"X" + good[1:], nps_file(b"\x00" * 10 + postcard_r1cs(r), fmt=b"NPSProof"), good[:20] + b"junk",
                nps_file(b"\x00" * 10 + bad_matrix), nps_file(varint(8 + 32 * 40) + (40).to_bytes(8, "little") + non_canonical * 40)):
        with pytest.raises(pk.PkError):
            pk.nps_read_r1cs(bad)


@pytest.mark.skipif(not os.path.exists(REF_NPS), reason="reference fixture only exists in the build container")
def test_reference_fixture_r1cs():
    got = pk.nps_read_r1cs(open(REF_NPS, "rb").read())
    assert (got["num_constraints"], got["num_witnesses"]) == (729_560, 860_637)
    assert [len(got[k][1]) for k in "abc"] == [740_508, 609_440, 1_915_568]
    assert got["interned"].shape == (366, 4) and got["num_public_inputs"] == 0
    assert all(v < P for v in arr_to_ints(got["interned"]))
    # the workload bench.py synthesises has exactly these shapes (tools/workload.py POSEIDON_1000)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools.workload import POSEIDON_1000 as W
    assert (W["num_constraints"], W["num_witnesses"]) == (got["num_constraints"], got["num_witnesses"])
    assert tuple(W["nnz"]) == tuple(len(got[k][1]) for k in "abc") and W["n_interned"] == 366
    for k in "abc":
        rs, col, val = got[k]
        assert rs[0] == 0 and np.all(np.diff(rs.astype(np.int64)) >= 0) and int(col.max()) < got["num_witnesses"] and int(val.max()) < 366


def _extremes(r):
    out = {}
    for k in "abc":
        rs, col, _ = r[k]
        cnt = np.diff(np.append(rs.astype(np.int64), len(col)))
        cc = np.bincount(col, minlength=r["num_witnesses"])
        out[k] = dict(nnz=len(col), row_max=int(cnt.max()), long_rows=int((cnt > 64).sum()), empty=int((cnt == 0).sum()),
                      col_max=int(cc.max()), hot_cols=int((cc > 64).sum()))
    return out


def test_synthetic_workload_has_the_fixture_sparsity_extremes():
    """bench.py's synthetic poseidon-1000 R1CS (tools/workload.py) is satisfiable and has the long rows / hot columns /
    empty rows of the real scheme; in the build container it is compared with the fixture itself."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools import workload as wl
    r = wl.synth_r1cs(**wl.POSEIDON_1000, seed=1)
    ex = _extremes(r)
    assert ex["a"]["long_rows"] == 2 and 65_536 <= ex["a"]["row_max"] <= 65_540 and ex["a"]["hot_cols"] == 2
    assert ex["b"]["row_max"] == 1 and ex["b"]["hot_cols"] == 1 and 131_076 <= ex["b"]["col_max"] <= 131_090
    assert ex["c"]["hot_cols"] == 1 and 598_479 <= ex["c"]["col_max"] <= 598_500
    # satisfiable: (A z) * (B z) == C z on the long rows and a sample of the others
    rinv = pow(1 << 256, P - 2, P)
    z = [v * rinv % P for v in arr_to_ints(r["witness"])]
    iv = [v * rinv % P for v in arr_to_ints(r["interned"])]

    def row_dot(m, i):
        rs, col, val = m
        s, e = int(rs[i]), int(rs[i + 1]) if i + 1 < len(rs) else len(col)
        return sum(iv[val[k]] * z[col[k]] for k in range(s, e)) % P

    cnt_a = np.diff(np.append(r["a"][0].astype(np.int64), len(r["a"][1])))
    rows = [int(i) for i in np.argsort(cnt_a)[-2:]] + list(range(0, r["num_constraints"], 9973))
    assert all(row_dot(r["a"], i) * row_dot(r["b"], i) % P == row_dot(r["c"], i) for i in rows)
    if os.path.exists(REF_NPS):
        real = _extremes(pk.nps_read_r1cs(open(REF_NPS, "rb").read()))
        for k in "abc":
            assert ex[k]["nnz"] == real[k]["nnz"] and ex[k]["long_rows"] == real[k]["long_rows"] and ex[k]["hot_cols"] <= real[k]["hot_cols"]
            assert abs(ex[k]["col_max"] - real[k]["col_max"]) <= 16
            assert abs(ex[k]["row_max"] - real[k]["row_max"]) <= (4 if real[k]["long_rows"] else 16)
            assert abs(ex[k]["empty"] - real[k]["empty"]) <= 16


def test_truncated_frame_and_ambiguous_scheme_are_errors():
    """(1) a zstd frame cut short is an error even when the cut falls behind the R1CS; (2) two well-formed R1CS
    candidates in one stream make the locate-by-shape heuristic ambiguous: refuse instead of returning the first"""
    r = SyntheticR1CS(100, 80, seed=3, n_interned=40)
    rng = np.random.default_rng(5)
    tail = rng.integers(0, 256, size=200_000, dtype=np.uint8).tobytes()  # incompressible: the frame's last blocks
    good = nps_file(b"\x00" * 10 + postcard_r1cs(r) + tail)
    check_equal(pk.nps_read_r1cs(good), r)
    with pytest.raises(pk.PkError):
        pk.nps_read_r1cs(good[:-1000])
    r2 = SyntheticR1CS(120, 90, seed=4, n_interned=40)
    with pytest.raises(pk.PkError):
        pk.nps_read_r1cs(nps_file(b"\x00" * 10 + postcard_r1cs(r) + b"\x00" * 7 + postcard_r1cs(r2)))
