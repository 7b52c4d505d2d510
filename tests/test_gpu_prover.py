"""End-to-end parity: pk_prove (GPU kernels + the library's host transcript) must emit, byte for byte,
the transcript the CPU oracle prover emits for the same witness, masks and R1CS, and the oracle
verifier (restating provekit/verifier + the Go recursive verifier) must accept it."""
import numpy as np
import pytest

from r1cs_util import SyntheticR1CS, oracle_prove, oracle_verify

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import provekit_b200 as pk
    c = pk.Context(0)
    yield c
    c.close()


def as_dict(r):
    return dict(num_constraints=r.nc, num_witnesses=r.nw, interned=r.interned, a=r.A, b=r.B, c=r.C)


def first_diff(a: bytes, b: bytes):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return n if len(a) != len(b) else -1


# sizes cover m mod 4 = 0..3 (final_sumcheck_rounds 0..3), 0..3 WHIR rounds, and a 2^17-coefficient witness polynomial
@pytest.mark.parametrize("nc,nfree", [(9, 4), (17, 5), (20, 30), (64, 40), (100, 300), (300, 400), (1000, 900), (5000, 3000),
                                      (5000, 9000), (1 << 15, 30000)])
def test_gpu_proof_is_byte_identical_to_oracle(ctx, orc, nc, nfree):
    import provekit_b200 as pk
    r = SyntheticR1CS(nc, nfree, seed=nc)
    expected = oracle_prove(orc, r)
    pr = pk.Prover(ctx, as_dict(r))
    got = pr.prove(r.witness, r.randomness())          # default: the transcript lives on the device
    assert len(got) == len(expected), (len(got), len(expected))
    assert first_diff(got, expected) == -1
    assert oracle_verify(orc, r, got) == 0
    assert pr.host_syncs <= 2, pr.host_syncs             # the host only waits for the finished proof string
    pr.set_host_transcript(True)                         # the same sponge on the host: a round trip per challenge
    assert pr.prove(r.witness, r.randomness()) == expected
    assert pr.host_syncs > 20
    pr.set_host_transcript(False)
    # proving twice with the same inputs is deterministic; other masks change the proof
    assert pr.prove(r.witness, r.randomness()) == got
    other = pr.prove(r.witness, r.randomness(seed=3))
    assert other != got and oracle_verify(orc, r, other) == 0
    pr.close()


def test_gpu_prover_rejects_malformed_r1cs(ctx):
    import provekit_b200 as pk
    r = SyntheticR1CS(64, 40, seed=9)
    d = as_dict(r)
    bad = dict(d)
    col = r.A[1].copy()
    col[3] = r.nw + 5  # column out of range
    bad["a"] = (r.A[0], col, r.A[2])
    with pytest.raises(pk.PkError):
        pk.Prover(ctx, bad)


def test_bad_witness_is_caught_by_verifier(ctx, orc):
    import provekit_b200 as pk
    r = SyntheticR1CS(64, 40, seed=11)
    w = r.witness.copy()
    w[45, 0] ^= np.uint64(1)
    pr = pk.Prover(ctx, as_dict(r))
    proof = pr.prove(w, r.randomness())
    assert oracle_verify(orc, r, proof) != 0
    pr.close()


def test_hot_column_long_rows(ctx, orc):
    """A witness column referenced by thousands of entries (constant-one / zero witness) makes one row of
    the transposed matrices very long: exercises the chunked long-row SpMV path.  Workload generator of
    bench.py at a small size, checked against the oracle prover and verifier."""
    import ctypes
    import provekit_b200 as pk
    from r1cs_util import CSRc, R1CSc, Randc
    from tools import workload as wl
    r = wl.synth_r1cs(5000, 6500, (5200, 4100, 21000), n_interned=40, seed=3)
    rnd = wl.randomness(r)

    def csr(t):
        return CSRc(r["num_constraints"], r["num_witnesses"], len(t[1]), t[0].ctypes.data, t[1].ctypes.data, t[2].ctypes.data)

    cs = R1CSc(r["num_constraints"], r["num_witnesses"], len(r["interned"]), r["interned"].ctypes.data,
               csr(r["a"]), csr(r["b"]), csr(r["c"]))
    rs = Randc(*[rnd[k].ctypes.data for k in ("mask_w", "g_w", "blind", "mask_h", "g_h")])
    out = ctypes.c_void_p()
    n = orc.orc_prove(ctypes.byref(cs), r["witness"].ctypes.data_as(ctypes.c_void_p), ctypes.byref(rs), 2, ctypes.byref(out))
    expected = ctypes.string_at(out, n)
    orc.orc_free(out)
    pr = pk.Prover(ctx, r)
    got = pr.prove(r["witness"], rnd)
    assert got == expected
    buf = np.frombuffer(got, dtype=np.uint8)
    assert orc.orc_verify(ctypes.byref(cs), buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(got)), 2) == 0
    pr.close()


def test_enqueue_collect_pipeline(ctx, orc):
    """asynchronous form: one host thread keeps several provers busy; every collected proof equals the oracle's"""
    import provekit_b200 as pk
    rs = [SyntheticR1CS(300 + 50 * i, 400, seed=70 + i) for i in range(3)]
    ctxs = [pk.Context(0) for _ in rs]
    provers = [pk.Prover(c, as_dict(r)) for c, r in zip(ctxs, rs)]
    expected = [oracle_prove(orc, r) for r in rs]
    for round_ in range(2):
        for p, r in zip(provers, rs):
            p.enqueue(r.witness, r.randomness())
        for p, e in zip(provers, expected):
            assert p.collect() == e
            assert p.host_syncs == 1
    # misuse: collect without enqueue, enqueue twice, and the host-transcript mode
    with pytest.raises(pk.PkError):
        provers[0].collect()
    provers[0].enqueue(rs[0].witness, rs[0].randomness())
    with pytest.raises(pk.PkError):
        provers[0].enqueue(rs[0].witness, rs[0].randomness())
    assert provers[0].collect() == expected[0]
    provers[1].set_host_transcript(True)
    with pytest.raises(pk.PkError):
        provers[1].enqueue(rs[1].witness, rs[1].randomness())
    for p in provers:
        p.close()
    for c in ctxs:
        c.close()
