"""Device mask generator (pk_rng_fill, pk_prove_seeded) against its CPU twin (oracle/rng.c, pinned by the RFC 8439
block-function vector in test_oracle_rng.py): bit-exact elements, and a seeded proof equals the oracle proof over the
oracle-generated masks."""
import ctypes

import numpy as np
import pytest

from r1cs_util import Randc, SyntheticR1CS, oracle_verify

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import provekit_b200 as pk
    c = pk.Context(0)
    yield c
    c.close()


def orc_fill(orc, n, seed, stream):
    out = np.zeros((n, 4), np.uint64)
    orc.orc_rng_fill(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), seed, ctypes.c_uint32(stream))
    return out


@pytest.mark.parametrize("n", [1, 31, 1000, (1 << 16) + 3, 1 << 20])
def test_rng_fill_matches_oracle(ctx, orc, n):
    seed = bytes((11 * i + n) & 0xFF for i in range(32))
    for stream in (0, 4):
        buf = ctx.buffer(n + 5).zero()
        buf.rng_fill(seed, stream, off=5, n=n)
        got = buf.download()
        assert not got[:5].any()  # the offset is honoured
        assert np.array_equal(got[5:], orc_fill(orc, n, seed, stream))
        buf.free()


def test_rng_fill_empty_and_bounds(ctx):
    import provekit_b200 as pk
    buf = ctx.buffer(8).zero()
    buf.rng_fill(b"\0" * 32, 0, off=8, n=0)
    assert not buf.download().any()
    with pytest.raises(pk.PkError):
        buf.rng_fill(b"\0" * 32, 0, off=4, n=5)
    buf.free()


@pytest.mark.parametrize("nc,nfree", [(20, 30), (1000, 900), (5000, 9000)])
def test_seeded_proof_equals_oracle_proof_over_oracle_masks(ctx, orc, nc, nfree):
    import provekit_b200 as pk
    r = SyntheticR1CS(nc, nfree, seed=nc + 1)
    seed = bytes((5 * i + nc) & 0xFF for i in range(32))
    masks = [np.zeros((n, 4), np.uint64) for n in (1 << (r.m - 1), 1 << r.m, 4 * r.m0, 1 << (r.mh - 1), 1 << r.mh)]
    orc.orc_rng_masks(seed, r.m, r.m0, r.mh, *[a.ctypes.data_as(ctypes.c_void_p) for a in masks])
    cs = r.c_struct()
    rs = Randc(*[a.ctypes.data for a in masks])
    out = ctypes.c_void_p()
    n = orc.orc_prove(ctypes.byref(cs), r.witness.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rs), 2, ctypes.byref(out))
    assert n > 0
    expected = ctypes.string_at(out, n)
    orc.orc_free(out)
    pr = pk.Prover(ctx, dict(num_constraints=r.nc, num_witnesses=r.nw, interned=r.interned, a=r.A, b=r.B, c=r.C))
    got = pr.prove_seeded(r.witness, seed)
    assert got == expected
    assert oracle_verify(orc, r, got) == 0
    # explicit-mask entry point over the same masks gives the same bytes; another seed another proof
    assert pr.prove(r.witness, dict(zip(("mask_w", "g_w", "blind", "mask_h", "g_h"), masks))) == got
    other = pr.prove_seeded(r.witness, bytes(32))
    assert other != got and oracle_verify(orc, r, other) == 0
    pr.close()
