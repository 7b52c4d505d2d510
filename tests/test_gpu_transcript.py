"""pk_prove_with_transcript: the wholesale GPU prover with the HOST's Fiat-Shamir transcript (include/pkwhir.h,
pk_transcript_vtbl = the spongefish ProverState surface of provekit/prover/src/whir_r1cs.rs:240-242,268-272,335-337).

The reference's spongefish internals are not in the tree, so byte identity with the reference can only be
reached by letting the Rust host own the sponge.  What is checked here: driven through the callbacks, the GPU prover
emits exactly the messages the CPU oracle prover emits when driven by the same callbacks — for the oracle's own sponge
AND for a deliberately different toy sponge with its own codecs — so no prover message depends on the in-tree guess of
spongefish."""
import ctypes

import numpy as np
import pytest

from r1cs_util import SyntheticR1CS, oracle_prove
from transcripts import OracleTranscript, ToyTranscript, orc_prove_with, orc_verify_with

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import provekit_b200 as pk
    c = pk.Context(0)
    yield c
    c.close()


def as_dict(r):
    return dict(num_constraints=r.nc, num_witnesses=r.nw, interned=r.interned, a=r.A, b=r.B, c=r.C)


def first_log_diff(a, b):
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            return i, x[0], y[0]
    return (min(len(a), len(b)), "length", "length") if len(a) != len(b) else None


@pytest.mark.parametrize("nc,nfree", [(20, 30), (300, 400), (5000, 3000), (1 << 15, 30000)])
def test_oracle_sponge_through_callbacks(ctx, orc, nc, nfree):
    """(a) callbacks = the oracle's transcript: same bytes as orc_prove and as pk_prove's in-tree sponge."""
    import provekit_b200 as pk
    r = SyntheticR1CS(nc, nfree, seed=nc + 1)
    expected = oracle_prove(orc, r)
    pr = pk.Prover(ctx, as_dict(r))
    t = OracleTranscript(orc, r.nc, r.nw)
    pr.prove_with_transcript(r.witness, r.randomness(), t.vtbl_ptr, t.user)
    got = t.narg()
    t.close()
    assert got == expected
    assert pr.prove(r.witness, r.randomness()) == expected
    pr.close()


@pytest.mark.parametrize("nc,nfree", [(20, 30), (300, 400), (5000, 3000)])
def test_toy_sponge_through_callbacks(ctx, orc, nc, nfree):
    """(b) a different sponge with different codecs: GPU and oracle provers agree call by call, oracle verifier accepts."""
    import provekit_b200 as pk
    r = SyntheticR1CS(nc, nfree, seed=nc + 2)
    ref = ToyTranscript()
    assert orc_prove_with(orc, r, ctypes.byref(ref.orc_vtbl), None) == 0
    pr = pk.Prover(ctx, as_dict(r))
    t = ToyTranscript()
    pr.prove_with_transcript(r.witness, r.randomness(), t.pk_vtbl)
    pr.close()
    assert first_log_diff(t.log, ref.log) is None
    assert bytes(t.narg) == bytes(ref.narg)
    v = ToyTranscript(proof=bytes(t.narg))
    assert orc_verify_with(orc, r, ctypes.byref(v.orc_vtbl), None) == 0 and v.exhausted()


def test_full_size_m21_through_callbacks(ctx, orc):
    """poseidon-1000 shapes (m = 21, m_0 = 20) through the toy sponge: byte-identical to the oracle prover under the same
    callbacks."""
    import provekit_b200 as pk
    from r1cs_util import CSRc, R1CSc, Randc
    from tools import workload as wl
    r = wl.synth_r1cs(**wl.POSEIDON_1000, seed=2)
    rnd = wl.randomness(r, seed=5)
    assert wl.shapes(r)[:2] == (21, 20)

    def csr(t):
        return CSRc(r["num_constraints"], r["num_witnesses"], len(t[1]), t[0].ctypes.data, t[1].ctypes.data, t[2].ctypes.data)

    cs = R1CSc(r["num_constraints"], r["num_witnesses"], len(r["interned"]), r["interned"].ctypes.data, csr(r["a"]), csr(r["b"]),
               csr(r["c"]))
    rs = Randc(*[rnd[k].ctypes.data for k in ("mask_w", "g_w", "blind", "mask_h", "g_h")])
    ref = ToyTranscript()
    orc.orc_prove_with_transcript.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.c_void_p] * 2
    assert orc.orc_prove_with_transcript(ctypes.byref(cs), r["witness"].ctypes.data_as(ctypes.c_void_p), ctypes.byref(rs), 2,
                                         ctypes.byref(ref.orc_vtbl), None) == 0
    pr = pk.Prover(ctx, r)
    t = ToyTranscript()
    pr.prove_with_transcript(r["witness"], rnd, t.pk_vtbl)
    pr.close()
    assert first_log_diff(t.log, ref.log) is None
    assert bytes(t.narg) == bytes(ref.narg)


def test_callback_failure_is_reported(ctx):
    import provekit_b200 as pk
    from transcripts import CB_SCALARS
    r = SyntheticR1CS(64, 40, seed=3)
    pr = pk.Prover(ctx, as_dict(r))
    t = ToyTranscript()
    t._cbs[0] = CB_SCALARS(lambda u, p, n: 9)
    t.pk_vtbl.add_scalars = t._cbs[0]
    with pytest.raises(pk.PkError) as e:
        pr.prove_with_transcript(r.witness, r.randomness(), t.pk_vtbl)
    assert e.value.code == -1
    # incomplete table
    bad = type(t.pk_vtbl)()
    with pytest.raises(pk.PkError):
        pr.prove_with_transcript(r.witness, r.randomness(), bad)
    pr.close()
