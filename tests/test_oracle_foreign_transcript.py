"""The oracle prover / verifier driven by a foreign Fiat-Shamir transcript (orc_prove_with_transcript): the CPU half of
the transcript-agnostic wholesale entry point (pk_prove_with_transcript, include/pkwhir.h).  GPU half:
tests/test_gpu_transcript.py."""
import ctypes

from r1cs_util import SyntheticR1CS, oracle_prove, oracle_verify
from transcripts import OracleTranscript, ToyTranscript, orc_prove_with, orc_verify_with


def test_oracle_sponge_behind_the_vtable_reproduces_orc_prove(orc):
    r = SyntheticR1CS(300, 200, seed=11)
    exp = oracle_prove(orc, r)
    t = OracleTranscript(orc, r.nc, r.nw)
    assert orc_prove_with(orc, r, t.vtbl_ptr, t.user) == 0
    assert t.narg() == exp
    t.close()
    # and the verifier through the same table
    v = OracleTranscript(orc, r.nc, r.nw, proof=exp)
    assert orc_verify_with(orc, r, v.vtbl_ptr, v.user) == 0
    v.close()


def test_toy_transcript_round_trip_and_tamper(orc):
    r = SyntheticR1CS(300, 200, seed=12)
    t = ToyTranscript()
    assert orc_prove_with(orc, r, ctypes.byref(t.orc_vtbl), None) == 0
    proof = bytes(t.narg)
    ops = [op for op, _ in t.log]
    assert ops.count("hint") >= 5 and ops.count("challenge_bytes") >= 4 and ops.count("add_bytes") >= 2
    # a different sponge gives different challenges, hence a different proof than the in-tree one
    assert proof != oracle_prove(orc, r)
    v = ToyTranscript(proof=proof)
    assert orc_verify_with(orc, r, ctypes.byref(v.orc_vtbl), None) == 0 and v.exhausted()
    # flip one bit of a sumcheck scalar: some check (or the canonical-scalar parse) must fail
    bad = bytearray(proof)
    bad[32 * 5 + 31] ^= 1
    v = ToyTranscript(proof=bytes(bad))
    assert orc_verify_with(orc, r, ctypes.byref(v.orc_vtbl), None) != 0
    # the in-tree verifier cannot accept a proof made under another sponge
    assert oracle_verify(orc, r, proof) != 0


def test_callback_failure_aborts(orc):
    r = SyntheticR1CS(100, 80, seed=13)
    t = ToyTranscript()
    calls = {"n": 0}
    orig = t._challenge_scalars

    def failing(u, p, n):
        calls["n"] += 1
        return 7 if calls["n"] == 3 else orig(u, p, n)

    from transcripts import CB_SCALARS
    t._cbs[1] = CB_SCALARS(failing)
    t.orc_vtbl.challenge_scalars = t._cbs[1]
    assert orc_prove_with(orc, r, ctypes.byref(t.orc_vtbl), None) != 0
