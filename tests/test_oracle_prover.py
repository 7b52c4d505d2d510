"""Oracle prover -> oracle verifier round trip on synthetic satisfiable R1CS instances, plus checks
that the verifier rejects tampering.  The verifier restates provekit/verifier/src/whir_r1cs.rs and
the Go recursive verifier (recursive-verifier/app/circuit/*.go)."""
import pytest

import fixture_walk as fw
from oracle import pyref as o
from r1cs_util import SyntheticR1CS, oracle_prove, oracle_verify


# (nc, nfree) chosen so that m = ceil(log2(nc + nfree)) + 1 covers every residue mod 4 (final_sumcheck_rounds 0..3)
@pytest.mark.parametrize("nc,nfree", [(64, 40), (100, 300), (1000, 900), (17, 5), (20, 30), (300, 400), (9, 4)])
def test_prove_verify_roundtrip(orc, nc, nfree):
    r = SyntheticR1CS(nc, nfree, seed=nc)
    proof = oracle_prove(orc, r)
    assert oracle_verify(orc, r, proof) == 0
    # determinism: same inputs (witness, masks) -> byte-identical transcript
    assert oracle_prove(orc, r) == proof
    # different masks -> different proof, still accepted (zero knowledge masks are inputs)
    p2 = oracle_prove(orc, r, seed=5)
    assert p2 != proof and oracle_verify(orc, r, p2) == 0


def test_verifier_rejects_tampering(orc):
    r = SyntheticR1CS(64, 40, seed=1)
    proof = bytearray(oracle_prove(orc, r))
    assert oracle_verify(orc, r, bytes(proof)) == 0
    for pos in (5, 100, 200, 300, len(proof) // 2, len(proof) - 40):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert oracle_verify(orc, r, bytes(bad)) != 0, pos
    assert oracle_verify(orc, r, bytes(proof[:-1])) != 0


def test_verifier_rejects_bad_witness(orc):
    r = SyntheticR1CS(64, 40, seed=2)
    r.witness[50, 0] ^= 1  # break one product witness
    proof = oracle_prove(orc, r)
    assert oracle_verify(orc, r, proof) != 0


def test_transcript_layout_matches_reference_shape(orc):
    """The oracle's proof for an instance with the fixture's m/m_0 class must walk with the same
    layout walker that walks the reference-produced proof — here on a small instance."""
    r = SyntheticR1CS(100, 300, seed=3)
    proof = oracle_prove(orc, r)
    rd = fw.Reader(proof)
    cfg_w, cfg_h = o.whir_config(r.m), o.whir_config(r.mh)
    rd.scalars(3)
    rd.scalars(3)
    rd.scalar()
    for _ in range(r.m0):
        rd.scalars(4)
    rd.scalars(2)
    wh = fw.walk_whir(rd, cfg_h)
    ce = rd.hint()
    assert len(ce) == 2 * (8 + 3 * 32)
    ww = fw.walk_whir(rd, cfg_w)
    assert rd.pos == len(proof)
    assert len(ww["deferred"]) == 3 and len(wh["deferred"]) == 1
    for rnd in ww["rounds"]:
        idx = rnd["multipath"][3]
        assert idx == sorted(set(idx))
    # Merkle paths in the proof verify with the v2 hash
    first = ww["rounds"][0] if ww["rounds"] else None
    if first:
        sib, pre, suf, idx = first["multipath"]
        paths = fw.decode_paths(pre, suf)
        root = int.from_bytes(proof[:32], "little")
        for j in range(min(5, len(idx))):
            assert o.merkle_verify_path(root, idx[j], first["answers"][j], sib[j], paths[j])
