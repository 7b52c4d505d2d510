"""Pins layout conventions against the reference-produced proof (tests/golden/).
The fixture predates the v2 hash switch, so Merkle checks here use compress_v1."""
import ctypes

import numpy as np
import pytest

import fixture_walk as fw
from helpers import from_mont, ptr, to_mont
from oracle import pyref as o

P = o.P


@pytest.fixture(scope="module")
def proof():
    return fw.walk_proof()


def test_whir_config_matches_fixture_tables(proof):
    """the two WhirConfigs stored in poseidon-1000.nps."""
    cw, ch = proof["cfg_w"], proof["cfg_h"]
    assert cw["max_pow_bits"] == 19 and ch["max_pow_bits"] == 6
    assert [(r["pow_bits"], r["num_queries"], r["ood_samples"], r["log_inv_rate"], r["num_variables"], r["domain_log"])
            for r in cw["rounds"]] == [(19, 109, 1, 1, 17, 22), (16, 28, 1, 4, 13, 21), (16, 16, 1, 7, 9, 20), (18, 11, 1, 10, 5, 19)]
    assert (cw["final_queries"], cw["final_pow_bits"], cw["final_log_inv_rate"], cw["final_sumcheck_rounds"]) == (9, 11, 13, 1)
    assert [(r["pow_bits"], r["num_queries"], r["ood_samples"], r["log_inv_rate"], r["num_variables"], r["domain_log"])
            for r in ch["rounds"]] == [(6, 122, 1, 1, 4, 9)]
    assert (ch["final_queries"], ch["final_pow_bits"], ch["final_log_inv_rate"], ch["final_sumcheck_rounds"]) == (31, 4, 4, 0)


def test_transcript_layout(proof):
    """The op sequence of SURVEY A.4 lands exactly on the last byte (asserted inside walk_proof);
    spot-check the documented offsets via sizes."""
    w0 = proof["whir_w"]["rounds"][0]
    assert len(w0["answers"]) == 109 and all(len(a) == 32 for a in w0["answers"])
    assert len(proof["whir_w"]["rounds"][1]["answers"][0]) == 16
    assert len(proof["whir_w"]["final_coeffs"]) == 2 and len(proof["whir_w"]["deferred"]) == 3
    assert len(proof["whir_h"]["deferred"]) == 1
    assert len(proof["claimed_evaluations_raw"]) == 2 * (8 + 3 * 32)
    for rnd in proof["whir_w"]["rounds"]:
        idx = rnd["multipath"][3]
        assert idx == sorted(set(idx))  # sorted + de-duplicated


def _check_paths(root, answers, multipath, comp):
    sib, pre, suf, idx = multipath
    paths = fw.decode_paths(pre, suf)
    return sum(o.merkle_verify_path(root, i, leaf, s, p, comp) for i, leaf, s, p in zip(idx, answers, sib, paths)), len(idx)


def test_merkle_paths_hiding_whir(proof):
    h = proof["whir_h"]
    ok, n = _check_paths(proof["commit_h"]["root"], h["rounds"][0]["answers"], h["rounds"][0]["multipath"], o.compress_v1)
    assert (ok, n) == (32, 32)
    ok, n = _check_paths(h["rounds"][0]["root"], h["final_answers"], h["final_multipath"], o.compress_v1)
    assert (ok, n) == (13, 13)
    # and they do NOT verify with v2 (documents the stale fixture)
    ok2, _ = _check_paths(proof["commit_h"]["root"], h["rounds"][0]["answers"], h["rounds"][0]["multipath"], o.compress)
    assert ok2 == 0


def test_merkle_paths_witness_whir_sample(proof):
    w = proof["whir_w"]
    r0 = w["rounds"][0]
    sib, pre, suf, idx = r0["multipath"]
    paths = fw.decode_paths(pre, suf)
    assert len(paths[0]) == 17  # 2^18 leaves: depth 18 = leaf-sibling level + 17
    for j in (0, 1, 57, 108):
        assert o.merkle_verify_path(proof["commit_w"]["root"], idx[j], r0["answers"][j], sib[j], paths[j], o.compress_v1)
    r1 = w["rounds"][1]
    ok, n = _check_paths(r0["root"], r1["answers"], r1["multipath"], o.compress_v1)
    assert (ok, n) == (28, 28)


def test_merkle_c_oracle_on_fixture_leaves(proof, orc):
    """C oracle leaf hash + tree on reference data: rebuild the 32-leaf hiding tree bottom level from
    the opened leaves (all 32 leaves are opened) and compare the root (v1)."""
    h = proof["whir_h"]["rounds"][0]
    idx = h["multipath"][3]
    assert idx == list(range(32))
    flat = [x for leaf in h["answers"] for x in leaf]
    nodes = np.zeros((64, 4), np.uint64)
    orc.orc_merkle_build(ptr(to_mont(flat)), ctypes.c_size_t(32), ctypes.c_size_t(32), ptr(nodes), 1)
    assert from_mont(nodes)[1] == proof["commit_h"]["root"]


def test_rs_encode_kat_from_reference_proof(proof, orc):
    """the witness WHIR's last commit holds a 5-variable polynomial (32 coefficients,
    rate 2^-13, domain 2^18, 2^14 leaves x 16).  Every leaf entry k is c_k + c_{k+16}*Y at
    Y = (g^16)^i.  Recover the 32 coefficients from two opened leaves, RS-encode them with the
    oracle, and compare ALL 9 opened leaves with the reference's bytes."""
    w = proof["whir_w"]
    leaves = w["final_answers"]
    idx = w["final_multipath"][3]
    assert len(leaves) == 9 and all(len(l) == 16 for l in leaves)
    g = o.root_of_unity(18)
    g16 = pow(g, 16, P)
    y = [pow(g16, i, P) for i in idx]
    inv = pow((y[1] - y[0]) % P, -1, P)
    hi = [(leaves[1][k] - leaves[0][k]) * inv % P for k in range(16)]
    lo = [(leaves[0][k] - hi[k] * y[0]) % P for k in range(16)]
    coeffs = lo + hi
    out = np.zeros(((1 << 14) * 16, 4), np.uint64)
    orc.orc_rs_encode(ptr(to_mont(coeffs)), 5, 13, 4, ptr(out), ctypes.c_size_t(16), ctypes.c_size_t(0))
    for i, leaf in zip(idx, leaves):
        assert from_mont(out[i * 16:(i + 1) * 16]) == leaf
    # pyref agrees on those leaves as well
    for i, leaf in zip(idx[:3], leaves[:3]):
        z = pow(g16, i, P)
        assert [o.eval_univariate(coeffs[k::16], z) for k in range(16)] == leaf
    # the final coefficients sent in clear are this polynomial folded by the round's challenges:
    # consistency of the verifier's final check shape (2 coefficients, final_sumcheck_rounds = 1)
    assert len(w["final_coeffs"]) == 2
