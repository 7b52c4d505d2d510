"""Full-size cases of BASELINE.json: configs[1] (poseidon-rounds shapes, m = 21, m_0 = 20) and configs[3]
(synthetic R1CS with 2^22 constraints, m = 23, m_0 = 22).  At these sizes the CUDA path is checked through
size-independent properties — every opened Merkle path hashes to the root, every opened leaf is the polynomial's
evaluation on the right coset, the oracle verifier accepts the proof and rejects a tampered one — and, because the
C oracle prover still finishes in seconds, through byte identity of the whole proof as well."""
import ctypes

import numpy as np
import pytest

from helpers import arr_to_ints, from_mont, to_mont
from oracle import pyref as o
from r1cs_util import CSRc, R1CSc, Randc

pytestmark = pytest.mark.gpu
P = o.P


@pytest.fixture(scope="module")
def ctx():
    import provekit_b200 as pk
    c = pk.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("log_n", [21, 23])
def test_first_commitment_full_size(ctx, log_n):
    """commit_batch of two random 2^log_n-coefficient polynomials (rate 1/2, fold 16): L = 2^(log_n-3) leaves of 32."""
    n = 1 << log_n
    seed = bytes([log_n]) * 32
    polys = [ctx.buffer(n).rng_fill(seed, s) for s in (0, 1)]
    cm = ctx.commit_batch(polys, log_n, 1)
    L = 1 << (log_n + 1 - 4)
    assert (cm.num_leaves, cm.leaf_width) == (L, 32)
    rng = np.random.default_rng(log_n)
    idx = sorted({0, 1, L // 2, L - 1} | {int(i) for i in rng.integers(0, L, size=4)})
    leaves, sib, pre, sufs = cm.open(idx)
    root = from_mont(cm.root.reshape(1, 4))[0]
    w = o.root_of_unity(log_n + 1)
    prev = []
    sib_i, pre_i = arr_to_ints(sib), [int(x) for x in pre]
    for j, i in enumerate(idx):
        leaf = from_mont(leaves[j])
        # Merkle: MultiPath prefix decoding + leaf fold + inner compress up to the root (whir_utilities.go:13-46)
        prev = prev[:pre_i[j]] + (arr_to_ints(sufs[j]) if len(sufs[j]) else [])
        assert o.merkle_verify_path(root, i, leaf, sib_i[j], prev)
        # Reed-Solomon: leaf i holds f_k(w^(16 i)), so f(x) = sum_k x^k leaf[k] for every x with x^16 = w^(16 i)
        for t in (0, 11):
            x = pow(w, i + t * L, P)
            for b in (0, 1):
                fx = from_mont(ctx.eval_univariate(polys[b], n, to_mont([x])).reshape(1, 4))[0]
                assert fx == sum(pow(x, k, P) * leaf[16 * b + k] for k in range(16)) % P
    cm.free()
    for p in polys:
        p.free()


def _structs(r, masks):
    def csr(t):
        return CSRc(r["num_constraints"], r["num_witnesses"], len(t[1]), t[0].ctypes.data, t[1].ctypes.data, t[2].ctypes.data)

    cs = R1CSc(r["num_constraints"], r["num_witnesses"], len(r["interned"]), r["interned"].ctypes.data,
               csr(r["a"]), csr(r["b"]), csr(r["c"]))
    return cs, Randc(*[a.ctypes.data for a in masks])


@pytest.mark.parametrize("which", ["poseidon-1000", "synthetic-2^22"])
def test_full_size_proof(ctx, orc, which):
    import provekit_b200 as pk
    from tools import workload as wl
    if which == "poseidon-1000":
        r = wl.synth_r1cs(**wl.POSEIDON_1000, seed=1)
    else:
        nc = (1 << 22) - 4096
        r = wl.synth_r1cs(nc, 1 << 22, (nc + 4000, nc - 100_000, 2 * nc), n_interned=64, seed=4)
    m, m0, mh = wl.shapes(r)
    assert (m, m0) == ((21, 20) if which == "poseidon-1000" else (23, 22))
    seed = bytes((3 * i + m) & 0xFF for i in range(32))
    pr = pk.Prover(ctx, r)
    proof = pr.prove_seeded(r["witness"], seed)
    assert pr.host_syncs <= 25, pr.host_syncs  # device transcript: no per-challenge round trips (was ~120)
    pr.set_host_transcript(True)
    assert pr.prove_seeded(r["witness"], seed) == proof
    pr.close()
    masks = [np.zeros((k, 4), np.uint64) for k in (1 << (m - 1), 1 << m, 4 * m0, 1 << (mh - 1), 1 << mh)]
    orc.orc_rng_masks(seed, m, m0, mh, *[a.ctypes.data_as(ctypes.c_void_p) for a in masks])
    cs, rs = _structs(r, masks)
    buf = np.frombuffer(proof, dtype=np.uint8).copy()
    vp = ctypes.c_void_p
    assert orc.orc_verify(ctypes.byref(cs), buf.ctypes.data_as(vp), ctypes.c_size_t(len(buf)), 2) == 0
    bad = buf.copy()
    bad[len(bad) // 2] ^= 1
    assert orc.orc_verify(ctypes.byref(cs), bad.ctypes.data_as(vp), ctypes.c_size_t(len(bad)), 2) != 0
    out = vp()
    n = orc.orc_prove(ctypes.byref(cs), r["witness"].ctypes.data_as(vp), ctypes.byref(rs), 2, ctypes.byref(out))
    assert n > 0
    expected = ctypes.string_at(out, n)
    orc.orc_free(out)
    assert proof == expected
