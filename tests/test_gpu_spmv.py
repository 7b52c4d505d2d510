"""Kernel-level parity of the R1CS sparse mat-vec (k_spmv, k_spmv_long_chunks, k_spmv_long_rows) through the C-ABI
(pk_prover_matvec) against the oracle's restatement of provekit/common/src/sparse_matrix.rs:148-184, on
  * the R1CS of the reference's own scheme fixture poseidon-1000.nps (tests/golden/poseidon-1000.r1cs.npz): two 65 536-entry
    rows and hot columns of up to 598 479 uses, i.e. both the thread-per-row and the chunked long-row kernels, and
  * small synthetic matrices including empty rows and a row exactly at / one past the long-row threshold."""
import ctypes

import numpy as np
import pytest

from helpers import ptr, rand_fr
from r1cs_util import CSRc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import provekit_b200 as pk
    c = pk.Context(0)
    yield c
    c.close()


def orc_matvec(orc, r, which, x, transposed):
    rs, col, val = r["abc"[which]]
    m = CSRc(r["num_constraints"], r["num_witnesses"], len(col), rs.ctypes.data, col.ctypes.data, val.ctypes.data)
    out = np.zeros((r["num_witnesses"] if transposed else r["num_constraints"], 4), np.uint64)
    interned = np.ascontiguousarray(r["interned"])
    orc.orc_r1cs_matvec.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
    orc.orc_r1cs_matvec(ctypes.byref(m), ptr(interned), ptr(x), ptr(out), 1 if transposed else 0)
    return out


def check_all_products(ctx, orc, r, seed):
    import provekit_b200 as pk
    rng = np.random.default_rng(seed)
    z = rand_fr(rng, r["num_witnesses"])
    e = rand_fr(rng, r["num_constraints"])
    pr = pk.Prover(ctx, r)
    dz, de = ctx.upload(z), ctx.upload(e)
    for which in (0, 1):       # A z, B z  (calculate_witness_bounds, sumcheck.rs:181-193)
        got = pr.matvec(which, dz).download()
        assert np.array_equal(got, orc_matvec(orc, r, which, z, False)), f"{'AB'[which]} z"
    for which in (0, 1, 2):    # e^T A, e^T B, e^T C  (calculate_external_row_of_r1cs_matrices, sumcheck.rs:207-218)
        got = pr.matvec(which, de, transposed=True).download()
        assert np.array_equal(got, orc_matvec(orc, r, which, e, True)), f"e^T {'ABC'[which]}"
    with pytest.raises(pk.PkError):
        pr.matvec(2, dz)
    pr.close()


def test_reference_scheme_r1cs(ctx, orc):
    import r1cs_fixture
    r = r1cs_fixture.load()
    cnt = np.diff(np.append(r["a"][0].astype(np.int64), len(r["a"][1])))
    assert int(cnt.max()) == 65536 and int(np.bincount(r["c"][1]).max()) == 598_479  # the long-row / hot-column extremes
    check_all_products(ctx, orc, r, seed=1)


@pytest.mark.parametrize("row_len", [0, 1, 63, 64, 65, 2048, 2049, 5000])
def test_row_lengths_around_the_thresholds(ctx, orc, row_len):
    """one row of `row_len` entries between short rows and empty rows (SPMV_LONG_ROW = 64, SPMV_CHUNK = 2048)"""
    rng = np.random.default_rng(row_len)
    nc, nw, n_int = 37, 6000, 11
    lens = rng.integers(0, 4, size=nc)
    lens[5] = row_len
    lens[6] = 0
    lens[nc - 1] = 0
    rs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    nnz = int(lens.sum())
    col = np.concatenate([np.sort(rng.choice(nw, size=int(n), replace=False)) for n in lens] + [np.zeros(0, np.int64)]).astype(np.uint32)
    val = rng.integers(0, n_int, size=nnz, dtype=np.uint32)
    mat = (rs, col, val)
    r = dict(num_constraints=nc, num_witnesses=nw, interned=rand_fr(rng, n_int), a=mat, b=mat, c=mat)
    check_all_products(ctx, orc, r, seed=2)
