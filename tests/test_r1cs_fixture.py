"""tests/golden/poseidon-1000.r1cs.npz decodes to the R1CS of the reference's poseidon-1000.nps, and — in
the build container, where the .nps itself is readable — to exactly what pk_nps_read_r1cs extracts from it."""
import os

import numpy as np
import pytest

import r1cs_fixture

REF_NPS = "/root/reference/tooling/provekit-bench/benches/poseidon-1000.nps"


def test_fixture_shapes():
    r = r1cs_fixture.load()
    assert (r["num_constraints"], r["num_witnesses"]) == (729_560, 860_637)
    assert [len(r[k][1]) for k in "abc"] == [740_508, 609_440, 1_915_568]
    assert r["interned"].shape == (366, 4)
    for k in "abc":
        rs, col, val = r[k]
        assert len(rs) == r["num_constraints"] and rs[0] == 0 and np.all(np.diff(rs.astype(np.int64)) >= 0)
        assert int(col.max()) < r["num_witnesses"] and int(val.max()) < 366


@pytest.mark.skipif(not os.path.exists(REF_NPS), reason="reference fixture only exists in the build container")
def test_fixture_equals_the_reference_scheme():
    import provekit_b200 as pk
    got = pk.nps_read_r1cs(open(REF_NPS, "rb").read())
    r = r1cs_fixture.load()
    assert np.array_equal(got["interned"], r["interned"])
    for k in "abc":
        for x, y in zip(got[k], r[k]):
            assert np.array_equal(x, y), k
