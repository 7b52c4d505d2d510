"""Loader of tests/golden/poseidon-1000.r1cs.npz (written by tests/golden/make_r1cs_fixture.py from the reference's own
poseidon-1000.nps): the R1CS in the dict form provekit_b200.Prover takes."""
import os

import numpy as np

from helpers import arr_to_ints, to_mont

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poseidon-1000.r1cs.npz")


def load():
    z = np.load(PATH)
    out = dict(num_constraints=int(z["nc"]), num_witnesses=int(z["nw"]), interned=to_mont(arr_to_ints(z["interned_canonical"])))
    for k in "abc":
        ln = z[k + "_row_len"].astype(np.uint64)
        rs = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.uint64)
        col = np.cumsum(z[k + "_col_delta"].astype(np.int64)).astype(np.uint32)
        out[k] = (rs, col, z[k + "_val"].astype(np.uint32))
    return out
