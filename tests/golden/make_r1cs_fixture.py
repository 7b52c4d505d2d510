"""Writes tests/golden/poseidon-1000.r1cs.npz: the R1CS (interner + three interned-CSR matrices) of the reference's own
scheme fixture tooling/provekit-bench/benches/poseidon-1000.nps, as decoded by pk_nps_read_r1cs.  The .nps (16.6 MB) cannot
travel to the GPU box and is too large for a fixture; its R1CS alone, delta-coded, is what the kernel-level SpMV parity test
(tests/test_gpu_spmv.py) needs.  Run in the build container (reads /root/reference):
    python tests/golden/make_r1cs_fixture.py
Layout: per matrix row_len (uint32 per-row entry counts), col_delta (int32: column minus the previous entry's column, wrapping
to the first column at row starts is NOT special-cased — plain running difference over the whole array), val (uint16);
interned (366 x 4 uint64, canonical form); nc, nw."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/tooling/provekit-bench/benches/poseidon-1000.nps"


def main():
    import provekit_b200 as pk
    from helpers import from_mont, ints_to_arr
    r = pk.nps_read_r1cs(open(REF, "rb").read())
    out = dict(nc=np.int64(r["num_constraints"]), nw=np.int64(r["num_witnesses"]),
               interned_canonical=ints_to_arr(from_mont(r["interned"])))
    for k in "abc":
        rs, col, val = r[k]
        ext = np.append(rs.astype(np.int64), len(col))
        out[k + "_row_len"] = np.diff(ext).astype(np.uint32)
        out[k + "_col_delta"] = np.diff(col.astype(np.int64), prepend=0).astype(np.int32)
        assert int(val.max()) < 65536
        out[k + "_val"] = val.astype(np.uint16)
    path = os.path.join(ROOT, "tests", "golden", "poseidon-1000.r1cs.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
