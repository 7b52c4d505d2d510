"""Synthetic satisfiable R1CS in the reference's interned-CSR form:
row i:  A_i = {(a1, alpha), (a2, beta)},  B_i likewise,  C_i = {(K + i, 1)} with a dedicated product
witness, so any random assignment of the K free witnesses extends to a satisfying one."""
import ctypes

import numpy as np

from helpers import from_mont, ptr, rand_fr, to_mont
from oracle import pyref as o

P = o.P


class CSRc(ctypes.Structure):
    _fields_ = [("num_rows", ctypes.c_uint64), ("num_cols", ctypes.c_uint64), ("nnz", ctypes.c_uint64),
                ("row_start", ctypes.c_void_p), ("col", ctypes.c_void_p), ("val", ctypes.c_void_p)]


class R1CSc(ctypes.Structure):
    _fields_ = [("num_constraints", ctypes.c_uint64), ("num_witnesses", ctypes.c_uint64),
                ("num_interned", ctypes.c_uint64), ("interned", ctypes.c_void_p),
                ("a", CSRc), ("b", CSRc), ("c", CSRc)]


class Randc(ctypes.Structure):
    _fields_ = [("mask_w", ctypes.c_void_p), ("g_w", ctypes.c_void_p), ("blind", ctypes.c_void_p),
                ("mask_h", ctypes.c_void_p), ("g_h", ctypes.c_void_p)]


def log2ceil(n):
    return max(0, (n - 1).bit_length())


class SyntheticR1CS:
    def __init__(self, num_constraints: int, num_free: int, seed: int = 4, n_interned: int = 17):
        rng = np.random.default_rng(seed)
        self.nc, self.K = num_constraints, num_free
        self.nw = num_free + num_constraints
        self.interned = rand_fr(rng, n_interned)           # Montgomery-form constants
        self.interned[0] = to_mont([1])[0]
        nc = num_constraints

        def two_term():
            cols = rng.integers(0, num_free, size=(nc, 2), dtype=np.uint32)
            cols[:, 1] = np.where(cols[:, 1] == cols[:, 0], (cols[:, 0] + 1) % num_free, cols[:, 1])
            cols.sort(axis=1)
            vals = rng.integers(0, n_interned, size=(nc, 2), dtype=np.uint32)
            return (np.arange(nc, dtype=np.uint64) * 2, np.ascontiguousarray(cols.reshape(-1)),
                    np.ascontiguousarray(vals.reshape(-1)))

        self.A = two_term()
        self.B = two_term()
        self.C = (np.arange(nc, dtype=np.uint64), (num_free + np.arange(nc)).astype(np.uint32),
                  np.zeros(nc, np.uint32))
        # witness
        free = rand_fr(rng, num_free)
        free[0] = to_mont([1])[0]
        fz = from_mont(free)
        iv = from_mont(self.interned)
        prod = []
        for i in range(nc):
            az = sum(iv[self.A[2][2 * i + t]] * fz[self.A[1][2 * i + t]] for t in range(2)) % P
            bz = sum(iv[self.B[2][2 * i + t]] * fz[self.B[1][2 * i + t]] for t in range(2)) % P
            prod.append(az * bz % P)
        self.witness = np.concatenate([free, to_mont(prod)])
        self.m = log2ceil(self.nw) + 1
        self.m0 = log2ceil(nc)
        self.mh = log2ceil(4 * self.m0) + 1

    def _csr(self, t, cls):
        rs, col, val = t
        return cls(self.nc, self.nw, len(col), rs.ctypes.data, col.ctypes.data, val.ctypes.data)

    def c_struct(self, csr_cls=CSRc, r1cs_cls=R1CSc):
        return r1cs_cls(self.nc, self.nw, len(self.interned), self.interned.ctypes.data,
                        self._csr(self.A, csr_cls), self._csr(self.B, csr_cls), self._csr(self.C, csr_cls))

    def randomness(self, seed=99):
        rng = np.random.default_rng(seed)
        self._rand = dict(mask_w=rand_fr(rng, 1 << (self.m - 1)), g_w=rand_fr(rng, 1 << self.m),
                          blind=rand_fr(rng, 4 * self.m0), mask_h=rand_fr(rng, 1 << (self.mh - 1)),
                          g_h=rand_fr(rng, 1 << self.mh))
        return self._rand

    def rand_struct(self, cls=Randc, seed=99):
        r = self.randomness(seed)
        return cls(*[r[k].ctypes.data for k in ("mask_w", "g_w", "blind", "mask_h", "g_h")])


def oracle_prove(orc, r1cs: SyntheticR1CS, hash_version=2, seed=99) -> bytes:
    cs = r1cs.c_struct()
    rs = r1cs.rand_struct(seed=seed)
    out = ctypes.c_void_p()
    orc.orc_prove.restype = ctypes.c_int64
    n = orc.orc_prove(ctypes.byref(cs), ptr(r1cs.witness), ctypes.byref(rs), hash_version, ctypes.byref(out))
    assert n > 0, n
    data = ctypes.string_at(out, n)
    orc.orc_free(out)
    return data


def oracle_verify(orc, r1cs: SyntheticR1CS, transcript: bytes, hash_version=2) -> int:
    cs = r1cs.c_struct()
    buf = np.frombuffer(transcript, dtype=np.uint8)
    return orc.orc_verify(ctypes.byref(cs), ptr(buf), ctypes.c_size_t(len(transcript)), hash_version)
