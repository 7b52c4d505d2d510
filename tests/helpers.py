"""Shared test helpers: numpy <-> field-element conversions and seeded inputs."""
import ctypes

import numpy as np

from oracle import pyref as o

P = o.P


def ints_to_arr(xs) -> np.ndarray:
    """list of python ints -> (n,4) uint64 little-endian limbs (no Montgomery conversion)."""
    out = np.empty((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        for k in range(4):
            out[i, k] = (x >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def arr_to_ints(a: np.ndarray):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in a]


def to_mont(xs) -> np.ndarray:
    return ints_to_arr([x * o.R % P for x in xs])


def from_mont(a: np.ndarray):
    return [x * o.R_INV % P for x in arr_to_ints(a)]


def ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def rand_fr(rng: np.random.Generator, n: int) -> np.ndarray:
    """n uniform-ish canonical field elements as (n,4) uint64 (top limb masked then rejected to < p).
    The result is a valid Montgomery-form array too (any value < p is)."""
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
    # reject >= p by clearing one more bit on the few offenders (keeps determinism, stays < p)
    top = np.uint64(0x30644e72e131a029)
    bad = a[:, 3] >= top
    a[bad, 3] &= np.uint64(0x1FFFFFFFFFFFFFFF)
    return a
