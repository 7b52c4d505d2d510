"""CPU oracle for the WHIR hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  `oracle.pyref` is the pure-Python root of the parity chain; `oracle.lib()` loads the
C restatement (oracle/*.c -> libpkoracle.so, built by `make -C oracle` / __graft_entry__.build()).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpkoracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpkoracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.orc_pow_solve.restype = ctypes.c_uint64
        L.orc_pow_solve.argtypes = [ctypes.c_void_p, ctypes.c_double]
        L.orc_pow_verify.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_uint64]
        L.orc_pow_threshold.argtypes = [ctypes.c_double, ctypes.c_void_p]
        if hasattr(L, "orc_prove"):
            L.orc_prove.restype = ctypes.c_int64
        _LIB = L
    return _LIB
