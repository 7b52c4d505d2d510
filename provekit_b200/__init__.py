"""provekit_b200 — B200-native (sm_100a) WHIR hot path for ProveKit's `noir-r1cs prove`.

The product is the C-ABI shared library `libpkwhir.so` (include/pkwhir.h, provekit_b200/csrc/);
this package is only the host-side harness the tests and bench.py drive it through.  Importing the
package never falls back to a CPU implementation: without the built CUDA library `Context()` raises.
"""
from ._abi import LIB_PATH, PkError, build, lib  # noqa: F401
from .api import Buffer, Commitment, Context, Prover, np_decode, np_encode, nps_read_r1cs  # noqa: F401
