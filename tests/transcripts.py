"""Foreign Fiat-Shamir transcripts for pk_prove_with_transcript / orc_prove_with_transcript.

The reference keeps `ProverState<SkyscraperSponge, FieldElement>` on the Rust side
(provekit/prover/src/whir_r1cs.rs:57-59) and its spongefish internals are not in the tree (parity
unpinned).  The wholesale prover therefore takes the transcript as a table of callbacks; these are the two test
doubles:

  ToyTranscript      a deliberately DIFFERENT sponge (SHA-256 chain, its own codecs): proves that prover messages
                     depend on the transcript only through the callbacks; records every call for step-by-step diffs
  OracleTranscript   the in-tree Skyscraper sponge of the CPU oracle (oracle/transcript.c) behind the same table
"""
import ctypes
import hashlib

import numpy as np

from oracle import pyref as o

P = o.P
_U64P = ctypes.POINTER(ctypes.c_uint64)
_U8P = ctypes.POINTER(ctypes.c_uint8)
CB_SCALARS = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, _U64P, ctypes.c_size_t)
CB_BYTES = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, _U8P, ctypes.c_size_t)
CB_NEXT_HINT = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(_U8P), ctypes.POINTER(ctypes.c_size_t))


class PkVtbl(ctypes.Structure):
    """pk_transcript_vtbl (include/pkwhir.h) = the first five entries of orc_transcript_vtbl (oracle/pk_oracle.h)"""
    _fields_ = [("add_scalars", CB_SCALARS), ("challenge_scalars", CB_SCALARS), ("add_bytes", CB_BYTES),
                ("challenge_bytes", CB_BYTES), ("hint", CB_BYTES)]


class OrcVtbl(ctypes.Structure):
    _fields_ = PkVtbl._fields_ + [("next_scalars", CB_SCALARS), ("next_bytes", CB_BYTES), ("next_hint", CB_NEXT_HINT)]


def _scalars_in(ptr, n):
    a = np.ctypeslib.as_array(ptr, shape=(4 * n,)).astype(np.uint64)
    return [int(a[4 * i]) | int(a[4 * i + 1]) << 64 | int(a[4 * i + 2]) << 128 | int(a[4 * i + 3]) << 192 for i in range(n)]


def _scalars_out(ptr, vals):
    for i, v in enumerate(vals):
        for k in range(4):
            ptr[4 * i + k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF


class ToyTranscript:
    """SHA-256 chain: state' = H(state || tag || data); challenge = H(state || 'C' || counter) mod p.
    Prover mode appends to `narg`; verifier mode (proof given) reads it back.  Wire format is its own: scalars as 32-byte
    BIG-endian canonical integers, hints with a u64-LE length — nothing like spongefish, on purpose."""

    def __init__(self, proof: bytes = None, label: bytes = b"toy"):
        self.state = hashlib.sha256(label).digest()
        self.narg = bytearray()
        self.proof, self.rd = proof, 0
        self.log = []           # (op, payload) per call: step-by-step diff between two provers
        self.failed = False
        self._ctr = 0
        self._hint_keep = None
        self._cbs = [CB_SCALARS(self._add_scalars), CB_SCALARS(self._challenge_scalars), CB_BYTES(self._add_bytes),
                     CB_BYTES(self._challenge_bytes), CB_BYTES(self._hint), CB_SCALARS(self._next_scalars),
                     CB_BYTES(self._next_bytes), CB_NEXT_HINT(self._next_hint)]
        self.pk_vtbl = PkVtbl(*self._cbs[:5])
        self.orc_vtbl = OrcVtbl(*self._cbs)

    # ---- sponge ----
    def _absorb(self, tag: bytes, data: bytes):
        self.state = hashlib.sha256(self.state + tag + data).digest()
        self._ctr = 0

    def _squeeze(self) -> bytes:
        out = hashlib.sha256(self.state + b"C" + self._ctr.to_bytes(8, "little")).digest()
        self._ctr += 1
        return out

    # ---- prover side ----
    def _add_scalars(self, _u, ptr, n):
        mont = _scalars_in(ptr, n)
        canon = [x * o.R_INV % P for x in mont]
        data = b"".join(c.to_bytes(32, "big") for c in canon)
        self._absorb(b"S", data)
        self.narg += data
        self.log.append(("add_scalars", tuple(canon)))
        return 0

    def _challenge_scalars(self, _u, ptr, n):
        vals = []
        for _ in range(n):
            v = int.from_bytes(self._squeeze() + self._squeeze(), "little") % P
            vals.append(v)
        _scalars_out(ptr, [v * o.R % P for v in vals])
        self._absorb(b"c", b"")   # ratchet
        self.log.append(("challenge_scalars", tuple(vals)))
        return 0

    def _add_bytes(self, _u, ptr, n):
        data = bytes(ptr[:n])
        self._absorb(b"B", data)
        self.narg += data
        self.log.append(("add_bytes", data))
        return 0

    def _challenge_bytes(self, _u, ptr, n):
        out = b""
        while len(out) < n:
            out += self._squeeze()
        out = out[:n]
        for i in range(n):
            ptr[i] = out[i]
        self._absorb(b"b", b"")
        self.log.append(("challenge_bytes", out))
        return 0

    def _hint(self, _u, ptr, n):
        data = bytes(ptr[:n])
        self.narg += n.to_bytes(8, "little") + data
        self.log.append(("hint", data))
        return 0

    # ---- verifier side ----
    def _take(self, n):
        if self.proof is None or self.rd + n > len(self.proof):
            self.failed = True
            return None
        d = self.proof[self.rd:self.rd + n]
        self.rd += n
        return d

    def _next_scalars(self, _u, ptr, n):
        data = self._take(32 * n)
        if data is None:
            return 1
        canon = [int.from_bytes(data[32 * i:32 * i + 32], "big") for i in range(n)]
        if any(c >= P for c in canon):
            self.failed = True
            return 1
        _scalars_out(ptr, [c * o.R % P for c in canon])
        self._absorb(b"S", data)
        return 0

    def _next_bytes(self, _u, ptr, n):
        data = self._take(n)
        if data is None:
            return 1
        for i in range(n):
            ptr[i] = data[i]
        self._absorb(b"B", data)
        return 0

    def _next_hint(self, _u, pptr, nptr):
        ln = self._take(8)
        if ln is None:
            return 1
        n = int.from_bytes(ln, "little")
        data = self._take(n)
        if data is None:
            return 1
        self._hint_keep = (ctypes.c_uint8 * max(n, 1)).from_buffer_copy(data + (b"\0" if n == 0 else b""))
        pptr[0] = ctypes.cast(self._hint_keep, _U8P)
        nptr[0] = n
        return 0

    def exhausted(self) -> bool:
        return self.proof is not None and self.rd == len(self.proof) and not self.failed


class OracleTranscript:
    """The oracle's in-tree Skyscraper sponge (oracle/transcript.c) as a foreign transcript."""

    def __init__(self, orc, num_constraints: int, num_witnesses: int, proof: bytes = None):
        self.orc = orc
        orc.orc_fs_create.restype = ctypes.c_void_p
        orc.orc_fs_create.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
        orc.orc_fs_vtbl.restype = ctypes.c_void_p
        orc.orc_fs_narg.restype = ctypes.c_void_p
        orc.orc_fs_narg.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
        orc.orc_fs_free.argtypes = [ctypes.c_void_p]
        self._proof = np.frombuffer(proof, np.uint8).copy() if proof is not None else None
        self.user = ctypes.c_void_p(orc.orc_fs_create(num_constraints, num_witnesses,
                                                      self._proof.ctypes.data if proof is not None else None,
                                                      len(proof) if proof is not None else 0))
        self.vtbl_ptr = ctypes.c_void_p(orc.orc_fs_vtbl())

    def narg(self) -> bytes:
        n = ctypes.c_size_t()
        p = self.orc.orc_fs_narg(self.user, ctypes.byref(n))
        return ctypes.string_at(p, n.value) if n.value else b""

    def close(self):
        if self.user:
            self.orc.orc_fs_free(self.user)
            self.user = None


def orc_prove_with(orc, r1cs, vtbl_ptr, user, hash_version=2, seed=99) -> int:
    """oracle prover (tests/r1cs_util.SyntheticR1CS) driven by a foreign transcript; returns its status"""
    from helpers import ptr
    cs = r1cs.c_struct()
    rs = r1cs.rand_struct(seed=seed)
    orc.orc_prove_with_transcript.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p]
    return orc.orc_prove_with_transcript(ctypes.byref(cs), ptr(r1cs.witness), ctypes.byref(rs), hash_version, vtbl_ptr, user)


def orc_verify_with(orc, r1cs, vtbl_ptr, user, hash_version=2) -> int:
    cs = r1cs.c_struct()
    orc.orc_verify_with_transcript.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return orc.orc_verify_with_transcript(ctypes.byref(cs), hash_version, vtbl_ptr, user)
