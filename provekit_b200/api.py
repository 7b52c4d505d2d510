"""Host-side mirror (Python, for the test/bench harness) of the reference's operator interface on the
WHIR hot path.  Every call goes through the C-ABI of libpkwhir.so; nothing here computes on the CPU.

Names follow the reference:
  compress_many            skyscraper::CompressManyFn                     skyscraper/core/src/lib.rs:26
  pow_solve                SkyscraperPoW::solve                           provekit/common/src/skyscraper/pow.rs:27-29
  evals_to_coeffs          EvaluationsList::to_coeffs  [whir]             provekit/prover/src/whir_r1cs.rs:195
  commit_batch             CommitmentWriter::commit_batch [whir]          provekit/prover/src/whir_r1cs.rs:200-206
  sumcheck_fold_map_reduce provekit_common::utils::sumcheck::...          provekit/common/src/utils/sumcheck.rs:16-39
  whir_sumcheck_round      SumcheckSingle::compute_sumcheck_polynomial [whir]
  prove                    WhirR1CSProver::prove                          provekit/prover/src/whir_r1cs.rs:42-100
Field elements are numpy uint64 arrays of shape (n, 4): arkworks' Montgomery limbs.
"""
import ctypes
from ctypes import POINTER, byref, c_double, c_float, c_size_t, c_uint64, c_void_p

import numpy as np

from . import _abi
from ._abi import PkError


def _p(a):
    return a.ctypes.data_as(c_void_p) if a is not None else None


def _fe(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, 4)


class Buffer:
    def __init__(self, ctx: "Context", n: int):
        self.ctx, self.n = ctx, n
        h = c_void_p()
        ctx._chk(ctx.L.pk_buf_alloc(ctx.h, n, byref(h)))
        self.h = h

    def upload(self, arr, off: int = 0) -> "Buffer":
        arr = _fe(arr)
        self.ctx._chk(self.ctx.L.pk_buf_upload(self.ctx.h, self.h, off, _p(arr), len(arr)))
        return self

    def download(self, n: int = None, off: int = 0) -> np.ndarray:
        n = self.n - off if n is None else n
        out = np.empty((n, 4), np.uint64)
        self.ctx._chk(self.ctx.L.pk_buf_download(self.ctx.h, self.h, off, _p(out), n))
        return out

    def clone(self) -> "Buffer":
        """device-to-device copy into a fresh buffer (pk_buf_copy)"""
        out = Buffer(self.ctx, self.n)
        self.ctx._chk(self.ctx.L.pk_buf_copy(self.ctx.h, out.h, 0, self.h, 0, self.n))
        return out

    def zero(self, off=0, n=None):
        self.ctx._chk(self.ctx.L.pk_buf_zero(self.ctx.h, self.h, off, self.n - off if n is None else n))
        return self

    def rng_fill(self, seed: bytes, stream: int, off=0, n=None):
        """n uniform field elements from the ChaCha12 counter stream (seed, stream) — pk_rng_fill."""
        assert len(seed) == 32
        self.ctx._chk(self.ctx.L.pk_rng_fill(self.ctx.h, self.h, off, self.n - off if n is None else n, seed, stream))
        return self

    def free(self):
        if self.h:
            self.ctx.L.pk_buf_free(self.ctx.h, self.h)
            self.h = None

    @property
    def device_ptr(self) -> int:
        return self.ctx.L.pk_buf_device_ptr(self.h)


class Commitment:
    """Witness{merkle_tree, merkle_leaves} of whir's commit_batch, resident on the device."""

    def __init__(self, ctx, h, root):
        self.ctx, self.h, self.root = ctx, h, root
        self.num_leaves = ctx.L.pk_commit_num_leaves(h)
        self.leaf_width = ctx.L.pk_commit_leaf_width(h)
        self.depth = self.num_leaves.bit_length() - 1

    def open(self, sorted_indexes):
        """STIR answers + ark MultiPath pieces for strictly increasing leaf indexes.
        Returns (leaves (n,w,4) Montgomery, siblings (n,4) canonical, prefix_lens, suffixes list of (k,4))."""
        idx = np.ascontiguousarray(sorted_indexes, dtype=np.uint64)
        n = len(idx)
        leaves = np.empty((n * self.leaf_width, 4), np.uint64)
        sib = np.empty((n, 4), np.uint64)
        pre = np.empty(n, np.uint64)
        slen = np.empty(n, np.uint64)
        cap = n * max(self.depth, 1)
        suf = np.empty((cap, 4), np.uint64)
        self.ctx._chk(self.ctx.L.pk_commit_open(self.ctx.h, self.h, _p(idx), n, _p(leaves), _p(sib), _p(pre), _p(suf),
                                               _p(slen), cap))
        sufs, pos = [], 0
        for k in slen:
            sufs.append(suf[pos:pos + int(k)].copy())
            pos += int(k)
        return leaves.reshape(n, self.leaf_width, 4), sib, pre, sufs

    def open_paths(self, sorted_indexes):
        """STIR answers + UNCOMPRESSED authentication paths (pk_commit_open_paths): (leaves (n,w,4) Montgomery,
        paths (n, depth, 4) canonical, level 0 = sibling leaf digest .. level depth-1 = sibling below the root)."""
        idx = np.ascontiguousarray(sorted_indexes, dtype=np.uint64)
        n = len(idx)
        leaves = np.empty((n * self.leaf_width, 4), np.uint64)
        paths = np.empty((n * max(self.depth, 1), 4), np.uint64)
        self.ctx._chk(self.ctx.L.pk_commit_open_paths(self.ctx.h, self.h, _p(idx), n, _p(leaves), _p(paths)))
        return leaves.reshape(n, self.leaf_width, 4), paths[:n * self.depth].reshape(n, self.depth, 4)

    def free(self):
        if self.h:
            self.ctx.L.pk_commit_free(self.ctx.h, self.h)
            self.h = None


class Context:
    def __init__(self, device: int = 0):
        self.L = _abi.lib()
        h = c_void_p()
        rc = self.L.pk_ctx_create(device, byref(h))
        if rc != 0:
            raise PkError(rc, "pk_ctx_create failed (no usable CUDA device? this package has no CPU fallback)")
        self.h = h

    def _chk(self, rc):
        if rc != 0:
            raise PkError(rc, self.L.pk_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.pk_ctx_destroy(self.h)
            self.h = None

    @property
    def launches(self) -> int:
        return self.L.pk_launch_count(self.h)

    @property
    def stream(self) -> int:
        return self.L.pk_ctx_stream(self.h)

    def sync(self):
        self._chk(self.L.pk_ctx_sync(self.h))

    def buffer(self, n: int) -> Buffer:
        return Buffer(self, n)

    def buffer_shared(self, n: int) -> Buffer:
        """IPC-exportable buffer (cudaMalloc) for the leaf block of a sharded commit."""
        b = Buffer.__new__(Buffer)
        b.ctx, b.n = self, n
        h = c_void_p()
        self._chk(self.L.pk_buf_alloc_shared(self.h, n, byref(h)))
        b.h = h
        return b

    def ipc_export(self, buf: Buffer) -> bytes:
        out = (ctypes.c_uint8 * 64)()
        self._chk(self.L.pk_ipc_export(self.h, buf.h, out))
        return bytes(out)

    def ipc_open(self, handle: bytes) -> int:
        h = (ctypes.c_uint8 * 64)(*handle)
        p = c_void_p()
        self._chk(self.L.pk_ipc_open(self.h, h, byref(p)))
        return p.value

    def ipc_close(self, dptr: int):
        self._chk(self.L.pk_ipc_close(self.h, c_void_p(dptr)))

    def rs_encode_sharded(self, coeffs: Buffer, log_n: int, log_inv_rate: int, col_first: int, n_cols: int, peer_ptrs,
                          leaf_stride: int, col_offset: int, fold: int = 4):
        """columns [col_first, col_first+n_cols) of one polynomial; row r -> peer_ptrs[r // (rows/len(peer_ptrs))]"""
        arr = (c_void_p * len(peer_ptrs))(*[c_void_p(p) for p in peer_ptrs])
        self._chk(self.L.pk_rs_encode_sharded(self.h, coeffs.h, log_n, log_inv_rate, fold, col_first, n_cols, arr,
                                              len(peer_ptrs), leaf_stride, col_offset))

    def commit_wrap(self, leaves: Buffer, nodes: Buffer, num_leaves: int, leaf_width: int) -> Commitment:
        """a shard's leaf block + sub-tree as a (non-owning) Commitment: open() / open_paths() with local row indexes"""
        h = c_void_p()
        self._chk(self.L.pk_commit_wrap(self.h, leaves.h, nodes.h, num_leaves, leaf_width, byref(h)))
        return Commitment(self, h, None)

    def multipath_build(self, paths):
        """(n, depth, 4) uncompressed paths -> ark MultiPath pieces (siblings (n,4), prefix_lens, suffixes), pk_multipath_build"""
        paths = np.ascontiguousarray(paths, dtype=np.uint64)
        n, depth = paths.shape[0], paths.shape[1]
        sib = np.empty((n, 4), np.uint64)
        pre = np.empty(n, np.uint64)
        slen = np.empty(n, np.uint64)
        cap = max(n * depth, 1)
        suf = np.empty((cap, 4), np.uint64)
        self._chk(self.L.pk_multipath_build(self.h, _p(paths), n, depth, _p(sib), _p(pre), _p(suf), _p(slen), cap))
        sufs, pos = [], 0
        for k in slen:
            sufs.append(suf[pos:pos + int(k)].copy())
            pos += int(k)
        return sib, pre, sufs

    def merkle_combine_roots(self, roots) -> np.ndarray:
        r = np.ascontiguousarray(roots, dtype=np.uint64).reshape(-1, 4)
        out = np.empty(4, np.uint64)
        self._chk(self.L.pk_merkle_combine_roots(self.h, _p(r), len(r), _p(out)))
        return out

    def upload(self, arr) -> Buffer:
        arr = _fe(arr)
        return Buffer(self, len(arr)).upload(arr)

    # ---- seams -------------------------------------------------------------------------------
    def compress_many(self, messages: bytes) -> bytes:
        if len(messages) % 64:
            raise ValueError("Message length not a multiple of 64")  # generic.rs:19
        n = len(messages) // 64
        src = np.frombuffer(messages, dtype=np.uint8)
        out = np.empty(n * 32, np.uint8)
        self._chk(self.L.pk_skyscraper_compress_many(self.h, _p(src), _p(out), n))
        return out.tobytes()

    def pow_solve(self, challenge, bits: float) -> int:
        ch = np.ascontiguousarray(challenge, dtype=np.uint64)
        nonce = c_uint64()
        self._chk(self.L.pk_pow_solve(self.h, _p(ch), c_double(bits), byref(nonce)))
        return nonce.value

    def evals_to_coeffs(self, buf: Buffer, log_n: int):
        self._chk(self.L.pk_evals_to_coeffs(self.h, buf.h, log_n))

    def coeffs_to_evals(self, buf: Buffer, log_n: int):
        self._chk(self.L.pk_coeffs_to_evals(self.h, buf.h, log_n))

    def commit_batch(self, polys, log_n: int, log_inv_rate: int = 1, fold: int = 4) -> Commitment:
        arr = (c_void_p * len(polys))(*[p.h for p in polys])
        h = c_void_p()
        root = np.empty(4, np.uint64)
        self._chk(self.L.pk_commit_batch(self.h, arr, len(polys), log_n, log_inv_rate, fold, byref(h), _p(root)))
        return Commitment(self, h, root)

    def rs_encode(self, coeffs: Buffer, log_n, log_inv_rate, leaves: Buffer, leaf_stride=16, col_offset=0, fold=4):
        self._chk(self.L.pk_rs_encode(self.h, coeffs.h, log_n, log_inv_rate, fold, leaves.h, leaf_stride, col_offset))

    def merkle_build(self, leaves: Buffer, num_leaves: int, leaf_width: int, nodes: Buffer):
        self._chk(self.L.pk_merkle_build(self.h, leaves.h, num_leaves, leaf_width, nodes.h))

    def eval_univariate(self, coeffs: Buffer, n: int, z) -> np.ndarray:
        out = np.empty(4, np.uint64)
        self._chk(self.L.pk_eval_univariate(self.h, coeffs.h, n, _p(_fe(z)), _p(out)))
        return out

    def eval_univariate_batch(self, polys, n: int, z) -> np.ndarray:
        arr = (c_void_p * len(polys))(*[p.h for p in polys])
        out = np.empty((len(polys), 4), np.uint64)
        self._chk(self.L.pk_eval_univariate_batch(self.h, arr, len(polys), n, _p(_fe(z)), _p(out)))
        return out

    def multi_dot(self, a_list, b_list, n: int) -> np.ndarray:
        """out[ja, jb] = <a_ja, b_jb> in one pass; shapes (3,2) or (1,2)."""
        aa = (c_void_p * len(a_list))(*[p.h for p in a_list])
        bb = (c_void_p * len(b_list))(*[p.h for p in b_list])
        out = np.empty((len(a_list) * len(b_list), 4), np.uint64)
        self._chk(self.L.pk_multi_dot(self.h, aa, len(a_list), bb, len(b_list), n, _p(out)))
        return out.reshape(len(a_list), len(b_list), 4)

    def mle_eval_batch(self, evals_list, log_n: int, point) -> np.ndarray:
        arr = (c_void_p * len(evals_list))(*[p.h for p in evals_list])
        out = np.empty((len(evals_list), 4), np.uint64)
        self._chk(self.L.pk_mle_eval_batch(self.h, arr, len(evals_list), log_n, _p(_fe(point)), _p(out)))
        return out

    def mle_eval_batch_prefix(self, evals_list, log_n: int, n_prefix: int, point) -> np.ndarray:
        """mle_eval_batch for arrays that vanish beyond their first n_prefix elements (only the prefix is read)"""
        arr = (c_void_p * len(evals_list))(*[p.h for p in evals_list])
        out = np.empty((len(evals_list), 4), np.uint64)
        self._chk(self.L.pk_mle_eval_batch_prefix(self.h, arr, len(evals_list), log_n, n_prefix, _p(_fe(point)), _p(out)))
        return out

    def axpy(self, y: Buffer, x: Buffer, a, n: int):
        self._chk(self.L.pk_axpy(self.h, y.h, x.h, _p(_fe(a)), n))

    def dot(self, a: Buffer, b: Buffer, n: int) -> np.ndarray:
        out = np.empty(4, np.uint64)
        self._chk(self.L.pk_dot(self.h, a.h, b.h, n, _p(out)))
        return out

    def eval_eq(self, points, scalars, n: int, out: Buffer):
        pts, sc = _fe(points), _fe(scalars)
        self._chk(self.L.pk_eval_eq_batch(self.h, _p(pts), len(sc), n, _p(sc), out.h))

    def eval_eq_roots(self, exps, log_d: int, n: int, scalars, out: Buffer):
        """out[x] += sum_k scalars[k] * eq(pow(omega_D^exps[k]), x) over n variables (pk_eval_eq_roots_batch)"""
        e = np.ascontiguousarray(exps, dtype=np.uint64)
        sc = _fe(scalars)
        self._chk(self.L.pk_eval_eq_roots_batch(self.h, _p(e), len(e), log_d, n, _p(sc), out.h))

    def mle_eval(self, evals: Buffer, log_n: int, point) -> np.ndarray:
        out = np.empty(4, np.uint64)
        self._chk(self.L.pk_mle_eval(self.h, evals.h, log_n, _p(_fe(point)), _p(out)))
        return out

    def fold_coeffs(self, coeffs: Buffer, log_n: int, r, out: Buffer):
        r = _fe(r)
        self._chk(self.L.pk_fold_coeffs(self.h, coeffs.h, log_n, _p(r), len(r), out.h))

    def sumcheck_fold_map_reduce(self, a: Buffer, b: Buffer, c: Buffer, eq: Buffer, log_n: int, fold=None) -> np.ndarray:
        """[f(0), f(-1), f(inf)] of the zk-sumcheck round; folds the four arrays in place first when
        `fold` is given (log_n = length before folding)."""
        out = np.empty((3, 4), np.uint64)
        f = _fe(fold) if fold is not None else None
        self._chk(self.L.pk_zk_sumcheck_round(self.h, a.h, b.h, c.h, eq.h, log_n, _p(f), _p(out)))
        return out

    def whir_sumcheck_round(self, p_in: Buffer, w_in: Buffer, log_n: int, fold=None, p_out: Buffer = None,
                            w_out: Buffer = None) -> np.ndarray:
        out = np.empty((3, 4), np.uint64)
        f = _fe(fold) if fold is not None else None
        self._chk(self.L.pk_whir_sumcheck_round(self.h, p_in.h, w_in.h, p_out.h if p_out else None,
                                                w_out.h if w_out else None, log_n, _p(f), _p(out)))
        return out

    # ---- sharded sumchecks (one process per GPU; see include/pkwhir.h "multi-GPU sumchecks") ----
    def shard_mailbox(self) -> Buffer:
        """this rank's zeroed mailbox (IPC-exportable); barrier across ranks before the first sharded round"""
        mb = self.buffer_shared(self.L.pk_shard_mailbox_elems())
        mb.zero()
        self.sync()
        return mb

    def shard_group(self, rank: int, world: int, mailbox_ptrs):
        arr = (c_void_p * world)(*[c_void_p(int(p)) for p in mailbox_ptrs])
        self._chk(self.L.pk_shard_group_set(self.h, rank, world, arr))

    def shard_barrier(self):
        """barrier on the stream across the shard group (asynchronous, pk_shard_barrier)"""
        self._chk(self.L.pk_shard_barrier(self.h))

    def shard_allgather(self, src: Buffer, off: int, world: int) -> np.ndarray:
        """(world, 4): rank r's src[off], gathered over the peer mailboxes (pk_shard_allgather; synchronises)"""
        out = np.empty((world, 4), np.uint64)
        self._chk(self.L.pk_shard_allgather(self.h, src.h, off, _p(out)))
        return out

    def sumcheck_fold_map_reduce_sharded(self, a: Buffer, b: Buffer, c: Buffer, eq: Buffer, log_n: int, fold=None) -> np.ndarray:
        """the round on the local low-bit shard + exchange over peer memory: the GLOBAL [f(0), f(-1), f(inf)]"""
        out = np.empty((3, 4), np.uint64)
        f = _fe(fold) if fold is not None else None
        self._chk(self.L.pk_zk_sumcheck_round_sharded(self.h, a.h, b.h, c.h, eq.h, log_n, _p(f), _p(out)))
        return out

    def whir_sumcheck_round_sharded(self, p_in: Buffer, w_in: Buffer, log_n: int, fold=None, p_out: Buffer = None,
                                    w_out: Buffer = None) -> np.ndarray:
        out = np.empty((3, 4), np.uint64)
        f = _fe(fold) if fold is not None else None
        self._chk(self.L.pk_whir_sumcheck_round_sharded(self.h, p_in.h, w_in.h, p_out.h if p_out else None,
                                                        w_out.h if w_out else None, log_n, _p(f), _p(out)))
        return out

    def modmul_bench(self, n_threads: int, iters: int, square: bool = False) -> float:
        ms = c_float()
        f = self.L.pk_modsqr_bench if square else self.L.pk_modmul_bench
        self._chk(f(self.h, n_threads, iters, byref(ms)))
        return ms.value


def np_encode(transcript: bytes) -> bytes:
    """WhirR1CSProof transcript -> bytes of a `.np` file (provekit/common/src/file/bin.rs write_bin)."""
    L = _abi.lib()
    src = np.frombuffer(transcript, dtype=np.uint8)
    out, n = c_void_p(), c_size_t()
    rc = L.pk_np_encode(_p(src) if len(src) else None, len(src), byref(out), byref(n))
    if rc != 0:
        raise PkError(rc, "pk_np_encode failed")
    data = ctypes.string_at(out, n.value)
    L.pk_free(out)
    return data


def np_decode(file_bytes: bytes) -> bytes:
    """bytes of a `.np` file -> transcript (read_bin + postcard)."""
    L = _abi.lib()
    src = np.frombuffer(file_bytes, dtype=np.uint8)
    out, n = c_void_p(), c_size_t()
    rc = L.pk_np_decode(_p(src), len(src), byref(out), byref(n))
    if rc != 0:
        raise PkError(rc, "pk_np_decode: not a valid .np container")
    data = ctypes.string_at(out, n.value)
    L.pk_free(out)
    return data


def nps_read_r1cs(file_bytes: bytes) -> dict:
    """bytes of a `.nps` file (NoirProofScheme) -> the R1CS dict Prover takes (pk_nps_read_r1cs): num_constraints,
    num_witnesses, interned (Montgomery, (k,4) uint64), a/b/c = (row_start u64, col u32, val u32), num_public_inputs."""
    L = _abi.lib()
    src = np.frombuffer(file_bytes, dtype=np.uint8)
    h = c_void_p()
    rc = L.pk_nps_read_r1cs(_p(src), len(src), byref(h))
    if rc != 0:
        raise PkError(rc, "pk_nps_read_r1cs: not a .nps container or no R1CS found in it")
    r = L.pk_nps_r1cs(h).contents

    def arr(ptr, n, dt):
        if n == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(ctypes.cast(ptr, POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()

    def csr(m):
        return (arr(m.row_start, m.num_rows, np.uint64), arr(m.col, m.nnz, np.uint32), arr(m.val, m.nnz, np.uint32))

    out = dict(num_constraints=int(r.num_constraints), num_witnesses=int(r.num_witnesses),
               interned=arr(r.interned, 4 * r.num_interned, np.uint64).reshape(-1, 4), a=csr(r.a), b=csr(r.b), c=csr(r.c),
               num_public_inputs=int(L.pk_nps_num_public_inputs(h)))
    L.pk_nps_free(h)
    return out


class Prover:
    """WhirR1CSProver::prove (provekit/prover/src/whir_r1cs.rs:42-100) over a device-resident R1CS.
    `r1cs` is a dict with num_constraints, num_witnesses, interned (k,4) and a/b/c = (row_start u64,
    col u32, val u32) in the reference's interned-CSR form (provekit/common/src/sparse_matrix.rs:19-26)."""

    def __init__(self, ctx: Context, r1cs: dict):
        self.ctx = ctx
        self._keep = []

        def csr(t):
            rs, col, val = (np.ascontiguousarray(t[0], np.uint64), np.ascontiguousarray(t[1], np.uint32),
                            np.ascontiguousarray(t[2], np.uint32))
            self._keep += [rs, col, val]
            return _abi.CSR(r1cs["num_constraints"], r1cs["num_witnesses"], len(col), rs.ctypes.data, col.ctypes.data,
                            val.ctypes.data)

        interned = _fe(r1cs["interned"])
        self._keep.append(interned)
        s = _abi.R1CS(r1cs["num_constraints"], r1cs["num_witnesses"], len(interned), interned.ctypes.data,
                      csr(r1cs["a"]), csr(r1cs["b"]), csr(r1cs["c"]))
        h = c_void_p()
        ctx._chk(ctx.L.pk_prover_create(ctx.h, byref(s), byref(h)))
        self.h = h
        self.num_witnesses = int(r1cs["num_witnesses"])
        self.num_constraints = int(r1cs["num_constraints"])
        m, m0, mh = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        ctx.L.pk_prover_shapes(h, byref(m), byref(m0), byref(mh))
        self.m, self.m0, self.mh = m.value, m0.value, mh.value

    def matvec(self, which: int, x: "Buffer", transposed: bool = False) -> "Buffer":
        """HydratedSparseMatrix * vector (sparse_matrix.rs:148-165) or vector * matrix (:168-184); which = 0 A, 1 B, 2 C"""
        n_out = self.num_witnesses if transposed else self.num_constraints
        out = Buffer(self.ctx, n_out)
        self.ctx._chk(self.ctx.L.pk_prover_matvec(self.h, which, 1 if transposed else 0, x.h, out.h))
        return out

    def _witness(self, witness):
        """the C-ABI reads exactly num_witnesses elements from a bare pointer: refuse anything else here
        (the reference asserts witness.len() == r1cs.num_witnesses(), provekit/prover/src/whir_r1cs.rs:48-54)"""
        w = _fe(witness)
        if len(w) != self.num_witnesses:
            raise PkError(-1, f"witness holds {len(w)} elements, the R1CS has {self.num_witnesses} witnesses")
        return w

    def _rand(self, rand: dict):
        want = {"mask_w": 1 << (self.m - 1), "g_w": 1 << self.m, "blind": 4 * self.m0, "mask_h": 1 << (self.mh - 1),
                "g_h": 1 << self.mh}
        arrs = [_fe(rand[k]) for k in ("mask_w", "g_w", "blind", "mask_h", "g_h")]
        for (k, n), a in zip(want.items(), arrs):
            if len(a) != n:
                raise PkError(-1, f"randomness array {k} holds {len(a)} elements, the scheme needs {n}")
        return _abi.Rand(*[a.ctypes.data for a in arrs]), arrs

    def _take(self, out, n) -> bytes:
        data = ctypes.string_at(out, n.value)
        self.ctx.L.pk_free(out)
        return data

    def prove(self, witness, rand: dict) -> bytes:
        """Host buffers in, spongefish NARG string (WhirR1CSProof.transcript) out."""
        w = self._witness(witness)
        rs, _keep = self._rand(rand)
        out, n = c_void_p(), c_size_t()
        self.ctx._chk(self.ctx.L.pk_prove(self.h, _p(w), byref(rs), byref(out), byref(n)))
        return self._take(out, n)

    def prove_with_transcript(self, witness, rand: dict, vtbl, user=None) -> None:
        """WhirR1CSProver::prove with the caller's Fiat-Shamir transcript: `vtbl` points to a pk_transcript_vtbl
        (ctypes structure or raw address), `user` is handed back to every callback.  The proof string lives on the
        caller's side (spongefish `ProverState::narg_string()`)."""
        w = self._witness(witness)
        rs, _keep = self._rand(rand)
        vt = vtbl if isinstance(vtbl, (int, c_void_p)) else ctypes.cast(ctypes.pointer(vtbl), c_void_p)
        self.ctx._chk(self.ctx.L.pk_prove_with_transcript(self.h, _p(w), byref(rs), vt, user))

    def upload_inputs(self, witness, rand: dict):
        w = self._witness(witness)
        rs, _keep = self._rand(rand)
        self.ctx._chk(self.ctx.L.pk_prover_upload_inputs(self.h, _p(w), byref(rs)))

    def prove_seeded(self, witness, seed: bytes) -> bytes:
        """pk_prove with the masks drawn on the device from a 32-byte seed: only the witness crosses PCIe."""
        assert len(seed) == 32
        w = self._witness(witness)
        out, n = c_void_p(), c_size_t()
        self.ctx._chk(self.ctx.L.pk_prove_seeded(self.h, _p(w), seed, byref(out), byref(n)))
        return self._take(out, n)

    def upload_inputs_seeded(self, witness, seed: bytes):
        assert len(seed) == 32
        w = self._witness(witness)
        self.ctx._chk(self.ctx.L.pk_prover_upload_inputs_seeded(self.h, _p(w), seed))

    def prove_staged(self) -> bytes:
        out, n = c_void_p(), c_size_t()
        self.ctx._chk(self.ctx.L.pk_prove_staged(self.h, byref(out), byref(n)))
        return self._take(out, n)

    # ---- asynchronous form (device transcript): enqueue returns at once, collect waits for the proof string ----
    def enqueue_staged(self):
        self.ctx._chk(self.ctx.L.pk_prove_staged_enqueue(self.h))

    def enqueue_seeded(self, witness, seed: bytes):
        assert len(seed) == 32
        self._w_keep = self._witness(witness)  # must stay alive until collect()
        self.ctx._chk(self.ctx.L.pk_prove_seeded_enqueue(self.h, _p(self._w_keep), seed))

    def enqueue(self, witness, rand: dict):
        self._w_keep = self._witness(witness)
        rs, self._r_keep = self._rand(rand)
        self.ctx._chk(self.ctx.L.pk_prove_enqueue(self.h, _p(self._w_keep), byref(rs)))

    def collect(self) -> bytes:
        out, n = c_void_p(), c_size_t()
        self.ctx._chk(self.ctx.L.pk_prove_collect(self.h, byref(out), byref(n)))
        return self._take(out, n)

    def set_host_transcript(self, on: bool):
        """in-tree sponge on the host (True) instead of on the device (default)"""
        self.ctx._chk(self.ctx.L.pk_prover_set_host_transcript(self.h, 1 if on else 0))

    @property
    def host_syncs(self) -> int:
        """stream synchronisations of the last proof"""
        return int(self.ctx.L.pk_prover_host_syncs(self.h))

    def timings(self):
        t = (c_double * 9)()
        self.ctx.L.pk_prover_timings(self.h, t)
        return list(t)

    def close(self):
        if self.h:
            self.ctx.L.pk_prover_destroy(self.h)
            self.h = None
