"""ctypes binding of libpkwhir.so (include/pkwhir.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback in this package."""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_size_t, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PKWHIR_LIB") or os.path.join(_HERE, "libpkwhir.so")  # override: kernel-variant experiments
_LIB = None


class PkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pkwhir error {code}: {msg}")
        self.code = code


class CSR(ctypes.Structure):
    _fields_ = [("num_rows", c_uint64), ("num_cols", c_uint64), ("nnz", c_uint64),
                ("row_start", c_void_p), ("col", c_void_p), ("val", c_void_p)]


class R1CS(ctypes.Structure):
    _fields_ = [("num_constraints", c_uint64), ("num_witnesses", c_uint64), ("num_interned", c_uint64),
                ("interned", c_void_p), ("a", CSR), ("b", CSR), ("c", CSR)]


class Rand(ctypes.Structure):
    _fields_ = [("mask_w", c_void_p), ("g_w", c_void_p), ("blind", c_void_p), ("mask_h", c_void_p),
                ("g_h", c_void_p)]


def build(force: bool = False) -> str:
    """Compile provekit_b200/csrc for sm_100a into provekit_b200/libpkwhir.so (nvcc cross-compiles
    without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-s", "clean"])
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-s"])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is required; there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, u64p = c_void_p, c_size_t, c_void_p
    sig = {
        "pk_set_blocking_sync": (c_int, [c_int, c_int]),
        "pk_ctx_create": (c_int, [c_int, POINTER(vp)]),
        "pk_ctx_destroy": (None, [vp]),
        "pk_last_error": (c_char_p, [vp]),
        "pk_version": (c_char_p, []),
        "pk_launch_count": (c_uint64, [vp]),
        "pk_ctx_stream": (vp, [vp]),
        "pk_ctx_sync": (c_int, [vp]),
        "pk_buf_alloc": (c_int, [vp, sz, POINTER(vp)]),
        "pk_buf_free": (None, [vp, vp]),
        "pk_buf_len": (sz, [vp]),
        "pk_buf_device_ptr": (vp, [vp]),
        "pk_buf_upload": (c_int, [vp, vp, sz, u64p, sz]),
        "pk_buf_download": (c_int, [vp, vp, sz, u64p, sz]),
        "pk_buf_copy": (c_int, [vp, vp, sz, vp, sz, sz]),
        "pk_buf_zero": (c_int, [vp, vp, sz, sz]),
        "pk_skyscraper_compress_many": (c_int, [vp, vp, vp, sz]),
        "pk_skyscraper_compress_many_dev": (c_int, [vp, vp, vp, sz]),
        "pk_pow_solve": (c_int, [vp, u64p, c_double, POINTER(c_uint64)]),
        "pk_evals_to_coeffs": (c_int, [vp, vp, c_int]),
        "pk_coeffs_to_evals": (c_int, [vp, vp, c_int]),
        "pk_commit_batch": (c_int, [vp, POINTER(vp), c_int, c_int, c_int, c_int, POINTER(vp), u64p]),
        "pk_commit_free": (None, [vp, vp]),
        "pk_commit_num_leaves": (sz, [vp]),
        "pk_commit_leaf_width": (sz, [vp]),
        "pk_rs_encode": (c_int, [vp, vp, c_int, c_int, c_int, vp, sz, sz]),
        "pk_merkle_build": (c_int, [vp, vp, sz, sz, vp]),
        "pk_buf_alloc_shared": (c_int, [vp, sz, POINTER(vp)]),
        "pk_ipc_export": (c_int, [vp, vp, vp]),
        "pk_ipc_open": (c_int, [vp, vp, POINTER(vp)]),
        "pk_ipc_close": (c_int, [vp, vp]),
        "pk_rs_encode_sharded": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, POINTER(vp), c_int, sz, sz]),
        "pk_merkle_combine_roots": (c_int, [vp, u64p, c_int, u64p]),
        "pk_commit_open": (c_int, [vp, vp, u64p, sz, u64p, u64p, u64p, u64p, u64p, sz]),
        "pk_commit_wrap": (c_int, [vp, vp, vp, sz, sz, POINTER(vp)]),
        "pk_commit_open_paths": (c_int, [vp, vp, u64p, sz, u64p, u64p]),
        "pk_multipath_build": (c_int, [vp, u64p, sz, c_int, u64p, u64p, u64p, u64p, sz]),
        "pk_eval_univariate": (c_int, [vp, vp, sz, u64p, u64p]),
        "pk_eval_univariate_batch": (c_int, [vp, POINTER(vp), c_int, sz, u64p, u64p]),
        "pk_multi_dot": (c_int, [vp, POINTER(vp), c_int, POINTER(vp), c_int, sz, u64p]),
        "pk_mle_eval_batch": (c_int, [vp, POINTER(vp), c_int, c_int, u64p, u64p]),
        "pk_mle_eval_batch_prefix": (c_int, [vp, POINTER(vp), c_int, c_int, sz, u64p, u64p]),
        "pk_axpy": (c_int, [vp, vp, vp, u64p, sz]),
        "pk_dot": (c_int, [vp, vp, vp, sz, u64p]),
        "pk_eval_eq": (c_int, [vp, u64p, c_int, u64p, vp]),
        "pk_eval_eq_batch": (c_int, [vp, u64p, sz, c_int, u64p, vp]),
        "pk_eval_eq_roots_batch": (c_int, [vp, u64p, sz, c_int, c_int, u64p, vp]),
        "pk_mle_eval": (c_int, [vp, vp, c_int, u64p, u64p]),
        "pk_fold_coeffs": (c_int, [vp, vp, c_int, u64p, c_int, vp]),
        "pk_zk_sumcheck_round": (c_int, [vp, vp, vp, vp, vp, c_int, u64p, u64p]),
        "pk_whir_sumcheck_round": (c_int, [vp, vp, vp, vp, vp, c_int, u64p, u64p]),
        "pk_shard_mailbox_elems": (sz, []),
        "pk_shard_group_set": (c_int, [vp, c_int, c_int, POINTER(vp)]),
        "pk_shard_group_clear": (c_int, [vp]),
        "pk_shard_barrier": (c_int, [vp]),
        "pk_shard_allgather": (c_int, [vp, vp, sz, u64p]),
        "pk_zk_sumcheck_round_sharded": (c_int, [vp, vp, vp, vp, vp, c_int, u64p, u64p]),
        "pk_whir_sumcheck_round_sharded": (c_int, [vp, vp, vp, vp, vp, c_int, u64p, u64p]),
        "pk_prover_create": (c_int, [vp, POINTER(R1CS), POINTER(vp)]),
        "pk_prover_destroy": (None, [vp]),
        "pk_prover_matvec": (c_int, [vp, c_int, c_int, vp, vp]),
        "pk_prover_set_host_transcript": (c_int, [vp, c_int]),
        "pk_prover_host_syncs": (c_uint64, [vp]),
        "pk_prover_shapes": (None, [vp, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
        "pk_prove": (c_int, [vp, u64p, POINTER(Rand), POINTER(vp), POINTER(sz)]),
        "pk_prove_enqueue": (c_int, [vp, u64p, POINTER(Rand)]),
        "pk_prove_seeded_enqueue": (c_int, [vp, u64p, c_char_p]),
        "pk_prove_staged_enqueue": (c_int, [vp]),
        "pk_prove_collect": (c_int, [vp, POINTER(vp), POINTER(sz)]),
        "pk_prove_with_transcript": (c_int, [vp, u64p, POINTER(Rand), vp, vp]),
        "pk_prove_staged_with_transcript": (c_int, [vp, vp, vp]),
        "pk_free": (None, [vp]),
        "pk_prover_timings": (None, [vp, POINTER(c_double)]),
        "pk_np_encode": (c_int, [vp, sz, POINTER(vp), POINTER(sz)]),
        "pk_np_decode": (c_int, [vp, sz, POINTER(vp), POINTER(sz)]),
        "pk_nps_read_r1cs": (c_int, [vp, sz, POINTER(vp)]),
        "pk_nps_r1cs": (POINTER(R1CS), [vp]),
        "pk_nps_num_public_inputs": (ctypes.c_int64, [vp]),
        "pk_nps_free": (None, [vp]),
        "pk_profile_begin": (c_int, [vp]),
        "pk_profile_end": (c_int, [vp, POINTER(c_double), POINTER(c_uint64), POINTER(c_double)]),
        "pk_prover_upload_inputs": (c_int, [vp, u64p, POINTER(Rand)]),
        "pk_prove_staged": (c_int, [vp, POINTER(vp), POINTER(sz)]),
        "pk_rng_fill": (c_int, [vp, vp, sz, sz, c_char_p, c_uint32]),
        "pk_prover_upload_inputs_seeded": (c_int, [vp, u64p, c_char_p]),
        "pk_prove_seeded": (c_int, [vp, u64p, c_char_p, POINTER(vp), POINTER(sz)]),
        "pk_modmul_bench": (c_int, [vp, sz, c_int, POINTER(c_float)]),
        "pk_modsqr_bench": (c_int, [vp, sz, c_int, POINTER(c_float)]),
    }
    missing = [n for n in sig if not hasattr(L, n)]
    if missing:
        raise ImportError(f"libpkwhir.so lacks symbols declared in include/pkwhir.h: {missing}")
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


EXPORTED = None  # filled lazily by tests via lib()
