"""ctypes binding of libtq100.so (include/tq100.h).  There is no CPU fallback: if the library is
missing or a call fails, the caller gets a RuntimeError carrying tq_last_error_string()."""

import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_PKG, "libtq100.so")

# enums of include/tq100.h
F32, F16, BF16 = 0, 1, 2
HESS_AUTO, HESS_FFMA, HESS_TCGEN05 = 0, 1, 2
AGA_NONE, AGA_HESSIAN, AGA_ACTIVATIONS = 0, 1, 2
ORDER_SEQUENTIAL, ORDER_SSR, ORDER_STATIC = 0, 1, 2
SWEEP_ROW_SHARD = 1
SWEEP_FFMA_FEEDBACK = 2
SWEEP_UNFUSED_STATS = 4
OP_INIT, OP_GRID, OP_ROUND, OP_ITF, OP_AGA = 0, 1, 2, 3, 4

_i64 = ctypes.c_int64
_int = ctypes.c_int
_ptr = ctypes.c_void_p
_dbl = ctypes.c_double

_SIGNATURES = {
    "tq_abi_version": (_int, []),
    "tq_last_error_string": (ctypes.c_char_p, []),
    "tq_launch_count": (ctypes.c_longlong, []),
    "tq_hessian_accum": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _int, _int, _ptr]),
    "tq_symmetrize": (_int, [_ptr, _i64, _i64, _ptr]),
    "tq_debug_set_diag_prof": (None, [_ptr]),
    "tq_hessian_finalize": (_int, [_ptr, _ptr, _i64, _dbl, _dbl, _ptr, _ptr]),
    "tq_chol_workspace_floats": (_i64, [_i64]),
    "tq_chol_inverse": (_int, [_ptr, _ptr, _i64, _ptr, _ptr, _ptr]),
    "tq_ssr_num_chunks": (_i64, [_i64]),
    "tq_ssr_stats": (_int, [_ptr, _i64, _i64, _ptr, _i64, _ptr, _ptr, _ptr]),
    "tq_ssr_select": (_int, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "tq_aga_vector": (_int, [_ptr, _i64, _ptr, _i64, _i64, _int, _ptr, _ptr]),
    "tq_atq_block": (_int, [_ptr, _i64, _i64, _ptr, _i64, _i64, _ptr, _int, _ptr, _i64, _ptr, _ptr, _i64,
                            _ptr, _i64, _ptr, _ptr]),
    "tq_atq_stage": (_int, [_int, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _int, _ptr, _ptr, _ptr, _ptr]),
    "tq_err_feedback": (_int, [_ptr, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _ptr, _i64, _i64, _ptr]),
    "tq_err_feedback_tc_workspace_floats": (_i64, [_i64, _i64, _i64]),
    "tq_err_feedback_tc": (_int, [_ptr, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _ptr, _i64, _i64, _ptr, _ptr]),
    "tq_split_tf32": (_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _i64, _ptr]),
    "tq_unpermute_codes": (_int, [_ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "tq_dequant": (_int, [_ptr, _ptr, _i64, _ptr, _i64, _i64, _ptr, _i64, _ptr, _ptr]),
    "tq_pack2b": (_int, [_ptr, _i64, _ptr, _ptr]),
    "tq_pack2b_f32": (_int, [_ptr, _i64, _ptr, _ptr]),
    "tq_unpack2b": (_int, [_ptr, _i64, _ptr, _ptr]),
    "tq_sweep_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "tq_sweep_layer": (_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _i64, _int, _int, _int, _ptr, _ptr,
                              _ptr, _ptr, _ptr, _ptr, _i64, _int, _ptr]),
    "tq_tl_words_per_row": (_i64, [_i64]),
    "tq_tl_pack": (_int, [_ptr, _i64, _i64, _ptr, _ptr, _i64, _ptr]),
    "tq_tl_wtab": (_int, [_ptr, _ptr, _i64, _i64, _int, _ptr, _ptr]),
    "tq_tl_gemv": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _ptr, _int, _i64, _i64, _ptr, _ptr, _ptr, _i64, _ptr]),
    "tq_tl_gemm_tc": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _ptr, _int, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "tq_tl_dequant": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _int, _i64, _ptr]),
    "tq_tl_unpack": (_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "tq_comm_unique_id": (_int, [_ptr]),
    "tq_comm_init": (_int, [_ptr, _int, _int]),
    "tq_comm_ready": (_int, []),
    "tq_comm_destroy": (_int, []),
    "tq_comm_allreduce_f32": (_int, [_ptr, _i64, _ptr]),
    "tq_comm_p2p_alloc": (_int, [_i64, _int, _int, _ptr]),
    "tq_comm_p2p_open": (_int, [_ptr]),
    "tq_comm_p2p_ready": (_int, []),
}

EXPORTS = tuple(_SIGNATURES)
_lib = None


def lib_path():
    return _LIB_PATH


def load():
    """Load (once) and return the ctypes handle; raise loudly when the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"libtq100.so not found at {_LIB_PATH}: the CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root). "
            "This package has no CPU fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.tq_abi_version() != 1:
        raise RuntimeError(f"libtq100.so ABI version {lib.tq_abi_version()} != 1")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().tq_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (this package runs on the GPU only; no CPU fallback)")


def ptr(t):
    return None if t is None else _ptr(t.data_ptr())


def stream():
    return _ptr(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(load().tq_launch_count())


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.float16:
        return F16
    if dt == torch.bfloat16:
        return BF16
    raise RuntimeError(f"unsupported activation dtype {dt}: use float32, float16 or bfloat16")


def comm_ready():
    """True once the in-library NCCL communicator (row-sharded SSR statistics) is initialised."""
    return load().tq_comm_ready() > 1
