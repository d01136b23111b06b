"""ctypes binding of the C-ABI library (include/swr_b200.h).

The CUDA extension is the product: there is no CPU or eager-PyTorch fallback.  Every
entry point raises ``RuntimeError`` when ``libswr_b200.so`` is missing (run
``python scenario-wise-rec_b200/build.py``) or when a call returns a non-zero status.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswr_b200.so")

ABI_VERSION = 2
REC_INTS, REC_FLOATS, REC_SLOTS = 32, 8, 32
REC_DTYPE = np.dtype([("kind", "<i4"), ("n_sub", "<i4"), ("i", "<i4", (REC_INTS,)),
                      ("f", "<f4", (REC_FLOATS,)), ("s", "<i4", (REC_SLOTS,))])
assert REC_DTYPE.itemsize == 8 + 4 * (REC_INTS + REC_FLOATS + REC_SLOTS)

# swr_op_kind
OP_ZERO, OP_GATHER, OP_SCATTER, OP_COLSTATS = 1, 2, 3, 4
OP_FC_FWD, OP_FC_DGRAD, OP_FC_WGRAD = 5, 6, 7
OP_POOL_FWD, OP_POOL_BWD, OP_HEAD_FWD, OP_HEAD_BWD = 8, 9, 10, 11
OP_BN_UPDATE, OP_BN_PGRAD, OP_GROUP = 12, 13, 100
OP_EW_FWD, OP_EW_BWD, OP_SUMGRAD, OP_SELECT_FWD, OP_SELECT_BWD = 14, 15, 16, 17, 18
OP_LN_FWD, OP_LN_BWD, OP_MIX_FWD, OP_MIX_BWD, OP_BMV_FWD, OP_BMV_BWD = 19, 20, 21, 22, 23, 24
OP_BCE, OP_ADAM, OP_FC_PRESPLIT, OP_ADAM_ROWS, OP_ADAM_FLUSH = 25, 26, 27, 28, 29
EW_MUL, EW_ADD, EW_COPY = 0, 1, 2
HEAD_SIG_SELECT_ADD, HEAD_SELECT_SIG, HEAD_NO_SELECT = 0, 1, 2
NORM_NONE, NORM_BATCH, NORM_RUNNING = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_LEAKY = 0, 1, 2, 3
W_NK, W_KN = 0, 1
# swr_dtype
DT_I8, DT_I16, DT_I32, DT_I64, DT_U8, DT_F16, DT_BF16, DT_F32, DT_F64 = 0, 1, 2, 3, 4, 8, 9, 10, 11

FC_SIMT, FC_TC, FC_AUTO = 0, 1, 2   # swr_fc_mode

EXPORTS = ("swr_abi_version", "swr_last_error", "swr_launch_count", "swr_device_check", "swr_set_fc_mode", "swr_get_fc_mode",
           "swr_profile_begin", "swr_profile_end", "swr_memcpy_async",
           "swr_embedding_gather_fwd", "swr_embedding_scatter_bwd", "swr_program_run",
           "swr_peer_alloc", "swr_peer_free", "swr_peer_handle", "swr_peer_open", "swr_peer_close")

_lib = None


def lib():
    """Load (once) and return the C-ABI library; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA extension has not been built "
            "(python scenario-wise-rec_b200/build.py). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.swr_abi_version.restype = ctypes.c_int
    L.swr_last_error.restype = ctypes.c_char_p
    L.swr_launch_count.restype = ctypes.c_int64
    L.swr_device_check.restype = ctypes.c_int
    L.swr_set_fc_mode.restype = ctypes.c_int
    L.swr_set_fc_mode.argtypes = [ctypes.c_int]
    L.swr_get_fc_mode.restype = ctypes.c_int
    L.swr_program_run.restype = ctypes.c_int
    L.swr_program_run.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.swr_embedding_gather_fwd.restype = ctypes.c_int
    L.swr_embedding_gather_fwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, vp, vp]
    L.swr_embedding_scatter_bwd.restype = ctypes.c_int
    L.swr_embedding_scatter_bwd.argtypes = [vp, i64, i64, vp, vp, vp, vp, i32, i32, vp]
    L.swr_memcpy_async.restype = ctypes.c_int
    L.swr_memcpy_async.argtypes = [vp, vp, i64, vp]
    L.swr_profile_begin.restype = ctypes.c_int
    L.swr_profile_end.restype = ctypes.c_int
    L.swr_profile_end.argtypes = [vp, vp, vp, i32]
    for fn in (L.swr_peer_alloc, L.swr_peer_free, L.swr_peer_handle, L.swr_peer_open, L.swr_peer_close):
        fn.restype = ctypes.c_int
    L.swr_peer_alloc.argtypes = [i64, vp]
    L.swr_peer_free.argtypes = [vp]
    L.swr_peer_handle.argtypes = [vp, vp]
    L.swr_peer_open.argtypes = [vp, vp]
    L.swr_peer_close.argtypes = [vp]
    if L.swr_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libswr_b200.so ABI {L.swr_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = L
    return L


def last_error() -> str:
    return lib().swr_last_error().decode()


def check(status: int, what: str):
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")


def launch_count() -> int:
    return int(lib().swr_launch_count())


def set_fc_mode(mode: int) -> int:
    """Select the arithmetic of the grouped FC ops (FC_SIMT / FC_TC / FC_AUTO); returns the previous mode.
    Programs already captured into a CUDA graph keep the kernels they were captured with."""
    return int(lib().swr_set_fc_mode(int(mode)))


def get_fc_mode() -> int:
    return int(lib().swr_get_fc_mode())


def program_run(recs: np.ndarray, slots: np.ndarray, stream: int):
    """recs: REC_DTYPE array; slots: uint64 array of device pointers; stream: cudaStream_t handle."""
    assert recs.dtype == REC_DTYPE and recs.flags.c_contiguous
    assert slots.dtype == np.uint64 and slots.flags.c_contiguous
    st = lib().swr_program_run(recs.ctypes.data, recs.shape[0], slots.ctypes.data, slots.shape[0], stream)
    check(st, "swr_program_run")


def memcpy_async(dst: int, src: int, nbytes: int, stream: int):
    check(lib().swr_memcpy_async(dst, src, nbytes, stream), "swr_memcpy_async")


def profile_begin():
    check(lib().swr_profile_begin(), "swr_profile_begin")


def profile_end(cap: int = 65536):
    """-> (op kinds, header record indices, device ms) of every op run since profile_begin()."""
    kinds = np.zeros(cap, dtype=np.int32)
    recs = np.zeros(cap, dtype=np.int32)
    ms = np.zeros(cap, dtype=np.float32)
    n = lib().swr_profile_end(kinds.ctypes.data, recs.ctypes.data, ms.ctypes.data, cap)
    if n < 0:
        check(n, "swr_profile_end")
    n = min(n, cap)
    return kinds[:n], recs[:n], ms[:n]


def torch_dtype_code(dt) -> int:
    import torch
    table = {torch.int8: DT_I8, torch.int16: DT_I16, torch.int32: DT_I32, torch.int64: DT_I64, torch.uint8: DT_U8,
             torch.float16: DT_F16, torch.bfloat16: DT_BF16, torch.float32: DT_F32, torch.float64: DT_F64,
             torch.bool: DT_U8}
    if dt not in table:
        raise TypeError(f"unsupported feature column dtype {dt}")
    return table[dt]
