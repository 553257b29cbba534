"""ctypes binding of libb200rag.so (the C ABI declared in include/b200rag.h).

The library is the product: there is no Python or CPU fallback.  Loading fails loudly when the shared object has
not been built (python advanced-rag-milvus_b200/build.py) and every call raises B200RagError on a non-zero status.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200rag.so")

F16, BF16 = 0, 1
DENSE_AUTO, DENSE_EXACT, DENSE_TENSOR, DENSE_APPROX = 0, 1, 2, 3
E_INVALID, E_WORKSPACE, E_CUDA, E_UNSUPPORTED = -1, -2, -3, -4
OP_EQ, OP_NE, OP_GE, OP_LE, OP_GT, OP_LT = range(6)
COL_F64, COL_I64, COL_I64_AS_F64, COL_CODE, COL_NEVER = range(5)


class FilterTerm(ctypes.Structure):
    """b200rag_filter_term (include/b200rag.h)."""
    _fields_ = [("column", c_void_p), ("lut", c_void_p), ("fvalue", c_double), ("ivalue", c_int64),
                ("kind", c_int32), ("op", c_int32), ("lut_size", c_int32), ("reserved", c_int32)]


class B200RagError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"b200rag error {code}: {message}")
        self.code = code


# name -> (restype, argtypes); must list every symbol include/b200rag.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "b200rag_last_error": (ctypes.c_char_p, []),
    "b200rag_abi_version": (ctypes.c_int, []),
    "b200rag_kernel_launch_count": (ctypes.c_uint64, []),
    "b200rag_device_info": (ctypes.c_int, [c_void_p, c_void_p, c_void_p]),
    "b200rag_prepare_rows": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "b200rag_dense_topk_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32, c_int32]),
    "b200rag_dense_topk": (ctypes.c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int64,
                                          c_void_p, c_void_p, c_void_p, c_double, c_void_p,
                                          c_void_p, c_size_t, c_int32, c_void_p]),
    "b200rag_dense_topk_masked": (ctypes.c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int64,
                                                 c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p,
                                                 c_void_p, c_size_t, c_int32, c_void_p]),
    "b200rag_profile_next_scan": (ctypes.c_int, [c_void_p, c_void_p]),
    "b200rag_debug_set_stats_buffer": (ctypes.c_int, [c_int32, c_void_p, c_size_t]),
    "b200rag_set_option": (ctypes.c_int, [ctypes.c_char_p, c_int64]),
    "b200rag_get_option": (c_int64, [ctypes.c_char_p]),
    "b200rag_filter_mask": (ctypes.c_int, [c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200rag_sparse_topk_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "b200rag_sparse_topk": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                           c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200rag_sparse_topk_masked": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                                  c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64,
                                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200rag_merge_topk_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "b200rag_merge_topk": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                          c_void_p, c_size_t, c_void_p]),
    "b200rag_merge_gathered": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200rag_rrf_fuse_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "b200rag_rrf_fuse": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200rag_fuse_select": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p]),
    "b200rag_rerank_learned": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_double, c_double, c_double,
                                              c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200rag_pairwise_jaccard_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "b200rag_pairwise_jaccard": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_void_p, c_size_t, c_void_p]),
    "b200rag_mmr_select_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "b200rag_mmr_select": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                          c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libb200rag.so.  Raises ImportError (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension has not been built. "
                "Run `python advanced-rag-milvus_b200/build.py` (there is no CPU fallback).")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def set_option(name: str, value: int) -> None:
    """A/B knob of the library (include/b200rag.h: b200rag_set_option); value < 0 restores the default."""
    check(load().b200rag_set_option(name.encode(), int(value)))


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200rag_last_error()
        raise B200RagError(rc, msg.decode() if msg else "")
