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
DENSE_AUTO, DENSE_EXACT, DENSE_TENSOR = 0, 1, 2
E_INVALID, E_WORKSPACE, E_CUDA, E_UNSUPPORTED = -1, -2, -3, -4


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
    "b200rag_debug_scan_stats": (ctypes.c_int, [c_int32, c_void_p, c_int32]),
    "b200rag_debug_sparse_stats": (ctypes.c_int, [c_int32, c_void_p, c_int32]),
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
    "b200rag_merge_gathered": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "b200rag_rrf_fuse_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "b200rag_rrf_fuse": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
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


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200rag_last_error()
        raise B200RagError(rc, msg.decode() if msg else "")
