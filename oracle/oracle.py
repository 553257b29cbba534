"""ctypes front-end of oracle/exact_scan.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
See the header of exact_scan.c for the parity status ("parity unpinned" for the dense / sparse scoring
that the reference delegates to a Milvus server; pinned for RRF / MMR through oracle/fusion.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
F16, BF16 = 0, 1
_lib: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> str:
    """Compile exact_scan.c with the committed Makefile (gcc only, no reference sources involved)."""
    src = os.path.join(_HERE, "exact_scan.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        c_i64, c_i32, c_int = ctypes.c_int64, ctypes.c_int32, ctypes.c_int
        P = ctypes.c_void_p
        L.orc_bits_to_double.restype = ctypes.c_double
        L.orc_bits_to_double.argtypes = [ctypes.c_uint16, c_int]
        L.orc_double_to_bits.restype = ctypes.c_uint16
        L.orc_double_to_bits.argtypes = [ctypes.c_double, c_int]
        L.orc_round_f32.argtypes = [P, P, c_i64, c_int]
        L.orc_bits_to_f32.argtypes = [P, P, c_i64, c_int]
        L.orc_normalize_rows.argtypes = [P, P, c_i64, c_int, c_int]
        L.orc_dense_scores.argtypes = [P, c_i64, c_int, c_int, P, P]
        L.orc_dense_topk.argtypes = [P, c_i64, c_int, c_int, P, c_int, c_int, c_i64, P, P]
        L.orc_dense_topk.restype = c_int
        L.orc_sparse_topk.argtypes = [P, P, P, c_i64, c_i32, P, P, P, c_int, c_int, c_i64, P, P, P]
        L.orc_sparse_topk.restype = c_int
        L.orc_merge_topk.argtypes = [P, P, c_int, c_int, c_int, P, P]
        L.orc_merge_topk.restype = c_int
        L.orc_num_threads.restype = c_int
        _lib = L
    return _lib


def _p(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def round_f32(x: np.ndarray, dtype: int) -> np.ndarray:
    """fp32 -> 16-bit patterns (uint16) of fp16 / bf16, round to nearest even."""
    x = _c(x, np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().orc_round_f32(_p(x), _p(out), x.size, dtype)
    return out


def bits_to_f32(bits: np.ndarray, dtype: int) -> np.ndarray:
    bits = _c(bits, np.uint16)
    out = np.empty(bits.shape, dtype=np.float32)
    lib().orc_bits_to_f32(_p(bits), _p(out), bits.size, dtype)
    return out


def normalize_rows(x: np.ndarray, dtype: int) -> np.ndarray:
    """Canonical cosine preparation: fp32 rows -> L2-normalised 16-bit patterns."""
    x = _c(x, np.float32)
    assert x.ndim == 2
    out = np.empty(x.shape, dtype=np.uint16)
    lib().orc_normalize_rows(_p(x), _p(out), x.shape[0], x.shape[1], dtype)
    return out


def dense_scores(corpus_bits: np.ndarray, query_bits: np.ndarray, dtype: int) -> np.ndarray:
    corpus_bits = _c(corpus_bits, np.uint16)
    query_bits = _c(query_bits, np.uint16)
    n, d = corpus_bits.shape
    out = np.empty(n, dtype=np.float64)
    lib().orc_dense_scores(_p(corpus_bits), n, d, dtype, _p(query_bits), _p(out))
    return out


def dense_topk(corpus_bits: np.ndarray, query_bits: np.ndarray, k: int, dtype: int,
               id_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Exact flat top-k; returns (scores f64 [B,k], ids i64 [B,k]) ordered (score desc, id asc)."""
    corpus_bits = _c(corpus_bits, np.uint16)
    query_bits = _c(query_bits, np.uint16)
    n, d = corpus_bits.shape
    b = query_bits.shape[0]
    assert query_bits.shape[1] == d
    scores = np.empty((b, k), dtype=np.float64)
    ids = np.empty((b, k), dtype=np.int64)
    rc = lib().orc_dense_topk(_p(corpus_bits), n, d, dtype, _p(query_bits), b, k, id_offset,
                              _p(scores), _p(ids))
    if rc != 0:
        raise ValueError(f"orc_dense_topk rc={rc}")
    return scores, ids


def sparse_topk(term_ptr, post_doc, post_w, n_docs: int, q_ptr, q_terms, q_vals, k: int,
                id_offset: int = 0):
    term_ptr = _c(term_ptr, np.int64)
    post_doc = _c(post_doc, np.int32)
    post_w = _c(post_w, np.float32)
    q_ptr = _c(q_ptr, np.int64)
    q_terms = _c(q_terms, np.int32)
    q_vals = _c(q_vals, np.float32)
    b = q_ptr.shape[0] - 1
    scores = np.empty((b, k), dtype=np.float32)
    ids = np.empty((b, k), dtype=np.int64)
    counts = np.empty(b, dtype=np.int32)
    rc = lib().orc_sparse_topk(_p(term_ptr), _p(post_doc), _p(post_w), n_docs, term_ptr.shape[0] - 1,
                               _p(q_ptr), _p(q_terms), _p(q_vals), b, k, id_offset,
                               _p(scores), _p(ids), _p(counts))
    if rc != 0:
        raise ValueError(f"orc_sparse_topk rc={rc}")
    return scores, ids, counts


def merge_topk(cand_scores, cand_ids, k: int):
    cand_scores = _c(cand_scores, np.float64)
    cand_ids = _c(cand_ids, np.int64)
    b, m = cand_ids.shape
    scores = np.empty((b, k), dtype=np.float64)
    ids = np.empty((b, k), dtype=np.int64)
    lib().orc_merge_topk(_p(cand_scores), _p(cand_ids), b, m, k, _p(scores), _p(ids))
    return scores, ids


# ---------------------------------------------------------------------------------------------
# "Fast CPU" variants: what a CPU deployment of the same exact search would run (BLAS sgemm + partial
# sort, scipy CSR product).  Used ONLY as the timed cpu_baseline in bench.py; never as a checker.
# ---------------------------------------------------------------------------------------------

def dense_topk_blas(corpus_f32: np.ndarray, queries_f32: np.ndarray, k: int):
    s = queries_f32 @ corpus_f32.T
    k = min(k, s.shape[1])
    part = np.argpartition(-s, k - 1, axis=1)[:, :k]
    ps = np.take_along_axis(s, part, axis=1)
    order = np.lexsort((part, -ps), axis=1)
    ids = np.take_along_axis(part, order, axis=1)
    return np.take_along_axis(ps, order, axis=1), ids.astype(np.int64)
