"""Restatement of the reference's whole retrieval chain on plain arrays -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

    profile -> dense top-2k + sparse top-2k (+ domain top-k) -> weighted RRF -> optional MMR -> [:top_k] -> learned re-rank

following reference src/advanced_rag/retrieval.py:249-339 (_retrieve_inner), :341-419 (K multipliers per index),
:421-491 (fusion), :493-516 (MMR), :518-563 + ranker.py:109-125 (re-rank).  Parity: PINNED by
tests/test_oracle.py::test_pipeline_restatement_matches_reference_e2e_golden against tests/golden/e2e_golden.json,
which oracle/gen_e2e_golden.py produced by running the unmodified reference over oracle/inmem_index.py.
Search scores come from oracle/exact_scan.c (see its header for the "parity unpinned" status of those two stages).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import fusion, oracle

# reference retrieval.py:142-213 for a base config of top_k=20, rerank_top_k=5 (values pinned by fusion_golden.json "profiles")
PROFILES_TOPK20 = {
    "default": dict(top_k=20, enable_mmr=False, mmr_lambda=0.7),
    "faq": dict(top_k=10, enable_mmr=False, mmr_lambda=0.7),
    "troubleshooting": dict(top_k=30, enable_mmr=True, mmr_lambda=0.5),
    "summary": dict(top_k=40, enable_mmr=False, mmr_lambda=0.7),
    "analysis": dict(top_k=30, enable_mmr=True, mmr_lambda=0.8),
}


class ArrayCorpus:
    """Stored (rounded, normalised) dense rows + term-major postings + token sets, all host numpy."""

    def __init__(self, semantic_f32: np.ndarray, domain_f32: Optional[np.ndarray], sp_ptr, sp_idx, sp_val, sparse_dim: int,
                 contents: Optional[Sequence[str]], dtype: int = oracle.F16):
        self.dtype = dtype
        self.n = semantic_f32.shape[0]
        self.sem = oracle.normalize_rows(semantic_f32, dtype)
        self.dom = oracle.normalize_rows(domain_f32, dtype) if domain_f32 is not None else None
        sp_ptr = np.asarray(sp_ptr, dtype=np.int64)
        sp_idx = np.asarray(sp_idx, dtype=np.int64)
        doc_of = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(sp_ptr))
        order = np.lexsort((doc_of, sp_idx))
        self.term_ptr = np.zeros(sparse_dim + 1, dtype=np.int64)
        np.cumsum(np.bincount(sp_idx, minlength=sparse_dim), out=self.term_ptr[1:])
        self.post_doc = doc_of[order].astype(np.int32)
        self.post_w = np.asarray(sp_val, dtype=np.float32)[order]
        self.tokens = [fusion.tokens(c) for c in contents] if contents is not None else None


def retrieve(c: ArrayCorpus, sem_q: np.ndarray, sparse_q: Dict[str, list], dom_q: Optional[np.ndarray], top_k: int,
             dense_weight: float = 0.7, sparse_weight: float = 0.3, enable_mmr: bool = False, mmr_lambda: float = 0.7,
             with_sparse: bool = True) -> Tuple[List[int], List[float], List[int]]:
    """One query through the chain.  Returns (row ids, fused scores, method bitmasks) of the final <= top_k hits."""
    q = oracle.normalize_rows(np.asarray(sem_q, dtype=np.float32).reshape(1, -1), c.dtype)
    _, si = oracle.dense_topk(c.sem, q, 2 * top_k, c.dtype)
    lists = [[int(i) for i in si[0] if i >= 0]]
    if with_sparse:
        qi = np.asarray(sparse_q["indices"], dtype=np.int32)
        qv = np.asarray(sparse_q["values"], dtype=np.float32)
        o = np.argsort(qi, kind="stable")
        _, pi, pc = oracle.sparse_topk(c.term_ptr, c.post_doc, c.post_w, c.n, np.asarray([0, qi.size], dtype=np.int64),
                                       qi[o], qv[o], 2 * top_k)
        lists.append([int(i) for i in pi[0, : int(pc[0])]])
    else:
        lists.append([])
    weights = [dense_weight, sparse_weight]
    if dom_q is not None:
        dq = oracle.normalize_rows(np.asarray(dom_q, dtype=np.float32).reshape(1, -1), c.dtype)
        _, di = oracle.dense_topk(c.dom, dq, top_k, c.dtype)
        lists.append([int(i) for i in di[0] if i >= 0])
        weights.append(0.2)
    ids, scores, masks = fusion.rrf_fuse(lists, weights)
    if enable_mmr and ids:
        picks = fusion.mmr_select(scores, [c.tokens[i] for i in ids], top_k, mmr_lambda)
        ids, scores, masks = [ids[p] for p in picks], [scores[p] for p in picks], [masks[p] for p in picks]
    return ids[:top_k], scores[:top_k], masks[:top_k]
