"""Restatement of the reference's rank fusion, MMR and learned re-rank on plain ids/floats.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Parity: PINNED -- tests/test_oracle_fusion.py checks every function
here against tests/golden/fusion_golden.json, which oracle/gen_golden.py produced by executing the
reference's own HybridRetriever._fuse_results / _mmr_diversify / rerank (imported unmodified through
oracle/ref_import.py).

All arithmetic is Python float (IEEE fp64) in the reference's operation order.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Sequence, Tuple


def rrf_fuse(lists: Sequence[Sequence[Hashable]], weights: Sequence[float], rrf_k: int = 60
             ) -> Tuple[List[Hashable], List[float], List[int]]:
    """Weighted Reciprocal Rank Fusion (reference src/advanced_rag/retrieval.py:421-487).

    lists   ranked id lists in the reference's processing order semantic, sparse[, domain]
    weights dense_weight, sparse_weight[, 0.2]  (retrieval.py:440,448,455)
    Per hit: fused[id] += (1.0 / (rrf_k + rank)) * w, rank starting at 1 (retrieval.py:437-440).
    Result order: score descending with a STABLE sort, so ties keep first-seen order (retrieval.py:487);
    first-seen order is dict insertion order over the lists as processed.
    Returns (ids, fused scores, method bitmask: bit i set if list i contained the id).
    """
    score: Dict[Hashable, float] = {}
    mask: Dict[Hashable, int] = {}
    for li, (ids, w) in enumerate(zip(lists, weights)):
        for rank, doc in enumerate(ids, start=1):
            contrib = (1.0 / (rrf_k + rank)) * w
            score[doc] = score.get(doc, 0.0) + contrib
            mask[doc] = mask.get(doc, 0) | (1 << li)
    order = sorted(score, key=score.__getitem__, reverse=True)
    return order, [score[d] for d in order], [mask[d] for d in order]


def jaccard(a: frozenset, b: frozenset) -> float:
    """Token-set Jaccard with the reference's `or 1` guard (retrieval.py:507)."""
    return len(a & b) / (len(a | b) or 1)


def tokens(content) -> frozenset:
    """The reference's tokeniser for MMR: set((content or "").lower().split()) (retrieval.py:497)."""
    return frozenset((content or "").lower().split())


def mmr_select(rel: Sequence[float], token_sets: Sequence[frozenset], k: int, mmr_lambda: float
               ) -> List[int]:
    """Greedy Maximal Marginal Relevance (reference retrieval.py:493-516).

    Candidates are given in fused order; returns the selected candidate indices in pick order.
    First pick: max rel.  Later picks: argmax of  mmr_lambda*rel - (1-mmr_lambda)*max_sim  with strict '>'
    (earliest candidate wins ties, retrieval.py:511), max_sim = max Jaccard against everything selected.
    The running maximum is kept incrementally (max is order independent, so this equals the reference's
    recomputation bit for bit).
    """
    n = len(rel)
    remaining = list(range(n))
    max_sim = [0.0] * n
    picked: List[int] = []
    while remaining and len(picked) < k:
        best = None
        best_score = -1e9
        for c in remaining:
            if not picked:
                s = rel[c]
            else:
                s = mmr_lambda * rel[c] - (1 - mmr_lambda) * max_sim[c]
            if s > best_score:
                best_score = s
                best = c
        picked.append(best)
        remaining.remove(best)
        tb = token_sets[best]
        for c in remaining:
            j = jaccard(token_sets[c], tb)
            if j > max_sim[c]:
                max_sim[c] = j
    return picked


def learned_rank(scores: Sequence[float], n_methods: Sequence[int], recency: Sequence[float],
                 top_k: int, base_weight: float = 1.0, method_bonus: float = 0.1,
                 recency_weight: float = 0.0) -> Tuple[List[int], List[float]]:
    """LearnedRanker.score + the sort/truncate of rerank (reference ranker.py:109-125,
    retrieval.py:555-563).  Returns (indices into the input in new order, their rerank scores)."""
    rs = [base_weight * float(s) + method_bonus * float(m) + recency_weight * float(r)
          for s, m, r in zip(scores, n_methods, recency)]
    order = sorted(range(len(rs)), key=rs.__getitem__, reverse=True)[:top_k]
    return order, [rs[i] for i in order]


def pairwise_similarity(token_sets: Sequence[frozenset]) -> Tuple[float, int]:
    """RAGEvaluator._calculate_pairwise_similarity (reference evaluation.py:327-344): mean token-set Jaccard over the pairs
    i < j whose two sets are non-empty, 0.0 when fewer than two results or no pair qualifies.  The mean is numpy's
    (np.mean of a Python list), as in the reference.  Returns (mean, number of pairs averaged)."""
    import numpy as np
    if len(token_sets) < 2:
        return 0.0, 0
    sims = []
    for i in range(len(token_sets)):
        for j in range(i + 1, len(token_sets)):
            a, b = token_sets[i], token_sets[j]
            if a and b:
                sims.append(len(a & b) / len(a | b))
    return (float(np.mean(sims)) if sims else 0.0), len(sims)
