"""BM25 weighting and the sparse encoder that feeds the sparse index.

The reference never computes BM25: its sparse path is an inner product of sparse vectors whose values come from a
user-supplied `embedding_generator.encode_sparse` or a random placeholder (reference
src/advanced_rag/indexing.py:629-654).  This module defines the weights this repo uses:

    idf(t)  = ln(1 + (N - df(t) + 0.5) / (df(t) + 0.5))
    w(t, d) = idf(t) * tf * (k1 + 1) / (tf + k1 * (1 - b + b * len(d) / avgdl)),   k1 = 1.2, b = 0.75

evaluated in fp64 in exactly this operation order and rounded once to fp32; a query holds value 1.0 per unique
term, so the sparse inner product equals the BM25 score.  Tokenisation is `text.lower().split()`, the same
tokeniser the reference's MMR uses (retrieval.py:497), so the BM25 vocabulary and the MMR token sets coincide.
Ingest-side host code (numpy); the search itself runs in sparse_bm25.cu.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

K1 = 1.2
B = 0.75


def bm25_idf(n_docs: int, df: np.ndarray) -> np.ndarray:
    df = df.astype(np.float64)
    return np.log(1.0 + (n_docs - df + 0.5) / (df + 0.5))


def bm25_weights(doc_ptr: np.ndarray, term_ids: np.ndarray, tf: np.ndarray, n_terms: int,
                 k1: float = K1, b: float = B, stats=None) -> np.ndarray:
    """Doc-major CSR of term frequencies -> fp32 BM25 weights (same nnz order).  stats = (n_docs, df, total_length) of the
    WHOLE corpus when this CSR is one row shard of it (distributed.bm25_global_stats): idf and avgdl are global statistics."""
    doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
    term_ids = np.asarray(term_ids, dtype=np.int64)
    tf = np.asarray(tf, dtype=np.float64)
    n_docs = doc_ptr.shape[0] - 1
    if term_ids.size == 0:
        return np.zeros(0, dtype=np.float32)
    counts = np.diff(doc_ptr)
    doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), counts)
    doc_len = np.bincount(doc_of, weights=tf, minlength=n_docs)          # sum of tf, exact in fp64
    if stats is None:
        n_all, df, total_len = n_docs, np.bincount(term_ids, minlength=n_terms), doc_len.sum()
    else:
        n_all, df, total_len = int(stats[0]), np.asarray(stats[1].cpu() if hasattr(stats[1], "cpu") else stats[1]), float(stats[2])
    avgdl = total_len / n_all
    idf = bm25_idf(n_all, df)
    norm = k1 * (1.0 - b + b * doc_len[doc_of] / avgdl)
    w = idf[term_ids] * tf * (k1 + 1.0) / (tf + norm)
    return w.astype(np.float32)


def bm25_weights_device(doc_ptr, term_ids, tf, n_terms: int, k1: float = K1, b: float = B, stats=None):
    """bm25_weights for CSR arrays that already live on a CUDA device (torch tensors): same fp64 operation order, so
    the fp32 result is bit-identical to the numpy version (idf, the only transcendental, is evaluated on the host for
    the n_terms vocabulary entries; +, *, / are correctly rounded on both sides)."""
    import torch
    dev = term_ids.device
    n_docs = doc_ptr.numel() - 1
    if term_ids.numel() == 0:
        return torch.zeros(0, dtype=torch.float32, device=dev)
    counts = (doc_ptr[1:] - doc_ptr[:-1]).to(dev)
    doc_of = torch.repeat_interleave(torch.arange(n_docs, device=dev), counts, output_size=term_ids.numel())
    tf64 = tf.to(torch.float64)
    doc_len = torch.zeros(n_docs, dtype=torch.float64, device=dev).index_add_(0, doc_of, tf64)   # integers: exact in fp64
    if stats is None:
        n_all, df, total_len = n_docs, torch.bincount(term_ids, minlength=n_terms).cpu().numpy(), float(doc_len.sum().item())
    else:                                  # one row shard of a larger corpus: global n_docs / df / total length
        n_all, df, total_len = int(stats[0]), np.asarray(stats[1].cpu() if hasattr(stats[1], "cpu") else stats[1]), float(stats[2])
    avgdl = total_len / n_all
    idf = torch.as_tensor(bm25_idf(n_all, df)).to(dev)
    norm = k1 * (1.0 - b + b * doc_len[doc_of] / avgdl)
    w = idf[term_ids] * tf64 * (k1 + 1.0) / (tf64 + norm)
    return w.to(torch.float32)


def tokenize(text) -> List[str]:
    return (text or "").lower().split()


class Bm25Encoder:
    """Vocabulary + encode_sparse hook with the reference's generator interface (indexing.py:634-643)."""

    def __init__(self):
        self.vocab: Dict[str, int] = {}

    def fit_transform(self, texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """texts -> doc-major CSR (doc_ptr, term_ids ascending per doc, tf); grows the vocabulary."""
        ptr = [0]
        terms: List[int] = []
        tfs: List[int] = []
        for text in texts:
            cnt: Dict[int, int] = {}
            for tok in tokenize(text):
                tid = self.vocab.setdefault(tok, len(self.vocab))
                cnt[tid] = cnt.get(tid, 0) + 1
            for tid in sorted(cnt):
                terms.append(tid)
                tfs.append(cnt[tid])
            ptr.append(len(terms))
        return (np.asarray(ptr, dtype=np.int64), np.asarray(terms, dtype=np.int64), np.asarray(tfs, dtype=np.int64))

    def encode_sparse(self, text: str) -> Dict[str, list]:
        ids = sorted({self.vocab[t] for t in tokenize(text) if t in self.vocab})
        return {"indices": ids, "values": [1.0] * len(ids)}

    def token_sets(self, texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
        """texts -> CSR of sorted unique token ids (MMR token sets); unseen tokens are added to the vocabulary."""
        ptr = [0]
        ids: List[int] = []
        for text in texts:
            s = sorted({self.vocab.setdefault(t, len(self.vocab)) for t in tokenize(text)})
            ids.extend(s)
            ptr.append(len(ids))
        return np.asarray(ptr, dtype=np.int64), np.asarray(ids, dtype=np.int32)
