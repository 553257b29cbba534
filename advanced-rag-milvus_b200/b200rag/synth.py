"""Seeded synthetic inputs shared by the tests and bench.py.

Dense: standard-normal rows (the index normalises them when metric=COSINE).  Sparse: documents of Poisson(mean_len)
tokens drawn from a Zipf(s) vocabulary; queries of n_terms unique terms drawn from the same law without the
`skip_top` most frequent ranks (stop-word analogue), value 1.0.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def dense_rows(n: int, dim: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, dim), dtype=np.float32)


def _zipf_p(vocab: int, s: float) -> np.ndarray:
    p = 1.0 / np.arange(1, vocab + 1, dtype=np.float64) ** s
    return p / p.sum()


def zipf_corpus(n_docs: int, vocab: int, seed: int, mean_len: int = 128, s: float = 1.07
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> doc-major CSR (doc_ptr i64, term_ids i64 ascending per doc, tf i64)."""
    rng = np.random.default_rng(seed)
    lens = rng.poisson(mean_len, size=n_docs).astype(np.int64)
    total = int(lens.sum())
    cdf = np.cumsum(_zipf_p(vocab, s))
    toks = np.searchsorted(cdf, rng.random(total), side="right").astype(np.int64)
    np.minimum(toks, vocab - 1, out=toks)
    doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), lens)
    key, tf = np.unique(doc_of * vocab + toks, return_counts=True)
    d = key // vocab
    doc_ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(np.bincount(d, minlength=n_docs), out=doc_ptr[1:])
    return doc_ptr, (key % vocab).astype(np.int64), tf.astype(np.int64)


def zipf_corpus_device(n_docs: int, vocab: int, seed: int, device, mean_len: int = 128, s: float = 1.07):
    """zipf_corpus on a CUDA device (same law, different random stream), for corpora too large to build on the host in a
    test: -> doc-major CSR as device tensors (doc_ptr i64, term_ids i64 ascending per doc, tf i64)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    lens = torch.poisson(torch.full((n_docs,), float(mean_len), device=dev), generator=g).to(torch.int64)
    p = 1.0 / torch.arange(1, vocab + 1, dtype=torch.float64, device=dev) ** s
    cdf = torch.cumsum(p / p.sum(), 0)
    keys = []
    step = 100_000
    for d0 in range(0, n_docs, step):
        ln = lens[d0:d0 + step]
        tot = int(ln.sum())
        toks = torch.searchsorted(cdf, torch.rand(tot, generator=g, device=dev, dtype=torch.float64), right=True).clamp_(max=vocab - 1)
        doc_of = torch.repeat_interleave(torch.arange(d0, d0 + ln.numel(), device=dev), ln, output_size=tot)
        keys.append(doc_of * vocab + toks)
    key, tf = torch.unique(torch.cat(keys), sorted=True, return_counts=True)
    doc_ptr = torch.zeros(n_docs + 1, dtype=torch.int64, device=dev)
    doc_ptr[1:] = torch.cumsum(torch.bincount(key // vocab, minlength=n_docs), 0)
    return doc_ptr, key % vocab, tf


def zipf_queries(n_q: int, vocab: int, seed: int, n_terms: int = 8, skip_top: int = 100, s: float = 1.07
                 ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> CSR queries (q_ptr i64, q_terms i32 ascending per query, q_vals f32 = 1)."""
    rng = np.random.default_rng(seed)
    skip = min(skip_top, max(0, vocab - n_terms))
    p = _zipf_p(vocab, s)[skip:]
    p = p / p.sum()
    terms = []
    ptr = [0]
    for _ in range(n_q):
        t = np.sort(rng.choice(vocab - skip, size=min(n_terms, vocab - skip), replace=False, p=p) + skip)
        terms.append(t)
        ptr.append(ptr[-1] + t.size)
    q_terms = np.concatenate(terms).astype(np.int32) if terms else np.zeros(0, np.int32)
    return np.asarray(ptr, dtype=np.int64), q_terms, np.ones(q_terms.size, dtype=np.float32)


def doc_major_to_term_major(doc_ptr: np.ndarray, term_ids: np.ndarray, w: np.ndarray, n_terms: int):
    """Doc-major CSR -> term-major CSR (term_ptr i64, post_doc i32 ascending per term, post_w f32)."""
    n_docs = doc_ptr.shape[0] - 1
    doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), np.diff(doc_ptr))
    order = np.lexsort((doc_of, term_ids))
    term_ptr = np.zeros(n_terms + 1, dtype=np.int64)
    np.cumsum(np.bincount(term_ids, minlength=n_terms), out=term_ptr[1:])
    return term_ptr, doc_of[order].astype(np.int32), np.asarray(w)[order].astype(np.float32)
