"""In-memory index manager ("Milvus mocked") -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Implements the duck type the reference's HybridRetriever calls (reference src/advanced_rag/retrieval.py:341-419,
634-648; result dict shape src/advanced_rag/indexing.py:534-551) on top of the CPU oracle (oracle/exact_scan.c):
exact cosine flat search for "semantic_index" / "domain_index", sparse inner product for "sparse_index", ties by row
ascending.  Driving the UNMODIFIED reference HybridRetriever with this manager is "the reference's own retrieval path
on identical synthetic embeddings and postings" that BASELINE.json's north_star names; oracle/gen_e2e_golden.py does
exactly that to produce tests/golden/e2e_golden.json.  Parity: "unpinned" for the dense / sparse scores themselves
(they live in a Milvus server the reference does not ship), pinned for everything downstream of them.
"""
from __future__ import annotations

import zlib
from typing import Any, Dict, List, Optional

import numpy as np

from . import oracle


def text_seed(text: str, salt: int) -> int:
    return (zlib.crc32(text.encode("utf-8")) * 2654435761 + salt) % (1 << 32)


class HashEmbeddingGenerator:
    """Deterministic stand-in for the user's embedding model (the reference's embedding_generator hook,
    indexing.py:119,601-676): dense vectors are standard normal draws seeded by crc32(text); the sparse query vector
    holds value 1.0 for every known token of text.lower().split()."""

    def __init__(self, semantic_dim: int, domain_dim: int, vocab: Dict[str, int]):
        self.semantic_dim, self.domain_dim, self.vocab = semantic_dim, domain_dim, vocab

    def encode_semantic(self, text: str) -> np.ndarray:
        return np.random.default_rng(text_seed(text, 1)).standard_normal(self.semantic_dim).astype(np.float32)

    def encode_domain(self, text: str, domain: str = "") -> np.ndarray:
        return np.random.default_rng(text_seed(text + "|" + (domain or ""), 2)).standard_normal(self.domain_dim).astype(np.float32)

    def encode_sparse(self, text: str) -> Dict[str, list]:
        ids = sorted({self.vocab[t] for t in (text or "").lower().split() if t in self.vocab})
        return {"indices": ids, "values": [1.0] * len(ids)}


class InMemoryIndexManager:
    def __init__(self, ids: List[str], contents: List[str], metadata: List[Dict[str, Any]], semantic: np.ndarray,
                 domain: Optional[np.ndarray], sp_ptr: np.ndarray, sp_idx: np.ndarray, sp_val: np.ndarray,
                 sparse_dim: int, generator, dtype: int = oracle.F16, with_sparse: bool = True):
        self.ids, self.contents, self.metadata = ids, contents, metadata
        self.dtype = dtype
        self.sem_bits = oracle.normalize_rows(semantic, dtype)
        self.dom_bits = oracle.normalize_rows(domain, dtype) if domain is not None else None
        self.n = len(ids)
        self.sparse_dim = sparse_dim
        # term-major postings for the oracle's sparse search
        doc_of = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(sp_ptr))
        order = np.lexsort((doc_of, sp_idx))
        self.term_ptr = np.zeros(sparse_dim + 1, dtype=np.int64)
        np.cumsum(np.bincount(sp_idx, minlength=sparse_dim), out=self.term_ptr[1:])
        self.post_doc = doc_of[order].astype(np.int32)
        self.post_w = np.asarray(sp_val, dtype=np.float32)[order]
        self.embedding_generator = generator
        self.collections = {"semantic_index": object(), "domain_index": object()}
        if with_sparse:
            self.collections["sparse_index"] = object()

    def _hit(self, row: int, score: float) -> Dict[str, Any]:
        md = self.metadata[row]
        return {"id": self.ids[row], "content": self.contents[row], "score": score,
                "metadata": {f: md.get(f) for f in ("doc_id", "chunk_index", "entropy", "redundancy", "domain_density", "timestamp")}}

    async def search(self, query_embedding, collection_name: str, top_k: int = 20, filters: Optional[str] = None,
                     search_params: Optional[Dict] = None) -> List[Dict[str, Any]]:
        if collection_name not in self.collections:
            raise ValueError(f"Collection {collection_name} not found")
        assert filters is None, "the mocked index does not evaluate filter expressions"
        if collection_name == "sparse_index":
            qi = np.asarray(query_embedding["indices"], dtype=np.int32)
            qv = np.asarray(query_embedding["values"], dtype=np.float32)
            order = np.argsort(qi, kind="stable")
            s, i, c = oracle.sparse_topk(self.term_ptr, self.post_doc, self.post_w, self.n,
                                         np.asarray([0, qi.size], dtype=np.int64), qi[order], qv[order], top_k)
            return [self._hit(int(i[0, r]), float(s[0, r])) for r in range(int(c[0]))]
        bits = self.sem_bits if collection_name == "semantic_index" else self.dom_bits
        q = oracle.normalize_rows(np.asarray(query_embedding, dtype=np.float32).reshape(1, -1), self.dtype)
        s, i = oracle.dense_topk(bits, q, top_k, self.dtype)
        return [self._hit(int(i[0, r]), float(s[0, r])) for r in range(min(top_k, self.n)) if i[0, r] >= 0]

    async def _generate_semantic_embedding(self, text: str):
        return self.embedding_generator.encode_semantic(text)

    async def _generate_sparse_embedding(self, text: str):
        return self.embedding_generator.encode_sparse(text)

    async def _generate_domain_embedding(self, text: str, domain: Optional[str] = None):
        return self.embedding_generator.encode_domain(text, domain or "")
