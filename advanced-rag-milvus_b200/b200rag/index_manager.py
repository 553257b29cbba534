"""B200IndexManager -- the drop-in for the reference's Multi-Index Manager on the retrieval path.

The reference's HybridRetriever only ever touches its index manager through a duck type (reference
src/advanced_rag/retrieval.py:113-131, 341-419, 634-648):

    await manager.search(query_embedding, collection_name, top_k, filters, search_params) -> List[dict]
    await manager._generate_semantic_embedding(text) / _generate_sparse_embedding(text) / _generate_domain_embedding(text, domain)
    manager.collections            (a dict; sparse search is skipped unless it contains "sparse_index")

and MilvusIndexManager (reference src/advanced_rag/indexing.py:80-713) implements it with three Milvus collections.
This class implements the same surface -- same names, argument meaning, result dict shape (indexing.py:534-551) and
error behaviour (ValueError for an unknown collection, :466-467; ValueError for a malformed sparse query, :497-498) --
over device-resident indexes searched by the CUDA kernels in libb200rag.so:

    "semantic_index" / "domain_index"   exact cosine flat scan on tcgen05 tensor cores    (engine.DenseIndex)
    "sparse_index"                      sparse inner product over blocked postings        (engine.SparseIndex)

Differences that a caller can observe, all deliberate:
  * search is EXACT (the reference's HNSW ef=64 is approximate); ties rank by insertion row ascending.
  * `search_batch` runs a whole query batch in one kernel launch; `search` is the batch-of-one special case.
  * payload columns (content, doc_id, ...) stay on the host and are gathered only for the returned rows.
There is no CPU fallback: constructing the manager without a CUDA device or without libb200rag.so raises.
"""
from __future__ import annotations

import asyncio
import os
import re
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import bm25 as _bm25
from . import engine

_DENSE = ("semantic_index", "domain_index")
_PAYLOAD_FIELDS = ("doc_id", "chunk_index", "entropy", "redundancy", "domain_density", "timestamp")
_FILTER_FIELDS = ("doc_id", "chunk_id", "domain_density", "timestamp", "entropy", "redundancy", "chunk_index", "token_count")


class _Collection:
    """What `manager.collections[name]` holds: enough of pymilvus.Collection's read-only surface for the reference's
    get_collection_stats (indexing.py:678-690)."""

    def __init__(self, name: str, kind: str, manager: "B200IndexManager"):
        self.name, self.kind, self._m = name, kind, manager

    @property
    def num_entities(self) -> int:
        return self._m.num_rows

    @property
    def schema(self) -> str:
        dim = {"semantic_index": self._m.semantic_dim, "sparse_index": self._m.sparse_dim,
               "domain_index": self._m.domain_dim}[self.name]
        return f"{self.name}(id VARCHAR PK, embedding {self.kind}[{dim}], payload on host)"

    @property
    def indexes(self) -> List[str]:
        return ["FLAT_EXACT/COSINE (tcgen05 scan)" if self.kind == "dense" else "BLOCKED_POSTINGS/IP"]

    def release(self) -> None:
        pass


class PayloadStore:
    """Host-side columns of the collection schema (reference indexing.py:191-225): one entry per row."""

    def __init__(self):
        self.ids: List[str] = []
        self.content: List[str] = []
        self.cols: Dict[str, list] = {f: [] for f in _PAYLOAD_FIELDS + ("token_count",)}
        self.row_of: Dict[str, int] = {}

    def __len__(self) -> int:
        return len(self.ids)

    def append(self, chunk_id: str, content: str, meta: Dict[str, Any]) -> None:
        self.row_of.setdefault(chunk_id, len(self.ids))
        self.ids.append(chunk_id)
        self.content.append(content)
        for f in self.cols:
            self.cols[f].append(meta.get(f))

    def hit(self, row: int, score: float) -> Dict[str, Any]:
        """A FRESH result dict per call -- downstream code mutates hits in place (retrieval.py:361-363,469-470)."""
        return {"id": self.ids[row], "content": self.content[row], "score": score,
                "metadata": {f: self.cols[f][row] for f in _PAYLOAD_FIELDS}}

    def column(self, name: str) -> np.ndarray:
        if name == "chunk_id":
            return np.asarray(self.ids, dtype=object)
        return np.asarray(self.cols[name], dtype=object)


# ------------------------------------------------------------------------------------------------ filter expressions
_TERM = re.compile(r'\s*([A-Za-z_][A-Za-z0-9_]*)\s*(==|!=|>=|<=|>|<)\s*("(?:[^"\\]|\\.)*"|[^\s]+)\s*')


def _parse_filter(expr: str) -> List[Tuple[str, str, Any]]:
    """Parse the boolean strings HybridRetriever._build_filter_expression emits (retrieval.py:573-632):
    `field op value` terms joined by ' and '; string values are double-quoted with backslash escapes."""
    terms, pos = [], 0
    while pos < len(expr):
        m = _TERM.match(expr, pos)
        if not m:
            raise ValueError(f"cannot parse filter expression at {expr[pos:pos + 40]!r}")
        field, op, raw = m.groups()
        if field not in _FILTER_FIELDS:
            raise ValueError(f"Invalid filter field: {field}")
        if raw.startswith('"'):
            val: Any = re.sub(r"\\(.)", r"\1", raw[1:-1])
        elif raw in ("True", "False"):
            val = raw == "True"
        else:
            val = float(raw) if any(c in raw for c in ".eE") or raw in ("inf", "nan") else int(raw)
        terms.append((field, op, val))
        pos = m.end()
        if pos < len(expr):
            if not expr.startswith("and", pos):
                raise ValueError(f"expected 'and' in filter expression at {expr[pos:pos + 20]!r}")
            pos += 3
    return terms


def _eval_filter(store: PayloadStore, expr: str) -> np.ndarray:
    """Row mask of a filter expression over the payload columns (rows with a missing value never match)."""
    mask = np.ones(len(store), dtype=bool)
    ops = {"==": lambda a, b: a == b, "!=": lambda a, b: a != b, ">=": lambda a, b: a >= b,
           "<=": lambda a, b: a <= b, ">": lambda a, b: a > b, "<": lambda a, b: a < b}
    for field, op, val in _parse_filter(expr):
        col = store.column(field)
        ok = np.zeros(len(store), dtype=bool)
        for i, v in enumerate(col):                      # object columns: compare only like with like
            if v is None:
                continue
            if isinstance(val, str) != isinstance(v, str):
                continue
            ok[i] = bool(ops[op](v, val))
        mask &= ok
    return mask


class _MicroBatcher:
    """Turns concurrent one-query `search` awaits into GPU-sized batches.

    The reference serves one query per call with up to 64 requests in flight (service.py:149) and hops to a thread for
    every Milvus call (indexing.py:505).  Here the awaits that arrive within `max_wait_s` of each other (or until `max_batch`
    of them are queued) for the same (collection, top_k, filter) are answered by ONE `search_batch` call, which runs on a
    single worker thread so that the event loop stays responsive and GPU calls never overlap."""

    def __init__(self, manager: "B200IndexManager", max_batch: int, max_wait_s: float):
        import concurrent.futures as cf
        self.m, self.max_batch, self.max_wait_s = manager, int(max_batch), float(max_wait_s)
        self.pending: Dict[Tuple, List[Tuple[Any, "asyncio.Future"]]] = {}
        self.timers: Dict[Tuple, Any] = {}
        self.pool = cf.ThreadPoolExecutor(max_workers=1, thread_name_prefix="b200rag-batch")
        self.batches = 0                                   # number of search_batch calls issued (observability / tests)

    async def submit(self, query: Any, collection: str, top_k: int, filters: Optional[str]) -> List[Dict[str, Any]]:
        loop = asyncio.get_running_loop()
        key = (collection, int(top_k), filters)
        fut = loop.create_future()
        self.pending.setdefault(key, []).append((query, fut))
        if len(self.pending[key]) >= self.max_batch:
            self._flush(key)
        elif key not in self.timers:
            self.timers[key] = loop.call_later(self.max_wait_s, self._flush, key)
        return await fut

    def _flush(self, key: Tuple) -> None:
        timer = self.timers.pop(key, None)
        if timer is not None:
            timer.cancel()
        items = self.pending.pop(key, [])
        if not items:
            return
        collection, top_k, filters = key
        queries = [q for q, _ in items]
        batch: Any = queries if collection == "sparse_index" else np.stack([np.asarray(q, dtype=np.float32).reshape(-1) for q in queries])
        loop = asyncio.get_running_loop()
        self.batches += 1
        task = loop.run_in_executor(self.pool, self.m.search_batch, batch, collection, top_k, filters)

        def deliver(done):
            try:
                results = done.result()
            except Exception as e:  # noqa: BLE001 - every waiter of the batch sees the failure
                for _, f in items:
                    if not f.done():
                        f.set_exception(e)
                return
            for (_, f), hits in zip(items, results):
                if not f.done():
                    f.set_result(hits)

        task.add_done_callback(deliver)

    def close(self) -> None:
        self.pool.shutdown(wait=False)


class B200IndexManager:
    """Device-resident semantic / sparse / domain indexes behind the reference's index-manager duck type."""

    def __init__(self, semantic_dim: int = 1536, sparse_dim: int = 10000, domain_dim: int = 768,
                 device: str = "cuda", dtype: str = "f16", enable_sparse: Optional[bool] = None,
                 sparse_block_docs: int = 16384, host: str = "", port: int = 0, connect: bool = True,
                 micro_batch: bool = False, max_batch: int = 256, max_wait_ms: float = 0.5, **_ignored):
        # host / port / connect / enable_sharding / num_shards are accepted for signature compatibility with
        # MilvusIndexManager(...) (indexing.py:86-96); there is no server to connect to.
        self.semantic_dim, self.sparse_dim, self.domain_dim = int(semantic_dim), int(sparse_dim), int(domain_dim)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("B200IndexManager needs a CUDA device (b200rag has no CPU path)")
        engine._lib.load()                                   # fail loudly when the CUDA library is missing
        self.dtype = dtype
        self.embedding_generator = None                      # set externally, as in the reference (indexing.py:119)
        self._batcher = _MicroBatcher(self, max_batch, max_wait_ms * 1e-3) if micro_batch else None
        self.payload = PayloadStore()
        self._sem = engine.DenseIndex(self.semantic_dim, dtype, "COSINE", self.device)
        self._dom = engine.DenseIndex(self.domain_dim, dtype, "COSINE", self.device)
        self._dom_rows = 0
        if enable_sparse is None:
            enable_sparse = os.getenv("ENABLE_SPARSE", "1") == "1"      # reference indexing.py:156-158
        self._sparse_block_docs = sparse_block_docs
        self._sp_ptr: List[int] = [0]                        # doc-major CSR of sparse document vectors (host, ingest side)
        self._sp_idx: List[np.ndarray] = []
        self._sp_val: List[np.ndarray] = []
        self._sparse: Optional[engine.SparseIndex] = None    # rebuilt lazily after inserts
        self._sparse_dirty = False
        self._tok_vocab: Dict[str, int] = {}                 # MMR token sets: content.lower().split() (retrieval.py:497)
        self._tok_ptr: List[int] = [0]
        self._tok_ids: List[np.ndarray] = []
        self._tok_dev: Optional[Tuple[torch.Tensor, torch.Tensor, int]] = None
        self._mask_cache: Dict[str, Tuple[int, int, Optional[torch.Tensor]]] = {}     # filter expression -> (rows, allowed, bit mask)
        self.collections: Dict[str, _Collection] = {"semantic_index": _Collection("semantic_index", "dense", self),
                                                    "domain_index": _Collection("domain_index", "dense", self)}
        if enable_sparse:
            self.collections["sparse_index"] = _Collection("sparse_index", "sparse", self)

    # ------------------------------------------------------------------------------------------- ingest
    @property
    def num_rows(self) -> int:
        return len(self.payload)

    def add(self, ids: Sequence[str], contents: Sequence[str], semantic: Any, sparse: Optional[Sequence[Dict]] = None,
            domain: Any = None, metadata: Optional[Sequence[Dict[str, Any]]] = None) -> None:
        """Append rows to all collections at once (row i of every index is the same chunk, as in the reference where
        the three collections hold the same chunk ids, indexing.py:346-347).

        semantic / domain: fp32 [n, dim] (numpy or torch, host or device); sparse: one {"indices","values"} dict per
        row (the reference's SPARSE_FLOAT_VECTOR payload, indexing.py:645-654) or None for an empty sparse index."""
        n = len(ids)
        sem = torch.as_tensor(np.asarray(semantic, dtype=np.float32) if not torch.is_tensor(semantic) else semantic)
        if sem.shape != (n, self.semantic_dim):
            raise ValueError(f"semantic embeddings must be [{n}, {self.semantic_dim}], got {tuple(sem.shape)}")
        self._sem.add(sem)
        if domain is not None:
            dom = torch.as_tensor(np.asarray(domain, dtype=np.float32) if not torch.is_tensor(domain) else domain)
            if dom.shape != (n, self.domain_dim):
                raise ValueError(f"domain embeddings must be [{n}, {self.domain_dim}], got {tuple(dom.shape)}")
            if self._dom_rows != self.num_rows:
                raise ValueError("domain embeddings must be supplied for every row or for none")
            self._dom.add(dom)
            self._dom_rows += n
        for r in range(n):
            entry = sparse[r] if sparse is not None else None
            idx = np.asarray(entry["indices"], dtype=np.int64) if entry else np.zeros(0, np.int64)
            val = np.asarray(entry["values"], dtype=np.float32) if entry else np.zeros(0, np.float32)
            if idx.size and (idx.min() < 0 or idx.max() >= self.sparse_dim):
                raise ValueError(f"sparse index out of range [0, {self.sparse_dim})")
            order = np.argsort(idx, kind="stable")
            self._sp_idx.append(idx[order])
            self._sp_val.append(val[order])
            self._sp_ptr.append(self._sp_ptr[-1] + idx.size)
            toks = sorted({self._tok_vocab.setdefault(t, len(self._tok_vocab)) for t in _bm25.tokenize(contents[r])})
            self._tok_ids.append(np.asarray(toks, dtype=np.int32))
            self._tok_ptr.append(self._tok_ptr[-1] + len(toks))
            self.payload.append(ids[r], contents[r], (metadata[r] if metadata is not None else {}) or {})
        self._sparse_dirty = True
        self._tok_dev = None
        self._mask_cache = {}

    async def index_chunks(self, chunks: List[Any], domain: Optional[str] = None) -> Dict[str, Any]:
        """Reference MilvusIndexManager.index_chunks (indexing.py:264-437): embed every chunk through the generator
        hooks, then insert.  `chunks` are the reference's Chunk objects (.text, .metadata.{chunk_id, doc_id, ...})."""
        summary = {"total_chunks": len(chunks), "indexed_semantic": 0, "indexed_sparse": 0, "indexed_domain": 0, "errors": []}
        ids, texts, sem, spa, dom, meta = [], [], [], [], [], []
        for ch in chunks:
            md = ch.metadata
            try:
                text = ch.text
                s = np.asarray(await self._generate_semantic_embedding(text), dtype=np.float32)
                d = np.asarray(await self._generate_domain_embedding(text, domain), dtype=np.float32)
                sp = await self._generate_sparse_embedding(text) if "sparse_index" in self.collections else None
            except Exception as e:  # noqa: BLE001 - per-chunk errors are collected, as in the reference (:360-364)
                summary["errors"].append({"chunk_id": getattr(md, "chunk_id", None), "error": str(e)})
                continue
            ids.append(md.chunk_id)
            texts.append(text[:65535])
            sem.append(s)
            dom.append(d)
            spa.append(sp)
            meta.append({"doc_id": md.doc_id, "chunk_index": md.chunk_index, "token_count": getattr(md, "token_count", None),
                         "entropy": getattr(md, "entropy", None), "redundancy": getattr(md, "redundancy", None),
                         "domain_density": getattr(md, "domain_density", None), "timestamp": getattr(md, "timestamp", None)})
        if ids:
            self.add(ids, texts, np.stack(sem), spa if "sparse_index" in self.collections else None, np.stack(dom), meta)
            summary["indexed_semantic"] = summary["indexed_domain"] = len(ids)
            if "sparse_index" in self.collections:
                summary["indexed_sparse"] = len(ids)
        return summary

    # ------------------------------------------------------------------------------------------- device views
    def _sparse_index(self) -> engine.SparseIndex:
        if self._sparse is None or self._sparse_dirty:
            idx = np.concatenate(self._sp_idx) if self._sp_idx else np.zeros(0, np.int64)
            val = np.concatenate(self._sp_val) if self._sp_val else np.zeros(0, np.float32)
            self._sparse = engine.SparseIndex(np.asarray(self._sp_ptr, dtype=np.int64), idx, val, self.sparse_dim,
                                              self.device, block_docs=self._sparse_block_docs)
            self._sparse_dirty = False
        return self._sparse

    def token_sets(self) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """Device CSR of every row's sorted unique token ids (input of the MMR kernel)."""
        if self._tok_dev is None:
            ptr = torch.as_tensor(np.asarray(self._tok_ptr, dtype=np.int64)).to(self.device)
            ids = torch.as_tensor(np.concatenate(self._tok_ids) if self._tok_ids else np.zeros(0, np.int32)).to(self.device)
            self._tok_dev = (ptr, ids, max(1, len(self._tok_vocab)))
        return self._tok_dev

    # ------------------------------------------------------------------------------------------- search
    def _dense_of(self, name: str) -> engine.DenseIndex:
        return self._sem if name == "semantic_index" else self._dom

    def _sparse_queries(self, queries: Sequence[Any]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        ptr, terms, vals = [0], [], []
        for q in queries:
            if isinstance(q, dict):
                i = np.asarray(q.get("indices", []), dtype=np.int64)
                v = np.asarray(q.get("values", []), dtype=np.float32)
            elif hasattr(q, "tocsr"):
                c = q.tocsr()
                i, v = c.indices.astype(np.int64), c.data.astype(np.float32)
            else:
                raise ValueError("Sparse query embedding must be dict with indices/values or a scipy.sparse matrix")
            keep = (i >= 0) & (i < self.sparse_dim)
            i, v = i[keep], v[keep]
            order = np.argsort(i, kind="stable")
            terms.append(i[order].astype(np.int32))
            vals.append(v[order])
            ptr.append(ptr[-1] + i.size)
        return (np.asarray(ptr, dtype=np.int64), np.concatenate(terms) if terms else np.zeros(0, np.int32),
                np.concatenate(vals) if vals else np.zeros(0, np.float32))

    def search_batch_ids(self, queries: Any, collection_name: str, top_k: int = 20, filters: Optional[str] = None
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Batched search that stays on the device: (scores f64 [B,k], row ids i64 [B,k] (-1 = empty), counts i32 [B])."""
        if collection_name not in self.collections:
            raise ValueError(f"Collection {collection_name} not found")
        n = self.num_rows
        k = int(top_k)
        if k <= 0:
            raise ValueError("top_k must be positive")
        if filters:
            return self._search_filtered(queries, collection_name, k, filters)
        if collection_name in _DENSE:
            idx = self._dense_of(collection_name)
            q = torch.as_tensor(np.asarray(queries, dtype=np.float32) if not torch.is_tensor(queries) else queries)
            if q.dim() == 1:
                q = q[None, :]
            b = q.shape[0]
            if idx.n == 0:
                return (torch.full((b, k), float("-inf"), dtype=torch.float64, device=self.device),
                        torch.full((b, k), -1, dtype=torch.int64, device=self.device),
                        torch.zeros(b, dtype=torch.int32, device=self.device))
            s, i, _ = idx.search(q, k)
            cnt = torch.full((b,), min(k, idx.n), dtype=torch.int32, device=self.device)
            return s, i, cnt
        qp, qt, qv = self._sparse_queries(list(queries))
        b = qp.shape[0] - 1
        if n == 0 or self._sp_ptr[-1] == 0:
            return (torch.full((b, k), float("-inf"), dtype=torch.float64, device=self.device),
                    torch.full((b, k), -1, dtype=torch.int64, device=self.device),
                    torch.zeros(b, dtype=torch.int32, device=self.device))
        s, i, c = self._sparse_index().search(qp, qt, qv, k)
        return s.to(torch.float64), i, c

    def _search_filtered(self, queries, collection_name: str, k: int, expr: str):
        """Exact filtered search: the predicate is evaluated to a row bit mask on the host columns (cached per expression
        until the next insert / delete) and applied INSIDE the scan kernels -- sample pass, epilogue survivors and exact
        fallback for the dense indexes, candidate collection for the sparse one (b200rag_*_topk_masked)."""
        cached = self._mask_cache.get(expr)
        if cached is None or cached[0] != self.num_rows:
            allowed = _eval_filter(self.payload, expr)
            words = engine.pack_row_mask(torch.as_tensor(allowed).to(self.device)) if allowed.size else None
            cached = (self.num_rows, int(allowed.sum()), words)
            self._mask_cache = {expr: cached}                 # one entry: serving loops repeat the same filter
        _, m, words = cached
        b = (len(queries) if collection_name == "sparse_index" else (1 if np.ndim(queries) == 1 else np.shape(queries)[0]))
        if m == 0 or words is None:
            return (torch.full((b, k), float("-inf"), dtype=torch.float64, device=self.device),
                    torch.full((b, k), -1, dtype=torch.int64, device=self.device),
                    torch.zeros(b, dtype=torch.int32, device=self.device))
        if collection_name in _DENSE:
            idx = self._dense_of(collection_name)
            q = torch.as_tensor(np.asarray(queries, dtype=np.float32) if not torch.is_tensor(queries) else queries)
            s, i, _ = idx.search(q if q.dim() == 2 else q[None, :], k, row_mask=words)
            return s, i, torch.full((b,), min(k, m), dtype=torch.int32, device=self.device)
        qp, qt, qv = self._sparse_queries(list(queries))
        s, i, c = self._sparse_index().search(qp, qt, qv, k, doc_mask=words)
        return s.to(torch.float64), i, c

    def search_batch(self, queries: Any, collection_name: str, top_k: int = 20, filters: Optional[str] = None,
                     search_params: Optional[Dict] = None) -> List[List[Dict[str, Any]]]:
        """One hit list per query, each in the reference's result format (indexing.py:534-551), best first."""
        s, i, c = self.search_batch_ids(queries, collection_name, top_k, filters)
        s_h, i_h, c_h = s.cpu().numpy(), i.cpu().numpy(), c.cpu().numpy()
        return [[self.payload.hit(int(i_h[b, r]), float(s_h[b, r])) for r in range(int(c_h[b])) if i_h[b, r] >= 0]
                for b in range(i_h.shape[0])]

    async def search(self, query_embedding: Any, collection_name: str, top_k: int = 20, filters: Optional[str] = None,
                     search_params: Optional[Dict] = None) -> List[Dict[str, Any]]:
        """Reference MilvusIndexManager.search (indexing.py:445-551) for one query.  `search_params` is accepted and
        ignored: the scan is exact, there is no ef / drop_ratio to tune."""
        if collection_name not in self.collections:
            raise ValueError(f"Collection {collection_name} not found")
        if collection_name == "sparse_index":
            if not (isinstance(query_embedding, dict) or hasattr(query_embedding, "tocsr")):
                raise ValueError("Sparse query embedding must be dict with indices/values or a scipy.sparse matrix")
            if self._batcher is not None:
                return await self._batcher.submit(query_embedding, collection_name, top_k, filters)
            batch: Any = [query_embedding]
        else:
            if self._batcher is not None:
                return await self._batcher.submit(query_embedding, collection_name, top_k, filters)
            batch = np.asarray(query_embedding, dtype=np.float32).reshape(1, -1)
        return self.search_batch(batch, collection_name, top_k, filters, search_params)[0]

    # ------------------------------------------------------------------------------------------- embedding hooks
    async def _call_generator(self, name: str, *args):
        fn = getattr(self.embedding_generator, name)
        if asyncio.iscoroutinefunction(fn):
            return await fn(*args)
        return await asyncio.get_event_loop().run_in_executor(None, lambda: fn(*args))

    async def _generate_semantic_embedding(self, text: str) -> np.ndarray:
        """indexing.py:601-627: the user generator, else a random placeholder vector."""
        if self.embedding_generator:
            return await self._call_generator("encode_semantic", text)
        return np.random.randn(self.semantic_dim).astype(np.float32)

    async def _generate_sparse_embedding(self, text: str):
        """indexing.py:629-654: the user generator, else a random 100-nnz placeholder with sorted indices."""
        if self.embedding_generator:
            return await self._call_generator("encode_sparse", text)
        nnz = min(100, self.sparse_dim)
        idx = np.sort(np.random.choice(self.sparse_dim, size=nnz, replace=False))
        return {"indices": idx.tolist(), "values": np.abs(np.random.randn(nnz)).astype(float).tolist()}

    async def _generate_domain_embedding(self, text: str, domain: Optional[str] = None) -> np.ndarray:
        """indexing.py:656-676."""
        if self.embedding_generator:
            return await self._call_generator("encode_domain", text, domain or "")
        return np.random.randn(self.domain_dim).astype(np.float32)

    # ------------------------------------------------------------------------------------------- housekeeping
    def get_collection_stats(self, collection_name: str) -> Dict[str, Any]:
        if collection_name not in self.collections:
            return {}
        c = self.collections[collection_name]
        return {"name": collection_name, "num_entities": c.num_entities, "schema": c.schema, "indexes": list(c.indexes)}

    def _keep_rows(self, keep: np.ndarray) -> None:
        """Compact every index and the payload to the rows `keep` (ascending row numbers) marks."""
        rows = np.flatnonzero(keep)
        dev_rows = torch.as_tensor(rows).to(self.device)
        for name in ("_sem", "_dom"):
            old = getattr(self, name)
            if old.n == 0:
                continue
            if old.n != keep.size:
                raise ValueError("domain index does not cover every row; cannot delete consistently")
            new = engine.DenseIndex(old.dim, self.dtype, "COSINE", self.device, capacity=rows.size)
            if rows.size:
                new.add_prepared(old.rows[dev_rows])
            setattr(self, name, new)
        self._dom_rows = self._dom.n
        self._sp_idx = [self._sp_idx[r] for r in rows]
        self._sp_val = [self._sp_val[r] for r in rows]
        self._sp_ptr = [0]
        for a in self._sp_idx:
            self._sp_ptr.append(self._sp_ptr[-1] + a.size)
        self._tok_ids = [self._tok_ids[r] for r in rows]
        self._tok_ptr = [0]
        for a in self._tok_ids:
            self._tok_ptr.append(self._tok_ptr[-1] + a.size)
        old_p = self.payload
        self.payload = PayloadStore()
        for r in rows:
            self.payload.append(old_p.ids[r], old_p.content[r], {f: old_p.cols[f][r] for f in old_p.cols})
        self._sparse, self._sparse_dirty, self._tok_dev = None, True, None
        self._mask_cache = {}

    async def delete_by_filter(self, collection_name: str, expr: str):
        """Reference MilvusIndexManager.delete_by_filter (indexing.py:692-695: `collection.delete(expr)`).  The three
        collections hold the same rows here, so the matching rows leave ALL of them (the reference deletes per collection;
        a chunk missing from one index but not the others is not a state this engine represents).  Returns the count."""
        if collection_name not in self.collections:
            raise ValueError(f"Collection {collection_name} not found")
        mask = _eval_filter(self.payload, expr)
        n_del = int(mask.sum())
        if n_del:
            self._keep_rows(~mask)
        return n_del

    # ------------------------------------------------------------------------------------------- checkpoint / resume
    def save(self, path: str) -> None:
        """Write the whole index (stored 16-bit rows, sparse CSR, token sets, payload) to one torch file.  The reference
        leaves durability to the Milvus server (collection.flush, indexing.py:430-431)."""
        state = {
            "version": 1, "dtype": self.dtype,
            "dims": (self.semantic_dim, self.sparse_dim, self.domain_dim),
            "sem": self._sem.rows.cpu(), "dom": self._dom.rows.cpu(),
            "sp_ptr": np.asarray(self._sp_ptr, dtype=np.int64),
            "sp_idx": np.concatenate(self._sp_idx) if self._sp_idx else np.zeros(0, np.int64),
            "sp_val": np.concatenate(self._sp_val) if self._sp_val else np.zeros(0, np.float32),
            "tok_vocab": self._tok_vocab, "tok_ptr": np.asarray(self._tok_ptr, dtype=np.int64),
            "tok_ids": np.concatenate(self._tok_ids) if self._tok_ids else np.zeros(0, np.int32),
            "ids": self.payload.ids, "content": self.payload.content, "cols": self.payload.cols,
            "sparse_enabled": "sparse_index" in self.collections,
        }
        torch.save(state, path)

    @classmethod
    def load(cls, path: str, device: str = "cuda", **kwargs) -> "B200IndexManager":
        st = torch.load(path, map_location="cpu", weights_only=False)
        if st.get("version") != 1:
            raise ValueError("unknown index file version")
        sd, pd, dd = st["dims"]
        m = cls(semantic_dim=sd, sparse_dim=pd, domain_dim=dd, device=device, dtype=st["dtype"],
                enable_sparse=st["sparse_enabled"], **kwargs)
        if st["sem"].shape[0]:
            m._sem.add_prepared(st["sem"])
        if st["dom"].shape[0]:
            m._dom.add_prepared(st["dom"])
        m._dom_rows = m._dom.n
        sp_ptr, tok_ptr = st["sp_ptr"], st["tok_ptr"]
        n = len(st["ids"])
        m._sp_ptr = [int(v) for v in sp_ptr]
        m._sp_idx = [st["sp_idx"][sp_ptr[r]: sp_ptr[r + 1]] for r in range(n)]
        m._sp_val = [st["sp_val"][sp_ptr[r]: sp_ptr[r + 1]] for r in range(n)]
        m._tok_vocab = dict(st["tok_vocab"])
        m._tok_ptr = [int(v) for v in tok_ptr]
        m._tok_ids = [st["tok_ids"][tok_ptr[r]: tok_ptr[r + 1]] for r in range(n)]
        for r in range(n):
            m.payload.append(st["ids"][r], st["content"][r], {f: st["cols"][f][r] for f in st["cols"]})
        m._sparse_dirty = True
        return m

    async def close(self):
        if self._batcher is not None:
            self._batcher.close()
        self._sparse = None
        self._tok_dev = None
