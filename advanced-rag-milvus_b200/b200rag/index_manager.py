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

The boundary has three result forms, all produced by the same kernels:

    search_batch_ids     device tensors (scores f64, rows i64, counts i32) -- what retrieve_batch consumes
    search_batch_arrays  SearchArrays: the same as numpy arrays on the host (one D2H copy each through pinned memory) plus a
                         lazy `hits()` view; this is the columnar plugin call bench.py's `e2e` times
    search_batch/search  the reference's List[dict] per query (indexing.py:534-551), fresh dicts on every call, built in bulk
                         from typed payload columns (one gather per column, not one Python call per hit)

Differences that a caller can observe, all deliberate:
  * search is EXACT (the reference's HNSW ef=64 is approximate); ties rank by insertion row ascending.
  * `search_batch*` run a whole query batch in one kernel launch; `search` is the batch-of-one special case.
  * payload columns are typed as the collection schema types them (indexing.py:191-225): INT64 chunk_index / token_count,
    FLOAT entropy / redundancy / domain_density, VARCHAR doc_id / timestamp (dictionary encoded); missing values are None.
  * metadata predicates are evaluated ON THE GPU over device copies of those columns (b200rag_filter_mask); deletes are
    tombstones in the same row mask until `compact()`.
There is no CPU fallback: constructing the manager without a CUDA device or without libb200rag.so raises.
"""
from __future__ import annotations

import asyncio
import json
import operator
import os
import re
import threading
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import bm25 as _bm25
from . import engine

_DENSE = ("semantic_index", "domain_index")
_PAYLOAD_FIELDS = ("doc_id", "chunk_index", "entropy", "redundancy", "domain_density", "timestamp")
_FILTER_FIELDS = ("doc_id", "chunk_id", "domain_density", "timestamp", "entropy", "redundancy", "chunk_index", "token_count")
_FLOAT_COLS = ("entropy", "redundancy", "domain_density")       # FLOAT in the collection schema (indexing.py:201-203)
_INT_COLS = ("chunk_index", "token_count")                      # INT64 (indexing.py:196-197)
_STR_COLS = ("doc_id", "timestamp")                             # VARCHAR (indexing.py:194,222)
_INT_MISSING = np.iinfo(np.int64).min
_OPS = {"==": _lib.OP_EQ, "!=": _lib.OP_NE, ">=": _lib.OP_GE, "<=": _lib.OP_LE, ">": _lib.OP_GT, "<": _lib.OP_LT}
_PY_OPS = {"==": operator.eq, "!=": operator.ne, ">=": operator.ge, "<=": operator.le, ">": operator.gt, "<": operator.lt}


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
        return f"{self.name}(id VARCHAR PK, embedding {self.kind}[{dim}], payload columns on host + device)"

    @property
    def indexes(self) -> List[str]:
        return ["FLAT_EXACT/COSINE (tcgen05 scan)" if self.kind == "dense" else "BLOCKED_POSTINGS/IP"]

    def release(self) -> None:
        pass


class _Grow:
    """Append-only numpy column with amortised growth (ingest appends batches; searches read the valid prefix)."""

    def __init__(self, dtype):
        self._a = np.empty(0, dtype=dtype)
        self.n = 0

    def extend(self, values: np.ndarray) -> None:
        m = self.n + values.shape[0]
        if m > self._a.shape[0]:
            new = np.empty(max(m, int(self._a.shape[0] * 1.5), 1024), dtype=self._a.dtype)
            new[: self.n] = self._a[: self.n]
            self._a = new
        self._a[self.n: m] = values
        self.n = m

    @property
    def view(self) -> np.ndarray:
        return self._a[: self.n]

    def replace(self, values: np.ndarray) -> None:
        self._a = np.ascontiguousarray(values)
        self.n = values.shape[0]


class PayloadStore:
    """Columns of the collection schema (reference indexing.py:191-225), one entry per row: chunk ids and contents as
    Python strings (they are handed out again verbatim), the scalar fields as typed numpy columns -- float64 with NaN for a
    missing value, int64 with INT64_MIN, dictionary codes (int32, -1 = missing) for the VARCHAR fields -- which is also the
    form the GPU predicate kernel reads them in."""

    def __init__(self):
        self.ids: List[str] = []
        self.content: List[str] = []
        self.num: Dict[str, _Grow] = {f: _Grow(np.float64) for f in _FLOAT_COLS}
        self.num.update({f: _Grow(np.int64) for f in _INT_COLS})
        self.codes: Dict[str, _Grow] = {f: _Grow(np.int32) for f in _STR_COLS}
        self.dict_values: Dict[str, List[str]] = {f: [] for f in _STR_COLS}
        self.dict_code: Dict[str, Dict[str, int]] = {f: {} for f in _STR_COLS}
        self._row_of: Optional[Dict[str, int]] = None          # chunk_id -> first row, built on demand (chunk_id filters)
        self.virtual_rows = 0                                  # > 0: payload-less bulk rows (benchmark-scale loads), see set_virtual

    def __len__(self) -> int:
        return self.virtual_rows or len(self.ids)

    def set_virtual(self, n_rows: int) -> None:
        """Rows without stored payload: id = f"c{row:09d}", empty content, every scalar field missing.  For corpora whose
        payload lives elsewhere (bench.py's 10M / 100M-row synthetic shards); cannot be mixed with real rows."""
        if self.ids:
            raise ValueError("payload-less rows cannot be mixed with stored payload")
        self.virtual_rows = int(n_rows)

    # ---- ingest: convert first (may raise), commit afterwards ---------------------------------------------------
    def convert(self, n: int, metadata: Optional[Sequence[Optional[Dict[str, Any]]]]) -> Dict[str, Any]:
        """Typed column batches for n new rows; raises ValueError on a value the schema type cannot hold.  Nothing is
        mutated except the dictionaries' pending entries, which `commit` applies."""
        if metadata is None:                                   # bulk loads without scalar fields: every value is missing
            out = {f: np.full(n, np.nan) for f in _FLOAT_COLS}
            out.update({f: np.full(n, _INT_MISSING, dtype=np.int64) for f in _INT_COLS})
            for f in _STR_COLS:
                out[f] = np.full(n, -1, dtype=np.int32)
                out["_new_" + f] = {}
            return out
        meta = [(m or {}) for m in metadata]
        if len(meta) != n:
            raise ValueError(f"metadata must hold one entry per row ({len(meta)} != {n})")
        out: Dict[str, Any] = {}
        for f in _FLOAT_COLS:
            try:
                out[f] = np.asarray([np.nan if m.get(f) is None else float(m[f]) for m in meta], dtype=np.float64)
            except (TypeError, ValueError) as e:
                raise ValueError(f"metadata field {f!r} must be a number: {e}") from None
        for f in _INT_COLS:
            try:
                out[f] = np.asarray([_INT_MISSING if m.get(f) is None else int(m[f]) for m in meta], dtype=np.int64)
            except (TypeError, ValueError, OverflowError) as e:
                raise ValueError(f"metadata field {f!r} must be an integer: {e}") from None
        for f in _STR_COLS:
            code, pending = self.dict_code[f], {}
            col = np.empty(n, dtype=np.int32)
            base = len(self.dict_values[f])
            for i, m in enumerate(meta):
                v = m.get(f)
                if v is None:
                    col[i] = -1
                    continue
                v = v if isinstance(v, str) else str(v)
                c = code.get(v)
                if c is None:
                    c = pending.get(v)
                    if c is None:
                        c = pending[v] = base + len(pending)
                col[i] = c
            out[f] = col
            out["_new_" + f] = pending
        return out

    def commit(self, ids: Sequence[str], contents: Sequence[str], cols: Dict[str, Any]) -> None:
        if self.virtual_rows:
            raise ValueError("payload-less rows cannot be mixed with stored payload")
        self.ids.extend(ids)
        self.content.extend(contents)
        for f in _FLOAT_COLS + _INT_COLS:
            self.num[f].extend(cols[f])
        for f in _STR_COLS:
            self.codes[f].extend(cols[f])
            for v, c in cols["_new_" + f].items():            # insertion order = code order
                assert c == len(self.dict_values[f])
                self.dict_values[f].append(v)
                self.dict_code[f][v] = c
        self._row_of = None

    def keep(self, rows: np.ndarray) -> None:
        """Compact to `rows` (ascending)."""
        get = rows.tolist()
        self.ids = [self.ids[r] for r in get]
        self.content = [self.content[r] for r in get]
        for g in list(self.num.values()) + list(self.codes.values()):
            g.replace(g.view[rows])
        self._row_of = None

    def row_of(self, chunk_id: str) -> int:
        if self._row_of is None:
            d: Dict[str, int] = {}
            for r, cid in enumerate(self.ids):
                d.setdefault(cid, r)
            self._row_of = d
        return self._row_of.get(chunk_id, -1)

    # ---- results ------------------------------------------------------------------------------------------------
    def _column_values(self, f: str, rows: np.ndarray) -> list:
        if f in _STR_COLS:
            c = self.codes[f].view[rows]
            vals = self.dict_values[f]
            return [vals[i] if i >= 0 else None for i in c.tolist()]
        a = self.num[f].view[rows]
        out = a.tolist()                                       # Python ints / floats in one C loop
        miss = np.isnan(a) if f in _FLOAT_COLS else a == _INT_MISSING
        if miss.any():
            for i in np.flatnonzero(miss).tolist():
                out[i] = None
        return out

    def hits(self, rows: np.ndarray, scores: np.ndarray) -> List[Dict[str, Any]]:
        """FRESH result dicts (reference indexing.py:534-551) for a flat list of rows -- downstream code mutates hits in
        place (retrieval.py:361-363,469-470).  One gather per column, then one dict display per hit."""
        n = int(rows.shape[0])
        if n == 0:
            return []
        get = rows.tolist()
        if self.virtual_rows:
            return [{"id": f"c{r:09d}", "content": "", "score": s, "metadata": dict.fromkeys(_PAYLOAD_FIELDS)}
                    for r, s in zip(get, scores.tolist())]
        if n == 1:
            ids, content = [self.ids[get[0]]], [self.content[get[0]]]
        else:
            pick = operator.itemgetter(*get)
            ids, content = pick(self.ids), pick(self.content)
        cols = [self._column_values(f, rows) for f in _PAYLOAD_FIELDS]
        return [{"id": i, "content": c, "score": s,
                 "metadata": {"doc_id": a, "chunk_index": b, "entropy": e, "redundancy": r, "domain_density": d, "timestamp": t}}
                for i, c, s, a, b, e, r, d, t in zip(ids, content, scores.tolist(), *cols)]

    def hit(self, row: int, score: float) -> Dict[str, Any]:
        return self.hits(np.asarray([row], dtype=np.int64), np.asarray([score], dtype=np.float64))[0]


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
            try:
                val = float(raw) if any(c in raw for c in ".eE") or raw in ("inf", "nan", "-inf") else int(raw)
            except ValueError:
                raise ValueError(f"cannot parse filter value {raw!r}") from None
        terms.append((field, op, val))
        pos = m.end()
        if pos < len(expr):
            if not expr.startswith("and", pos):
                raise ValueError(f"expected 'and' in filter expression at {expr[pos:pos + 20]!r}")
            pos += 3
    return terms


def eval_filter_host(store: PayloadStore, expr: str) -> np.ndarray:
    """Row mask of a filter expression over the HOST columns, vectorised (rows with a missing value never match; a literal
    whose type does not fit the column matches nothing).  The product evaluates predicates on the GPU
    (B200IndexManager._filter_words); this numpy form is what the CPU tests compare the device kernel's semantics against."""
    n = len(store)
    mask = np.ones(n, dtype=bool)
    for field, op, val in _parse_filter(expr):
        fn = _PY_OPS[op]
        if field == "chunk_id" or field in _STR_COLS:
            if not isinstance(val, str):
                ok = np.zeros(n, dtype=bool)
            elif field == "chunk_id":
                ok = np.fromiter((fn(c, val) for c in store.ids), dtype=bool, count=n)
            else:
                lut = np.fromiter((fn(v, val) for v in store.dict_values[field]), dtype=bool, count=len(store.dict_values[field]))
                codes = store.codes[field].view
                ok = (codes >= 0) & np.concatenate([lut, [False]])[codes]
        elif isinstance(val, str):
            ok = np.zeros(n, dtype=bool)
        elif field in _FLOAT_COLS:
            a = store.num[field].view
            with np.errstate(invalid="ignore"):
                ok = ~np.isnan(a) & fn(a, float(val))
        else:
            a = store.num[field].view
            ok = (a != _INT_MISSING) & fn(a.astype(np.float64) if isinstance(val, float) else a, val)
        mask &= ok
    return mask


# ------------------------------------------------------------------------------------------------ result containers
@dataclass
class SearchArrays:
    """Columnar result of a batched search on the host: row r of query b is valid iff r < counts[b]."""
    rows: np.ndarray        # int64 [B, k]  corpus rows, -1 padded
    scores: np.ndarray      # float64 [B, k]
    counts: np.ndarray      # int32 [B]
    store: PayloadStore

    def chunk_ids(self) -> List[List[str]]:
        if self.store.virtual_rows:
            return [[f"c{r:09d}" for r in self.rows[b, : int(self.counts[b])].tolist()] for b in range(self.rows.shape[0])]
        ids = self.store.ids
        return [[ids[r] for r in self.rows[b, : int(self.counts[b])].tolist()] for b in range(self.rows.shape[0])]

    def hits(self) -> "HitLists":
        return HitLists(self)


class HitLists(Sequence):
    """Lazy List[List[dict]] over a SearchArrays: the reference's hit dicts (indexing.py:534-551) are built when a query's
    list is first asked for -- fresh dicts on every access, like a fresh `search` call."""

    def __init__(self, arrays: SearchArrays):
        self.arrays = arrays

    def __len__(self) -> int:
        return int(self.arrays.rows.shape[0])

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[i] for i in range(*b.indices(len(self)))]
        a = self.arrays
        if b < 0:
            b += len(self)
        if not 0 <= b < len(self):
            raise IndexError(b)
        c = int(a.counts[b])
        return a.store.hits(a.rows[b, :c], a.scores[b, :c])

    def materialize(self) -> List[List[Dict[str, Any]]]:
        """All lists at once: one column gather for the whole batch."""
        a = self.arrays
        k = a.rows.shape[1]
        valid = (np.arange(k)[None, :] < a.counts[:, None]) & (a.rows >= 0)
        flat = a.store.hits(a.rows[valid], a.scores[valid])
        out, pos = [], 0
        for c in valid.sum(1).tolist():
            out.append(flat[pos: pos + c])
            pos += c
        return out


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


def _csr_take(ptr: np.ndarray, data: Sequence[np.ndarray], rows: np.ndarray) -> Tuple[np.ndarray, List[np.ndarray]]:
    """Rows `rows` of a CSR (ptr, data arrays sharing the nnz axis), vectorised."""
    lens = ptr[rows + 1] - ptr[rows]
    new_ptr = np.zeros(rows.shape[0] + 1, dtype=np.int64)
    np.cumsum(lens, out=new_ptr[1:])
    src = np.repeat(ptr[rows] - new_ptr[:-1], lens) + np.arange(int(new_ptr[-1]), dtype=np.int64)
    return new_ptr, [d[src] for d in data]


class B200IndexManager:
    """Device-resident semantic / sparse / domain indexes behind the reference's index-manager duck type."""

    MASK_CACHE_SIZE = 16
    _compactable = True                                      # (the row-sharded subclass keeps tombstones: row ranges are fixed)

    def __init__(self, semantic_dim: int = 1536, sparse_dim: int = 10000, domain_dim: int = 768,
                 device: str = "cuda", dtype: str = "f16", enable_sparse: Optional[bool] = None,
                 sparse_block_docs: int = 16384, host: str = "", port: int = 0, connect: bool = True,
                 micro_batch: bool = False, max_batch: int = 256, max_wait_ms: float = 0.5,
                 use_graphs: Optional[bool] = None, **_ignored):
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
        # one lock around every GPU search and every mutation: the micro-batcher's worker thread, the event loop thread and
        # ingest may all call in; the kernels of one call are ordered by the stream, calls by this lock
        self._lock = threading.RLock()
        self.payload = PayloadStore()
        self._sem = engine.DenseIndex(self.semantic_dim, dtype, "COSINE", self.device)
        self._dom = engine.DenseIndex(self.domain_dim, dtype, "COSINE", self.device)
        if enable_sparse is None:
            enable_sparse = os.getenv("ENABLE_SPARSE", "1") == "1"      # reference indexing.py:156-158
        self._sparse_block_docs = int(sparse_block_docs)
        self._sparse: Optional[engine.SparseIndex] = None    # created by the first insert that carries sparse vectors
        self._has_domain: Optional[bool] = None              # domain vectors come for every row or for none
        self._tok_fixed: Optional[Tuple[torch.Tensor, torch.Tensor, int]] = None      # set_token_sets
        self._tok_vocab: Dict[str, int] = {}                 # MMR token sets: content.lower().split() (retrieval.py:497)
        self._tok_ptr = _Grow(np.int64)
        self._tok_ptr.extend(np.zeros(1, np.int64))
        self._tok_ids = _Grow(np.int32)
        self._tok_vocab_size = 0                             # > 0: token ids were supplied directly (add(token_sets=...))
        self._tok_dev: Optional[Tuple[torch.Tensor, torch.Tensor, int]] = None
        self._live = _Grow(np.bool_)                         # tombstones: False = deleted, still occupying its row
        self._n_dead = 0
        self._gen = 0                                        # bumped by every mutation; device mirrors carry the generation they saw
        self._dev_cols: Dict[str, Tuple[int, torch.Tensor]] = {}
        self._live_words: Optional[Tuple[int, torch.Tensor]] = None
        self._mask_cache: "OrderedDict[str, Tuple[int, int, Optional[torch.Tensor]]]" = OrderedDict()
        self._pinned: Dict[Tuple, torch.Tensor] = {}
        # CUDA graphs of the dense search chain (prepare, sample, scan, finish, gated fallbacks: ~15 launches), replayed by
        # search_batch_arrays for repeated (collection, batch, k, filter) shapes: the host cost of a search drops to one copy +
        # one launch, which is what bounds batch-1 latency and small shards.  Off when B200RAG_GRAPHS=0.
        self.use_graphs = (os.getenv("B200RAG_GRAPHS", "1") != "0") if use_graphs is None else bool(use_graphs)
        self._graphs: "OrderedDict[Tuple, Any]" = OrderedDict()
        self._graph_seen: Dict[Tuple, int] = {}
        self.collections: Dict[str, _Collection] = {"semantic_index": _Collection("semantic_index", "dense", self),
                                                    "domain_index": _Collection("domain_index", "dense", self)}
        if enable_sparse:
            self.collections["sparse_index"] = _Collection("sparse_index", "sparse", self)

    # ------------------------------------------------------------------------------------------- ingest
    @property
    def num_rows(self) -> int:
        """Live rows (what Collection.num_entities reports)."""
        return len(self.payload) - self._n_dead

    @property
    def n_slots(self) -> int:
        """Rows held by the indexes, deleted ones included (row ids are stable until compact())."""
        return len(self.payload)

    def _sparse_csr(self, n: int, sparse: Any) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """The reference's per-row {"indices","values"} dicts (indexing.py:645-654), a scipy sparse matrix, or a CSR triple
        (indptr, indices, values) -> validated doc-major CSR with ascending indices per row.  Vectorised."""
        if sparse is None:
            return np.zeros(n + 1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32)
        if isinstance(sparse, tuple) and len(sparse) == 3 and torch.is_tensor(sparse[1]) and sparse[1].is_cuda:
            # bulk load from a CSR that already lives on the device (1e8 postings are not round-tripped through the host):
            # checked there, must already be ascending inside every row
            ptr = sparse[0].to("cpu", torch.int64)
            idx, val = sparse[1].to(torch.int64), sparse[2].to(torch.float32)
            if ptr.numel() != n + 1 or int(ptr[0]) != 0 or int(ptr[-1]) != idx.numel() or idx.numel() != val.numel() \
                    or bool((ptr[1:] < ptr[:-1]).any()):
                raise ValueError("sparse CSR is inconsistent with the number of rows")
            if idx.numel():
                if int(idx.min()) < 0 or int(idx.max()) >= self.sparse_dim:
                    raise ValueError(f"sparse index out of range [0, {self.sparse_dim})")
                bad = idx[1:] <= idx[:-1]
                starts = ptr[1:-1]
                starts = starts[(starts > 0) & (starts < idx.numel())].to(idx.device)
                bad[starts - 1] = False
                if bool(bad.any()):
                    raise ValueError("a device CSR must hold ascending unique indices inside every row")
            return ptr, idx, val
        if hasattr(sparse, "tocsr"):
            c = sparse.tocsr()
            ptr, idx, val = c.indptr.astype(np.int64), c.indices.astype(np.int64), c.data.astype(np.float32)
        elif isinstance(sparse, tuple) and len(sparse) == 3:
            ptr, idx, val = (np.asarray(sparse[0], np.int64), np.asarray(sparse[1], np.int64), np.asarray(sparse[2], np.float32))
        else:
            rows = list(sparse)
            if len(rows) != n:
                raise ValueError(f"sparse must hold one entry per row ({len(rows)} != {n})")
            idx_l = [np.asarray(e["indices"], dtype=np.int64).reshape(-1) if e else np.zeros(0, np.int64) for e in rows]
            val_l = [np.asarray(e["values"], dtype=np.float32).reshape(-1) if e else np.zeros(0, np.float32) for e in rows]
            lens = np.asarray([a.size for a in idx_l], dtype=np.int64)
            if any(a.size != b.size for a, b in zip(idx_l, val_l)):
                raise ValueError("sparse entry with different numbers of indices and values")
            ptr = np.zeros(n + 1, np.int64)
            np.cumsum(lens, out=ptr[1:])
            idx = np.concatenate(idx_l) if idx_l else np.zeros(0, np.int64)
            val = np.concatenate(val_l) if val_l else np.zeros(0, np.float32)
        if ptr.shape[0] != n + 1 or ptr[0] != 0 or int(ptr[-1]) != idx.size or idx.size != val.size or (np.diff(ptr) < 0).any():
            raise ValueError("sparse CSR is inconsistent with the number of rows")
        if idx.size and (idx.min() < 0 or idx.max() >= self.sparse_dim):
            raise ValueError(f"sparse index out of range [0, {self.sparse_dim})")
        row = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr))
        order = np.lexsort((idx, row))                       # stable: ascending index inside each row
        return ptr, idx[order], val[order]

    def _token_csr(self, contents: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, Dict[str, int]]:
        """Sorted unique token ids per new row; new vocabulary entries are returned, not applied."""
        vocab, pending = self._tok_vocab, {}
        base = len(vocab)
        ptr = np.zeros(len(contents) + 1, np.int64)
        out: List[int] = []
        for r, text in enumerate(contents):
            ids = set()
            for tok in set(_bm25.tokenize(text)):
                c = vocab.get(tok)
                if c is None:
                    c = pending.get(tok)
                    if c is None:
                        c = pending[tok] = base + len(pending)
                ids.add(c)
            out.extend(sorted(ids))
            ptr[r + 1] = len(out)
        return ptr, np.asarray(out, dtype=np.int32), pending

    def add(self, ids: Sequence[str], contents: Sequence[str], semantic: Any, sparse: Any = None,
            domain: Any = None, metadata: Optional[Sequence[Dict[str, Any]]] = None,
            token_sets: Optional[Tuple[Any, Any, int]] = None) -> None:
        """Append rows to all collections at once (row i of every index is the same chunk, as in the reference where
        the three collections hold the same chunk ids, indexing.py:346-347).

        semantic / domain: fp32 [n, dim] (numpy or torch, host or device); sparse: one {"indices","values"} dict per
        row (the reference's SPARSE_FLOAT_VECTOR payload, indexing.py:645-654), a scipy sparse matrix or a CSR triple, or
        None; token_sets: optional (ptr, ids, vocab_size) CSR of sorted unique token ids per row for MMR, for bulk loads
        that already hold tokenised text (otherwise the contents are tokenised here with content.lower().split()).

        Everything is validated and converted BEFORE any index is touched: a bad row leaves the manager unchanged."""
        n = len(ids)
        if len(contents) != n:
            raise ValueError(f"contents must hold one entry per row ({len(contents)} != {n})")
        sem = semantic if torch.is_tensor(semantic) else torch.as_tensor(np.asarray(semantic, dtype=np.float32))
        if tuple(sem.shape) != (n, self.semantic_dim):
            raise ValueError(f"semantic embeddings must be [{n}, {self.semantic_dim}], got {tuple(sem.shape)}")
        dom = None
        if domain is not None:
            dom = domain if torch.is_tensor(domain) else torch.as_tensor(np.asarray(domain, dtype=np.float32))
            if tuple(dom.shape) != (n, self.domain_dim):
                raise ValueError(f"domain embeddings must be [{n}, {self.domain_dim}], got {tuple(dom.shape)}")
        with self._lock:
            if self._has_domain is not None and self.n_slots > 0 and (dom is not None) != self._has_domain:
                raise ValueError("domain embeddings must be supplied for every row or for none")
            sp_ptr, sp_idx, sp_val = self._sparse_csr(n, sparse)
            if token_sets is not None:
                if self._tok_vocab:
                    raise ValueError("token_sets cannot be mixed with tokenised contents")
                t_ptr = np.asarray(token_sets[0].cpu() if torch.is_tensor(token_sets[0]) else token_sets[0], dtype=np.int64)
                t_ids = np.asarray(token_sets[1].cpu() if torch.is_tensor(token_sets[1]) else token_sets[1], dtype=np.int32)
                if t_ptr.shape[0] != n + 1 or int(t_ptr[-1]) != t_ids.size:
                    raise ValueError("token_sets CSR is inconsistent with the number of rows")
                tok_pending, tok_vocab_size = {}, max(self._tok_vocab_size, int(token_sets[2]))
            else:
                if self._tok_vocab_size:
                    raise ValueError("token_sets cannot be mixed with tokenised contents")
                t_ptr, t_ids, tok_pending = self._token_csr(contents)
                tok_vocab_size = 0
            cols = self.payload.convert(n, metadata)
            if n == 0:
                return
            # ---- commit (GPU appends first: they are the only steps that can still fail, and they roll back)
            lo, hi = self._local_part(self.n_slots, n)             # rows of this batch whose vectors / postings live here
            sem_n, dom_n = self._sem.n, self._dom.n
            try:
                self._append_vectors(sem[lo:hi], dom[lo:hi] if dom is not None else None,
                                     (sp_ptr[lo: hi + 1] - sp_ptr[lo], sp_idx[sp_ptr[lo]: sp_ptr[hi]], sp_val[sp_ptr[lo]: sp_ptr[hi]]),
                                     sparse is not None)
            except Exception:
                self._sem.n, self._dom.n = sem_n, dom_n
                raise
            self._has_domain = dom is not None
            self.payload.commit(list(ids), [c if isinstance(c, str) else str(c or "") for c in contents], cols)
            self._tok_ids.extend(t_ids)
            self._tok_ptr.extend(t_ptr[1:] + self._tok_ptr.view[-1])
            self._tok_vocab.update(tok_pending)
            self._tok_vocab_size = tok_vocab_size
            self._live.extend(np.ones(n, dtype=np.bool_))
            self._invalidate()

    def _local_part(self, first_row: int, n: int) -> Tuple[int, int]:
        """[lo, hi) of a batch of n rows starting at global row first_row that this manager stores vectors for (all of it;
        the row-sharded subclass keeps its own range only)."""
        return 0, n

    def _append_vectors(self, sem: torch.Tensor, dom: Optional[torch.Tensor], csr: Tuple[np.ndarray, np.ndarray, np.ndarray],
                        has_sparse: bool) -> None:
        n = int(sem.shape[0])
        if n == 0:
            return
        self._sem.add(sem)
        if dom is not None:
            self._dom.add(dom)
        if "sparse_index" in self.collections and (has_sparse or self._sparse is not None):
            if self._sparse is None:
                # rows that came before without sparse vectors are empty documents
                pre = np.zeros(self._sem.n - n + 1, np.int64)
                self._sparse = engine.SparseIndex(pre, np.zeros(0, np.int64), np.zeros(0, np.float32), self.sparse_dim,
                                                  self.device, block_docs=self._sparse_block_docs, id_offset=self._sem.id_offset)
            self._sparse.append(*csr)

    def add_vectors(self, semantic: Any, sparse: Any = None, domain: Any = None) -> None:
        """Payload-less bulk ingest: vectors / postings only, rows are addressed by number (ids read f"c{row:09d}").  For
        corpora whose payload lives elsewhere -- bench.py's synthetic 10M / 100M-row shards.  Cannot be mixed with add()."""
        sem = semantic if torch.is_tensor(semantic) else torch.as_tensor(np.asarray(semantic, dtype=np.float32))
        n = int(sem.shape[0])
        if sem.dim() != 2 or sem.shape[1] != self.semantic_dim:
            raise ValueError(f"semantic embeddings must be [n, {self.semantic_dim}], got {tuple(sem.shape)}")
        dom = None
        if domain is not None:
            dom = domain if torch.is_tensor(domain) else torch.as_tensor(np.asarray(domain, dtype=np.float32))
            if tuple(dom.shape) != (n, self.domain_dim):
                raise ValueError(f"domain embeddings must be [{n}, {self.domain_dim}], got {tuple(dom.shape)}")
        with self._lock:
            if self.payload.ids:
                raise ValueError("add_vectors cannot be mixed with add()")
            if self._has_domain is not None and self._sem.n > 0 and (dom is not None) != self._has_domain:
                raise ValueError("domain embeddings must be supplied for every row or for none")
            csr = self._sparse_csr(n, sparse)
            sem_n, dom_n = self._sem.n, self._dom.n
            try:
                self._append_vectors(sem, dom, csr, sparse is not None)
            except Exception:
                self._sem.n, self._dom.n = sem_n, dom_n
                raise
            self._has_domain = dom is not None
            self._grow_virtual(n)
            self._invalidate()

    def _grow_virtual(self, n_new: int) -> None:
        self.payload.set_virtual(self.payload.virtual_rows + n_new)
        self._live.extend(np.ones(n_new, dtype=np.bool_))

    def set_token_sets(self, tok_ptr: Any, tok_ids: Any, vocab_size: int) -> None:
        """MMR token sets for payload-less rows: CSR of sorted unique token ids per row (device tensors are kept as given)."""
        with self._lock:
            ptr = tok_ptr if torch.is_tensor(tok_ptr) else torch.as_tensor(np.asarray(tok_ptr, dtype=np.int64))
            ids = tok_ids if torch.is_tensor(tok_ids) else torch.as_tensor(np.asarray(tok_ids, dtype=np.int32))
            self._tok_fixed = (ptr.to(self.device, torch.int64).contiguous(), ids.to(self.device, torch.int32).contiguous(),
                               int(vocab_size))
            self._tok_dev = None

    def _invalidate(self) -> None:
        self._gen += 1
        self._tok_dev = None
        self._mask_cache.clear()
        self._graphs.clear()                                 # captured pointers / sizes are stale
        self._graph_seen.clear()

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
                if s.shape != (self.semantic_dim,) or d.shape != (self.domain_dim,):
                    raise ValueError(f"embedding shapes {s.shape} / {d.shape} do not match the collection dimensions")
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
            try:
                self.add(ids, texts, np.stack(sem), spa if "sparse_index" in self.collections else None, np.stack(dom), meta)
            except Exception as e:  # noqa: BLE001 - the batch is rejected as a whole and nothing was inserted
                summary["errors"].append({"chunk_id": None, "error": f"batch rejected: {e}"})
                return summary
            summary["indexed_semantic"] = summary["indexed_domain"] = len(ids)
            if "sparse_index" in self.collections:
                summary["indexed_sparse"] = len(ids)
        return summary

    # ------------------------------------------------------------------------------------------- device views
    def token_sets(self) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """Device CSR of every row's sorted unique token ids (input of the MMR kernel)."""
        with self._lock:
            if self._tok_fixed is not None:
                return self._tok_fixed
            if self._tok_dev is None:
                ptr = torch.from_numpy(self._tok_ptr.view.copy()).to(self.device)
                ids = torch.from_numpy(self._tok_ids.view.copy() if self._tok_ids.n else np.zeros(1, np.int32)).to(self.device)
                self._tok_dev = (ptr, ids, max(1, len(self._tok_vocab), self._tok_vocab_size))
            return self._tok_dev

    def _device_column(self, name: str) -> torch.Tensor:
        got = self._dev_cols.get(name)
        if got is None or got[0] != self._gen:
            src = self.payload.codes[name].view if name in _STR_COLS else self.payload.num[name].view
            got = (self._gen, torch.from_numpy(np.ascontiguousarray(src)).to(self.device))
            self._dev_cols[name] = got
        return got[1]

    def _live_mask(self) -> Optional[torch.Tensor]:
        """Bit mask of the rows that are not deleted, or None while nothing is deleted."""
        if self._n_dead == 0:
            return None
        if self._live_words is None or self._live_words[0] != self._gen:
            words = engine.pack_row_mask(torch.from_numpy(self._live.view.copy()).to(self.device))
            self._live_words = (self._gen, words)
        return self._live_words[1]

    def _filter_words(self, expr: Optional[str]) -> Tuple[int, Optional[torch.Tensor]]:
        """(number of allowed rows, bit mask or None for "every row").  The predicate runs on the GPU over the typed device
        columns (b200rag_filter_mask), ANDed with the live-row mask; a small LRU keeps the masks of recent expressions
        until the next insert / delete."""
        n = self.n_slots
        live = self._live_mask()
        if not expr:
            return self.num_rows, live
        cached = self._mask_cache.get(expr)
        if cached is not None and cached[0] == self._gen:
            self._mask_cache.move_to_end(expr)
            return cached[1], cached[2]
        if n == 0:
            return 0, None
        if self.payload.virtual_rows:
            _parse_filter(expr)                                # (still refuses malformed expressions)
            return 0, None                                     # payload-less rows: every scalar field is missing, nothing matches
        terms, keep_alive, and_mask = [], [], live
        for field, op, val in _parse_filter(expr):
            t = _lib.FilterTerm()
            t.op = _OPS[op]
            if field == "chunk_id":
                # the primary key: evaluated on the host (== is one dictionary lookup), handed over as a bit mask
                if not isinstance(val, str):
                    ok = np.zeros(n, dtype=bool)
                elif op == "==":
                    ok = np.zeros(n, dtype=bool)
                    ids = self.payload.ids
                    r = self.payload.row_of(val)
                    while 0 <= r < n:                          # duplicates of a chunk id are consecutive re-inserts at most
                        ok[r] = True
                        try:
                            r = ids.index(val, r + 1)
                        except ValueError:
                            break
                else:
                    fn = _PY_OPS[op]
                    ok = np.fromiter((fn(c, val) for c in self.payload.ids), dtype=bool, count=n)
                words = engine.pack_row_mask(torch.from_numpy(ok).to(self.device))
                and_mask = words if and_mask is None else (and_mask & words)
                continue
            if field in _STR_COLS:
                if not isinstance(val, str):
                    t.kind = _lib.COL_NEVER
                else:
                    t.kind = _lib.COL_CODE
                    col = self._device_column(field)
                    t.column = col.data_ptr()
                    keep_alive.append(col)
                    if op in ("==", "!="):
                        t.ivalue = self.payload.dict_code[field].get(val, -2)
                    else:
                        fn = _PY_OPS[op]
                        values = self.payload.dict_values[field]
                        lut = torch.from_numpy(np.fromiter((fn(v, val) for v in values), dtype=np.uint8, count=len(values))
                                               if values else np.zeros(1, np.uint8)).to(self.device)
                        t.lut, t.lut_size = lut.data_ptr(), len(values)
                        keep_alive.append(lut)
            elif isinstance(val, str):
                t.kind = _lib.COL_NEVER
            else:
                col = self._device_column(field)
                t.column = col.data_ptr()
                keep_alive.append(col)
                if field in _FLOAT_COLS:
                    t.kind, t.fvalue = _lib.COL_F64, float(val)
                elif isinstance(val, float):
                    t.kind, t.fvalue = _lib.COL_I64_AS_F64, float(val)
                else:
                    t.kind, t.ivalue = _lib.COL_I64, int(val)
            terms.append(t)
        words, count = engine.filter_mask(terms, n, self.device, and_mask)
        m = int(count.item())                                  # (synchronises: the tables above may be freed afterwards)
        del keep_alive
        self._mask_cache[expr] = (self._gen, m, words)
        while len(self._mask_cache) > self.MASK_CACHE_SIZE:
            self._mask_cache.popitem(last=False)
        return m, words

    # ------------------------------------------------------------------------------------------- search
    def _dense_of(self, name: str) -> engine.DenseIndex:
        return self._sem if name == "semantic_index" else self._dom

    def _sparse_queries(self, queries: Sequence[Any]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        if isinstance(queries, tuple) and len(queries) == 3 and not isinstance(queries[0], dict):
            qp, qt, qv = (np.asarray(queries[0], np.int64), np.asarray(queries[1], np.int32), np.asarray(queries[2], np.float32))
            if qp.ndim != 1 or qp.size < 1 or qp[0] != 0 or int(qp[-1]) != qt.size or qt.size != qv.size or (np.diff(qp) < 0).any():
                raise ValueError("CSR sparse queries: q_ptr does not describe q_terms / q_vals")
            if qt.size:
                bad = np.diff(qt) <= 0                         # ascending unique term ids inside every query
                starts = qp[1:-1]
                bad[starts[(starts > 0) & (starts < qt.size)] - 1] = False
                if qt.min() < 0 or qt.max() >= self.sparse_dim or bad.any():
                    raise ValueError("CSR sparse queries need ascending unique in-range term ids per query")
            return qp, qt, qv
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

    def _empty(self, b: int, k: int):
        return (torch.full((b, k), float("-inf"), dtype=torch.float64, device=self.device),
                torch.full((b, k), -1, dtype=torch.int64, device=self.device),
                torch.zeros(b, dtype=torch.int32, device=self.device))

    def search_batch_ids(self, queries: Any, collection_name: str, top_k: int = 20, filters: Optional[str] = None
                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Batched search that stays on the device: (scores f64 [B,k], row ids i64 [B,k] (-1 = empty), counts i32 [B]).
        Dense queries: fp32 [B, dim] (numpy, or a torch tensor on the host -- pinned memory makes the copy asynchronous --
        or on the device); sparse queries: a list of {"indices","values"} dicts / scipy rows, or a CSR triple
        (q_ptr, q_terms ascending per query, q_vals).  The metadata predicate and the deleted rows are applied INSIDE the
        scan kernels through one row bit mask (b200rag_*_topk_masked)."""
        if collection_name not in self.collections:
            raise ValueError(f"Collection {collection_name} not found")
        k = int(top_k)
        if k <= 0:
            raise ValueError("top_k must be positive")
        with self._lock:
            m, words = self._filter_words(filters)
            return self._search_masked(queries, collection_name, k, m, words)

    def _search_masked(self, queries: Any, collection_name: str, k: int, m: int, words: Optional[torch.Tensor]):
        """The search itself, given the number of allowed rows and their bit mask (None = every row)."""
        if collection_name in _DENSE:
            idx = self._dense_of(collection_name)
            q = queries if torch.is_tensor(queries) else torch.as_tensor(np.asarray(queries, dtype=np.float32))
            if q.dim() == 1:
                q = q[None, :]
            b = q.shape[0]
            if idx.n == 0 or m == 0:
                return self._empty(b, k)
            s, i, _ = idx.search(q, k, row_mask=words)
            return s, i, torch.full((b,), min(k, m), dtype=torch.int32, device=self.device)
        qp, qt, qv = self._sparse_queries(queries)
        b = qp.shape[0] - 1
        if self._sparse is None or self._sparse.nnz == 0 or m == 0:
            return self._empty(b, k)
        s, i, c = self._sparse.search(qp, qt, qv, k, doc_mask=words)
        return s.to(torch.float64), i, c

    GRAPH_CACHE_SIZE = 4
    GRAPH_AFTER_CALLS = 3                                    # a shape is captured when it shows up for the third time

    def _search_for_host(self, queries: Any, collection_name: str, k: int, filters: Optional[str]):
        """search_batch_ids for callers that copy the result to the host right away (search_batch_arrays): dense searches of a
        repeated shape are replayed from a CUDA graph.  The graph's outputs are reused by the next replay, which is why this
        path is not offered to callers that keep device tensors."""
        if (not self.use_graphs or collection_name not in _DENSE or collection_name not in self.collections or k <= 0):
            return self.search_batch_ids(queries, collection_name, k, filters)
        idx = self._dense_of(collection_name)
        q = queries if torch.is_tensor(queries) else torch.as_tensor(np.asarray(queries, dtype=np.float32))
        if q.dim() == 1:
            q = q[None, :]
        if idx.n == 0 or q.dim() != 2 or q.shape[1] != idx.dim or q.shape[0] == 0:
            return self.search_batch_ids(q, collection_name, k, filters)
        m, words = self._filter_words(filters)
        if m == 0:
            return self._search_masked(q, collection_name, k, m, words)
        stream = torch.cuda.current_stream(self.device)
        key = (collection_name, int(q.shape[0]), k, filters or "", self._gen, stream.cuda_stream, threading.get_ident())
        entry = self._graphs.get(key)
        if entry is None:
            seen = self._graph_seen.get(key, 0) + 1
            self._graph_seen[key] = seen
            if seen < self.GRAPH_AFTER_CALLS or seen > self.GRAPH_AFTER_CALLS + 2:      # (a failed capture is not retried forever)
                return self._search_masked(q, collection_name, k, m, words)
            entry = self._capture_search(q, collection_name, k, m, words)
            if entry is None:
                return self._search_masked(q, collection_name, k, m, words)
            self._graphs[key] = entry
            while len(self._graphs) > self.GRAPH_CACHE_SIZE:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        graph, q_static, out, _keep = entry
        q_static.copy_(q, non_blocking=True)
        graph.replay()
        return out

    def _capture_search(self, q: torch.Tensor, collection_name: str, k: int, m: int, words: Optional[torch.Tensor]):
        """Capture one dense search into a CUDA graph (static fp32 query buffer in, static result tensors out)."""
        q_static = torch.empty(tuple(q.shape), dtype=torch.float32, device=self.device)
        q_static.copy_(q, non_blocking=True)
        bufs = engine._WS.begin_capture()
        try:
            torch.cuda.current_stream(self.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._search_masked(q_static, collection_name, k, m, words)
        except Exception:  # noqa: BLE001 - anything that cannot be captured keeps running eagerly
            return None
        finally:
            engine._WS.end_capture()
        return graph, q_static, out, (list(bufs), words)

    def _pinned_like(self, tag: str, t: torch.Tensor) -> torch.Tensor:
        key = (tag, tuple(t.shape), t.dtype, threading.get_ident())
        buf = self._pinned.get(key)
        if buf is None:
            if len(self._pinned) > 64:
                self._pinned.clear()
            buf = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            self._pinned[key] = buf
        return buf

    def search_batch_arrays(self, queries: Any, collection_name: str, top_k: int = 20, filters: Optional[str] = None
                            ) -> SearchArrays:
        """The columnar plugin call: the batched search with HOST results -- numpy rows / scores / counts, copied through
        pinned buffers with one synchronisation -- and the payload one gather away (`.hits()`, `.chunk_ids()`)."""
        with self._lock:
            s, i, c = self._search_for_host(queries, collection_name, int(top_k), filters)
            hs, hi, hc = self._pinned_like("s", s), self._pinned_like("i", i), self._pinned_like("c", c)
            hs.copy_(s, non_blocking=True)
            hi.copy_(i, non_blocking=True)
            hc.copy_(c, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return SearchArrays(hi.numpy().copy(), hs.numpy().copy(), hc.numpy().copy(), self.payload)

    def search_batch(self, queries: Any, collection_name: str, top_k: int = 20, filters: Optional[str] = None,
                     search_params: Optional[Dict] = None) -> List[List[Dict[str, Any]]]:
        """One hit list per query, each in the reference's result format (indexing.py:534-551), best first."""
        return self.search_batch_arrays(queries, collection_name, top_k, filters).hits().materialize()

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

    def compact(self) -> None:
        """Drop the deleted rows from every index (row numbers change; chunk ids do not)."""
        with self._lock:
            if self._n_dead == 0:
                return
            if self.payload.virtual_rows:
                raise ValueError("payload-less rows are addressed by row number: they cannot be compacted")
            live = self._live.view.copy()
            rows = np.flatnonzero(live)
            dev_rows = torch.from_numpy(rows).to(self.device)
            for name in ("_sem", "_dom"):
                old = getattr(self, name)
                if old.n == 0:
                    continue
                new = engine.DenseIndex(old.dim, self.dtype, "COSINE", self.device, capacity=rows.size)
                if rows.size:
                    new.add_prepared(old.rows[dev_rows])
                setattr(self, name, new)
            if self._sparse is not None:
                dp, ti, w = self._sparse.to_doc_major()
                lens = (dp[1:] - dp[:-1])[dev_rows]
                new_ptr = torch.zeros(rows.size + 1, dtype=torch.int64, device=self.device)
                new_ptr[1:] = torch.cumsum(lens, 0)
                tot = int(new_ptr[-1])
                src = torch.repeat_interleave(dp[dev_rows] - new_ptr[:-1], lens, output_size=tot) + torch.arange(tot, device=self.device)
                self._sparse = engine.SparseIndex(new_ptr, ti[src], w[src], self.sparse_dim, self.device,
                                                  block_docs=self._sparse_block_docs)
            t_ptr, (t_ids,) = _csr_take(self._tok_ptr.view, [self._tok_ids.view], rows)
            self._tok_ptr.replace(t_ptr)
            self._tok_ids.replace(t_ids)
            self.payload.keep(rows)
            self._live.replace(np.ones(rows.size, dtype=np.bool_))
            self._n_dead = 0
            self._invalidate()

    async def delete_by_filter(self, collection_name: str, expr: str):
        """Reference MilvusIndexManager.delete_by_filter (indexing.py:692-695: `collection.delete(expr)`).  The three
        collections hold the same rows here, so the matching rows leave ALL of them (the reference deletes per collection;
        a chunk missing from one index but not the others is not a state this engine represents).  Deleted rows become
        tombstones in the row mask every search already applies inside the kernels; the indexes are compacted once more
        than half of the rows are dead (or on compact()).  Returns the number of rows deleted."""
        if collection_name not in self.collections:
            raise ValueError(f"Collection {collection_name} not found")
        with self._lock:
            m, words = self._filter_words(expr)
            if m == 0 or words is None:
                return 0
            w = words.cpu().numpy().view(np.uint32)
            hit = np.unpackbits(w.view(np.uint8), bitorder="little")[: self.n_slots].astype(bool)
            live = self._live.view
            live[hit] = False
            self._n_dead = int((~live).sum())
            self._invalidate()
            if self._n_dead * 2 > self.n_slots and not self.payload.virtual_rows and self._compactable:
                self.compact()
            return m

    # ------------------------------------------------------------------------------------------- checkpoint / resume
    def save(self, path: str) -> None:
        """Write the whole index (stored 16-bit rows, sparse CSR, token sets, payload columns) to one .npz-format file: plain
        arrays plus JSON-encoded strings, nothing pickled.  The reference leaves durability to the Milvus server
        (collection.flush, indexing.py:430-431).  Deleted rows are compacted away first."""
        with self._lock:
            if not self.payload.virtual_rows:
                self.compact()
            elif self._n_dead:
                raise ValueError("cannot checkpoint payload-less rows with deletions")
            def blob(obj) -> np.ndarray:
                return np.frombuffer(json.dumps(obj, ensure_ascii=False).encode("utf-8"), dtype=np.uint8)
            if self._sparse is not None:
                dp, ti, w = (t.cpu().numpy() for t in self._sparse.to_doc_major())
            else:
                dp, ti, w = np.zeros(self.n_slots + 1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32)
            arrays = {
                "header": blob({"version": 2, "dtype": self.dtype, "dims": [self.semantic_dim, self.sparse_dim, self.domain_dim],
                                "sparse_enabled": "sparse_index" in self.collections, "sparse_built": self._sparse is not None,
                                "tok_vocab_size": self._tok_vocab_size, "rows": self.n_slots,
                                "virtual": bool(self.payload.virtual_rows)}),
                "sem": self._sem.rows.view(torch.int16).cpu().numpy(), "dom": self._dom.rows.view(torch.int16).cpu().numpy(),
                "sp_ptr": dp, "sp_idx": ti, "sp_val": w,
                "tok_ptr": self._tok_ptr.view, "tok_ids": self._tok_ids.view,
                "tok_vocab": blob(sorted(self._tok_vocab, key=self._tok_vocab.get)),
                "ids": blob(self.payload.ids), "content": blob(self.payload.content),
            }
            for f in _FLOAT_COLS + _INT_COLS:
                arrays["col_" + f] = self.payload.num[f].view
            for f in _STR_COLS:
                arrays["col_" + f] = self.payload.codes[f].view
                arrays["dict_" + f] = blob(self.payload.dict_values[f])
            with open(path, "wb") as fh:
                np.savez(fh, **arrays)

    @classmethod
    def load(cls, path: str, device: str = "cuda", **kwargs) -> "B200IndexManager":
        def unblob(a: np.ndarray):
            return json.loads(a.tobytes().decode("utf-8"))
        with np.load(path, allow_pickle=False) as z:
            hd = unblob(z["header"])
            if hd.get("version") != 2:
                raise ValueError("unknown index file version")
            sd, pd, dd = hd["dims"]
            m = cls(semantic_dim=sd, sparse_dim=pd, domain_dim=dd, device=device, dtype=hd["dtype"],
                    enable_sparse=hd["sparse_enabled"], **kwargs)
            n = int(hd["rows"])
            virtual = bool(hd.get("virtual"))
            ids, content = unblob(z["ids"]), unblob(z["content"])
            sem, dom = z["sem"], z["dom"]
            if (not virtual and (len(ids) != n or len(content) != n)) or sem.shape != (n, sd) or dom.shape not in ((n, dd), (0, dd)):
                raise ValueError("index file is inconsistent (row counts / dimensions)")
            tdt = engine._TORCH_DTYPE[engine.dtype_code(hd["dtype"])]
            if n:
                m._sem.add_prepared(torch.from_numpy(sem).view(tdt))
            if dom.shape[0]:
                m._dom.add_prepared(torch.from_numpy(dom).view(tdt))
            sp_ptr, sp_idx, sp_val = z["sp_ptr"], z["sp_idx"], z["sp_val"]
            if hd["sparse_built"]:
                if sp_ptr.shape[0] != n + 1 or int(sp_ptr[-1]) != sp_idx.size or sp_idx.size != sp_val.size:
                    raise ValueError("index file is inconsistent (sparse CSR)")
                m._sparse = engine.SparseIndex(sp_ptr, sp_idx, sp_val, pd, m.device, block_docs=m._sparse_block_docs)
            tok_ptr, tok_ids = z["tok_ptr"].astype(np.int64), z["tok_ids"].astype(np.int32)
            if tok_ptr.shape[0] != n + 1 or int(tok_ptr[-1]) != tok_ids.size:
                raise ValueError("index file is inconsistent (token sets)")
            m._tok_ptr.replace(tok_ptr)
            m._tok_ids.replace(tok_ids)
            m._tok_vocab = {t: i for i, t in enumerate(unblob(z["tok_vocab"]))}
            m._tok_vocab_size = int(hd["tok_vocab_size"])
            m._has_domain = bool(dom.shape[0]) if n else None
            st = m.payload
            if virtual:
                st.set_virtual(n)
            st.ids, st.content = ([], []) if virtual else (ids, content)
            for f in () if virtual else _FLOAT_COLS + _INT_COLS:
                col = z["col_" + f]
                if col.shape != (n,):
                    raise ValueError("index file is inconsistent (payload columns)")
                st.num[f].replace(col.astype(np.float64 if f in _FLOAT_COLS else np.int64))
            for f in () if virtual else _STR_COLS:
                col, values = z["col_" + f].astype(np.int32), unblob(z["dict_" + f])
                if col.shape != (n,) or (col.size and int(col.max()) >= len(values)):
                    raise ValueError("index file is inconsistent (dictionary columns)")
                st.codes[f].replace(col)
                st.dict_values[f] = values
                st.dict_code[f] = {v: i for i, v in enumerate(values)}
            m._live.replace(np.ones(n, dtype=np.bool_))
            m._invalidate()
        return m

    async def close(self):
        if self._batcher is not None:
            self._batcher.close()
        with self._lock:
            self._sparse = None
            self._tok_dev = None
            self._dev_cols.clear()
            self._mask_cache.clear()
            self._pinned.clear()
            self._graphs.clear()
