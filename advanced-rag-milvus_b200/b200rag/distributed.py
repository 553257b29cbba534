"""Row-sharded search across the GPUs of one box (one process per GPU, torch.distributed).

The corpus is cut into contiguous row ranges, one per rank (the reference's own knob is the
Milvus collection's num_shards, src/advanced_rag/indexing.py:91,234-239); every rank scans its shard for the exact
local top-k, ONE all-gather carries the k (score, id) candidates per query and rank, and the merge kernel (K4) reduces
G*k -> k by (score desc, id asc) on every rank.  Scores travel as fp64 so the merge compares the same canonical values
the single-GPU path ranks by -- the merged result is bit-identical to a single-shard search.

The message is a [2, B, k] i64 buffer: a plane of fp64 score bit patterns and a plane of ids.  The local search writes
its two output arrays straight into the planes (`out=`), so nothing is packed between the search and the collective, and
the merge kernel reads the gathered [G, 2, B, k] buffer in place.

    ShardedDenseIndex     one dense collection, sharded
    ShardedIndexManager   the whole B200IndexManager surface, sharded: dense AND sparse collections hold this rank's rows
                          only; payload columns, token sets and row masks are replicated (they are small next to the vectors);
                          every search ends with the same all-gather + merge.  RRF / MMR run on the merged lists, split by
                          query across the ranks (b200rag/retriever.py).

The exchange is the only collective on the path (NCCL over NVLink/NVSwitch on the GPU box; the same code runs over gloo in
the CPU tests).  It is latency bound: B=1024, k=100 is 1.6 MB per rank.
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def shard_range(n_total: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous balanced row range [start, end) of `rank`.  align = 1: the first n_total % world ranks hold one extra row;
    align > 1: ranges are cut at multiples of `align` (whole postings blocks / whole 32-row mask words per shard)."""
    units = -(-n_total // align)
    base, extra = divmod(units, world)
    start = rank * base + min(rank, extra)
    end = start + base + (1 if rank < extra else 0)
    return min(n_total, start * align), min(n_total, end * align)


def pack_candidates(scores: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """(f64 [B,k], i64 [B,k]) -> one i64 [2, B, k] message (a plane of score bit patterns, a plane of ids)."""
    return torch.stack([scores.contiguous().view(torch.int64), ids.contiguous()])


def unpack_gathered(gathered: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """i64 [G, 2, B, k] -> candidate lists (f64 [B, G*k], i64 [B, G*k]), rank-major inside a query."""
    g, _, b, _ = gathered.shape
    sc = gathered[:, 0].permute(1, 0, 2).reshape(b, g * k).contiguous().view(torch.float64)
    ids = gathered[:, 1].permute(1, 0, 2).reshape(b, g * k).contiguous()
    return sc, ids


class SendBuffer:
    """Persistent [2, B, k] i64 message per (B, k) whose planes the local search writes into (`planes()` -> (f64 view,
    i64 view) to pass as `out=`)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._buf = {}

    def planes(self, b: int, k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        msg = self._buf.get((b, k))
        if msg is None:
            if len(self._buf) > 8:
                self._buf.clear()
            msg = torch.empty((2, b, k), dtype=torch.int64, device=self.device)
            self._buf[(b, k)] = msg
        return msg, msg[0].view(torch.float64), msg[1]


def gather_and_merge(scores: torch.Tensor, ids: torch.Tensor, k: int,
                     merge_fn: Callable[[torch.Tensor, torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]],
                     group: Optional[dist.ProcessGroup] = None,
                     merge_gathered_fn: Optional[Callable[[torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]]] = None,
                     message: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather every rank's local top-k and merge.  Identity when not running distributed.

    message: the [2, B, k] buffer `scores` / `ids` already live in (SendBuffer) -- then nothing is packed.
    merge_gathered_fn (engine.merge_gathered on the GPU) reduces the gathered [G, 2, B, k] buffer in place; without it the
    buffer is unpacked into [B, G*k] candidate lists for merge_fn (the CPU tests stand the oracle in there)."""
    world = _world(group)
    if world == 1:
        return scores, ids
    msg = message if message is not None else pack_candidates(scores, ids)
    b = msg.shape[1]
    # concatenated (not stacked) output layout: the one form both NCCL and gloo accept
    out = torch.empty((world * 2 * b, k), dtype=msg.dtype, device=msg.device)
    dist.all_gather_into_tensor(out, msg.view(2 * b, k), group=group)
    gathered = out.view(world, 2, b, k)
    if merge_gathered_fn is not None:
        return merge_gathered_fn(gathered, k)
    cs, ci = unpack_gathered(gathered, k)
    return merge_fn(cs, ci, k)


class ShardedDenseIndex:
    """DenseIndex over this rank's row range of a corpus of `n_total` rows."""

    def __init__(self, dim: int, n_total: int, dtype="f16", metric: str = "COSINE", device="cuda",
                 group: Optional[dist.ProcessGroup] = None):
        from . import engine
        self._engine = engine
        self.group = group
        self.rank, self.world = _rank(group), _world(group)
        self.n_total = n_total
        self.start, self.end = shard_range(n_total, self.rank, self.world)
        self.local = engine.DenseIndex(dim, dtype, metric, device, id_offset=self.start, capacity=self.end - self.start)
        self._send = SendBuffer(device)

    def search(self, queries_f32: torch.Tensor, k: int, mode: Optional[int] = None):
        """Replicated queries -> global exact top-k on every rank: (scores f64 [B,k], ids i64 [B,k])."""
        eng = self._engine
        mode = eng.DENSE_AUTO if mode is None else mode
        if self.world == 1:
            s, i, _ = self.local.search(queries_f32, k, mode)
            return s, i
        q16 = self.local.prepare_queries(queries_f32)
        msg, ps, pi = self._send.planes(q16.shape[0], k)
        self.local.search_prepared(q16, k, mode, out=(ps, pi))
        return gather_and_merge(ps, pi, k, eng.merge_topk, self.group, eng.merge_gathered, message=msg)


def bm25_global_stats(doc_ptr, term_ids, tf, n_terms: int, group: Optional[dist.ProcessGroup] = None):
    """(n_docs, df i64 [n_terms], total length) of the WHOLE corpus from this rank's doc-major tf CSR: three all-reduces at
    build time.  BM25 weights computed from them (bm25.bm25_weights*(..., stats=...)) are bit-identical to a single-shard
    build, so a document scores the same on whichever GPU holds it (idf and avgdl are global)."""
    t = torch.as_tensor(term_ids).to(torch.int64)
    f = torch.as_tensor(tf).to(torch.int64)
    n_local = int(torch.as_tensor(doc_ptr).numel() - 1)
    df = torch.bincount(t, minlength=n_terms) if t.numel() else torch.zeros(n_terms, dtype=torch.int64, device=t.device)
    scal = torch.stack([torch.tensor(n_local, dtype=torch.int64, device=df.device), f.sum()])
    if _world(group) > 1:
        dist.all_reduce(df, group=group)
        dist.all_reduce(scal, group=group)
    return int(scal[0]), df, int(scal[1])


def _popcount_words(words: torch.Tensor) -> int:
    table = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=words.device)
    return int(table[words.contiguous().view(torch.uint8).to(torch.int64)].sum().item())


def make_sharded_manager_class():
    """ShardedIndexManager is a subclass of B200IndexManager; built lazily so that importing b200rag.distributed on a box
    without the CUDA library (the CPU tests exercise the collective helpers above) does not load it."""
    from . import engine
    from .index_manager import B200IndexManager

    class ShardedIndexManager(B200IndexManager):
        """B200IndexManager whose dense and sparse collections hold one contiguous row range per rank.

        Two ways to load it, both called on EVERY rank:
          add(ids, contents, semantic, sparse, domain, metadata)   the same global batch everywhere: each rank keeps the
              vectors / postings of its own range and the payload columns + token sets of all rows (replicated);
          add_vectors(semantic_local, sparse_local, domain_local)  this rank's rows only, payload-less (bulk benchmark loads).
        Searches return the GLOBAL exact top-k on every rank.  Row ranges are fixed, so deletes stay tombstones
        (compact / save are per-process operations of the single-GPU manager and are refused here)."""

        _compactable = False

        def __init__(self, n_total: int, *args, group: Optional[dist.ProcessGroup] = None, **kwargs):
            kwargs.setdefault("use_graphs", False)             # (the search chain contains a collective; opt in explicitly)
            super().__init__(*args, **kwargs)
            self.group = group
            self.rank, self.world = _rank(group), _world(group)
            self.n_total = int(n_total)
            # whole postings blocks (and whole mask words) per shard
            self.start, self.end = shard_range(self.n_total, self.rank, self.world, align=max(32, self._sparse_block_docs))
            cap = self.end - self.start
            self._sem = engine.DenseIndex(self.semantic_dim, self.dtype, "COSINE", self.device, id_offset=self.start, capacity=cap)
            self._dom = engine.DenseIndex(self.domain_dim, self.dtype, "COSINE", self.device, id_offset=self.start)
            self._send = SendBuffer(self.device)

        def _local_part(self, first_row: int, n: int) -> Tuple[int, int]:
            if first_row + n > self.n_total:
                raise ValueError(f"more rows than the declared corpus size ({first_row + n} > {self.n_total})")
            return min(max(self.start - first_row, 0), n), min(max(self.end - first_row, 0), n)

        def _grow_virtual(self, n_new: int) -> None:
            if self._sem.n > self.end - self.start:
                raise ValueError("more local rows than this rank's range holds")
            if not self.payload.virtual_rows:                  # declared once for the whole corpus
                self.payload.set_virtual(self.n_total)
                self._live.extend(np.ones(self.n_total, dtype=np.bool_))

        def _local_words(self, m: int, words: Optional[torch.Tensor]):
            """This rank's slice of the replicated global row mask and the number of rows it allows here."""
            if words is not None:
                words = words[self.start // 32: (self.end + 31) // 32].contiguous()
                return (_popcount_words(words) if m else 0), words
            return self._sem.n, None                           # no filter, nothing deleted: every local row

        def _local_dense(self, q: torch.Tensor, collection_name: str, k: int, m_local: int, words: Optional[torch.Tensor]):
            """Exact local top-k of this rank's row range, written by the finish kernel straight into the [2, B, k] message."""
            idx = self._dense_of(collection_name)
            msg, ps, pi = self._send.planes(q.shape[0], k)
            if idx.n == 0 or m_local == 0:
                ps.fill_(float("-inf"))
                pi.fill_(-1)
            else:
                idx.search(q, k, row_mask=words, out=(ps, pi))
            return msg, ps, pi

        def _search_masked(self, queries: Any, collection_name: str, k: int, m: int, words: Optional[torch.Tensor]):
            m_local, words = self._local_words(m, words)
            if self.world == 1:
                return super()._search_masked(queries, collection_name, k, m_local, words)
            if collection_name != "sparse_index":
                q = queries if torch.is_tensor(queries) else torch.as_tensor(np.asarray(queries, dtype=np.float32))
                q = q[None, :] if q.dim() == 1 else q
                msg, ps, pi = self._local_dense(q, collection_name, k, m_local, words)
            else:
                s, i, _ = super()._search_masked(queries, collection_name, k, m_local, words)
                msg, ps, pi = self._send.planes(s.shape[0], k)
                ps.copy_(s)
                pi.copy_(i)
            return gather_and_merge(ps, pi, k, engine.merge_topk, self.group,
                                    lambda g, kk: engine.merge_gathered(g, kk, with_counts=True), message=msg)

        def search_own_queries_arrays(self, queries_local: Any, collection_name: str, top_k: int = 20,
                                      filters: Optional[str] = None):
            """The SPMD front door of a row-sharded dense collection: EVERY rank calls it with ITS OWN queries (host or device
            fp32 [per, dim], the same `per` on every rank -- each rank serves its own clients) and gets the global exact
            top-k of those queries back as host arrays (SearchArrays, like search_batch_arrays).

            The queries are all-gathered over NVLink (instead of every rank copying the whole batch from its host), every rank
            searches its row range for the whole batch, and ONE all-to-all hands each rank the candidates of its own queries
            from all row ranges, so a rank merges and copies back 1/world of the result."""
            from .index_manager import SearchArrays
            if collection_name not in self.collections:
                raise ValueError(f"Collection {collection_name} not found")
            if collection_name == "sparse_index":
                raise ValueError("search_own_queries_arrays serves the dense collections")
            k = int(top_k)
            with self._lock:
                q = queries_local if torch.is_tensor(queries_local) else torch.as_tensor(np.asarray(queries_local, dtype=np.float32))
                q = q[None, :] if q.dim() == 1 else q
                q = q.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
                per, world = q.shape[0], self.world
                if world == 1:
                    return self.search_batch_arrays(q, collection_name, k, filters)
                q_all = torch.empty((world * per, q.shape[1]), dtype=torch.float32, device=self.device)
                dist.all_gather_into_tensor(q_all, q, group=self.group)
                m, words = self._filter_words(filters)
                m_local, words = self._local_words(m, words)
                msg, _, _ = self._local_dense(q_all, collection_name, k, m_local, words)
                # [2, world * per, k] planes -> one [2, per, k] block per destination rank
                send = msg.view(2, world, per, k).permute(1, 0, 2, 3).contiguous()
                recv = torch.empty_like(send)
                dist.all_to_all_single(recv, send, group=self.group)
                ms, mi, cnt = engine.merge_gathered(recv, k, with_counts=True)
                hs, hi, hc = self._pinned_like("s", ms), self._pinned_like("i", mi), self._pinned_like("c", cnt)
                hs.copy_(ms, non_blocking=True)
                hi.copy_(mi, non_blocking=True)
                hc.copy_(cnt, non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
                return SearchArrays(hi.numpy().copy(), hs.numpy().copy(), hc.numpy().copy(), self.payload)

        def compact(self) -> None:
            if self._n_dead:
                raise NotImplementedError("row ranges of a sharded index are fixed: deleted rows stay tombstones")

        def save(self, path: str) -> None:
            raise NotImplementedError("checkpoint the shards with one B200IndexManager.save per rank")

    return ShardedIndexManager


def ShardedIndexManager(*args, **kwargs):
    """Construct a row-sharded B200IndexManager (see make_sharded_manager_class)."""
    return make_sharded_manager_class()(*args, **kwargs)
