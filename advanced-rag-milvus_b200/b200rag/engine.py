"""Device-resident indexes and batched kernels: thin torch-tensor front-end of the C ABI (include/b200rag.h).

PyTorch is plumbing here (device memory, streams); every computation is a hand-written CUDA kernel in
libb200rag.so.  There is no CPU path: tensors must live on a CUDA device and the library must be built.

    DenseIndex   row shard of fp16/bf16 vectors  -> exact top-k      (replaces Collection.search on the
                 semantic / domain collections, reference src/advanced_rag/indexing.py:505-523)
    SparseIndex  doc-range-blocked postings      -> sparse IP top-k  (replaces Collection.search on sparse_index)
    rrf_fuse / mmr_select / merge_topk           -> reference retrieval.py:421-491 / :493-516 / shard reduce
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BF16, DENSE_APPROX, DENSE_AUTO, DENSE_EXACT, DENSE_TENSOR, F16, check  # noqa: F401

_TORCH_DTYPE = {F16: torch.float16, BF16: torch.bfloat16}
_DTYPE_CODE = {"f16": F16, "fp16": F16, "float16": F16, torch.float16: F16,
               "bf16": BF16, "bfloat16": BF16, torch.bfloat16: BF16, F16: F16, BF16: BF16}


def dtype_code(dtype) -> int:
    try:
        return _DTYPE_CODE[dtype]
    except KeyError:
        raise ValueError(f"unsupported vector dtype {dtype!r} (use 'f16' or 'bf16')") from None


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (b200rag has no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


class _Workspace:
    """Grow-only scratch buffers handed to the C ABI (the library never allocates), one per (device, CUDA stream, host
    thread): the multi-kernel sequences behind dense_topk / sparse_topk keep state in their workspace between launches, so
    two host threads -- or two streams -- must never share one (the micro-batcher's worker thread and the event
    loop thread used to).  Work queued on ONE stream by one thread is ordered by the stream."""

    def __init__(self):
        self._buf = {}
        self._lock = threading.Lock()
        self._capture = threading.local()       # .bufs: list collecting the buffers handed out while a CUDA graph is captured

    def get(self, device: torch.device, nbytes: int) -> torch.Tensor:
        bufs = getattr(self._capture, "bufs", None)
        if bufs is not None:
            # CUDA-graph capture (B200IndexManager): the scratch must belong to the graph -- allocated from its private pool,
            # kept alive by it, never shared with eager calls or other graphs
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            bufs.append(buf)
            return buf
        key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream, threading.get_ident())
        with self._lock:
            buf = self._buf.get(key)
            if buf is None or buf.numel() < nbytes:
                buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
                self._buf[key] = buf
            return buf

    def release(self) -> None:
        with self._lock:
            self._buf.clear()

    def begin_capture(self) -> list:
        self._capture.bufs = []
        return self._capture.bufs

    def end_capture(self) -> None:
        self._capture.bufs = None


_WS = _Workspace()


def device_info() -> Tuple[int, int, int]:
    import ctypes
    sm, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(_lib.load().b200rag_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr)))
    return sm.value, maj.value, mnr.value


def prepare_rows(x_f32: torch.Tensor, dtype, normalize: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 rows -> stored 16-bit rows (canonical L2 normalisation when `normalize`)."""
    code = dtype_code(dtype)
    x = x_f32.to(torch.float32).contiguous()
    _require_cuda(x, "rows")
    if x.dim() != 2:
        raise ValueError("rows must be [n, dim]")
    if out is None:
        out = torch.empty(x.shape, dtype=_TORCH_DTYPE[code], device=x.device)
    else:
        _require_cuda(out, "out")
        if out.shape != x.shape or out.dtype != _TORCH_DTYPE[code]:
            raise ValueError("out has the wrong shape or dtype")
    with torch.cuda.device(x.device):
        check(_lib.load().b200rag_prepare_rows(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], code,
                                               1 if normalize else 0, _stream_ptr(x.device)))
    return out


def pack_row_mask(allowed: torch.Tensor) -> torch.Tensor:
    """bool [n] (device) -> the bit mask the *_masked entry points take: int32 [ceil(n / 32)], bit (row & 31) of word row >> 5."""
    _require_cuda(allowed, "allowed")
    n = allowed.numel()
    pad = (-n) % 32
    bits = torch.nn.functional.pad(allowed.to(torch.int64).view(-1), (0, pad)).view(-1, 32)
    weights = torch.ones(32, dtype=torch.int64, device=allowed.device) << torch.arange(32, device=allowed.device)
    words = (bits * weights).sum(dim=1)                              # 0 .. 2^32 - 1
    return torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).contiguous()


def filter_mask(terms: Sequence["_lib.FilterTerm"], n_rows: int, device, and_mask: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Parsed predicate terms over device columns -> (bit mask int32 [ceil(n/32)], number of allowed rows as a device i64
    scalar).  The columns / tables the terms point to must stay alive until the stream has run the kernel."""
    import ctypes
    dev = torch.device(device)
    words = torch.empty(((n_rows + 31) // 32,), dtype=torch.int32, device=dev)
    count = torch.empty((1,), dtype=torch.int64, device=dev)
    arr = (_lib.FilterTerm * max(1, len(terms)))(*terms)
    with torch.cuda.device(dev):
        check(_lib.load().b200rag_filter_mask(ctypes.cast(arr, ctypes.c_void_p), len(terms), n_rows,
                                              and_mask.data_ptr() if and_mask is not None else None,
                                              words.data_ptr(), count.data_ptr(), _stream_ptr(dev)))
    return words, count


def dense_topk(corpus16: torch.Tensor, queries16: torch.Tensor, k: int, id_offset: int = 0, mode: int = DENSE_AUTO,
               n_rows: Optional[int] = None, row_norm_bound: float = 1.001, out_err: Optional[torch.Tensor] = None,
               row_mask: Optional[torch.Tensor] = None, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Exact top-k of stored rows (of the rows `row_mask` allows, see pack_row_mask).  Returns (scores f64 [B,k],
    ids i64 [B,k], flags i32 [B]).  `out` = preallocated (scores, ids) to write into (e.g. the two planes of an all-gather
    send buffer)."""
    _require_cuda(corpus16, "corpus")
    _require_cuda(queries16, "queries")
    code = dtype_code(corpus16.dtype)
    if queries16.dtype != corpus16.dtype:
        raise ValueError("queries and corpus must share the 16-bit dtype")
    n = corpus16.shape[0] if n_rows is None else n_rows
    dim = corpus16.shape[1]
    b = queries16.shape[0]
    if queries16.shape[1] != dim:
        raise ValueError(f"query dim {queries16.shape[1]} != corpus dim {dim}")
    dev = corpus16.device
    if out is not None:
        scores, ids = out
        if (scores.dtype, ids.dtype) != (torch.float64, torch.int64) or tuple(scores.shape) != (b, k) or tuple(ids.shape) != (b, k) \
                or not (scores.is_contiguous() and ids.is_contiguous() and scores.is_cuda and ids.is_cuda):
            raise ValueError("out must be contiguous CUDA (f64 [B,k], i64 [B,k]) tensors")
    else:
        scores = torch.empty((b, k), dtype=torch.float64, device=dev)
        ids = torch.empty((b, k), dtype=torch.int64, device=dev)
    if b == 0:
        return scores, ids, torch.zeros((0,), dtype=torch.int32, device=dev)
    flags = torch.empty((b,), dtype=torch.int32, device=dev)      # every mode writes all b entries
    if out_err is not None and (out_err.dtype != torch.float32 or out_err.numel() < b or not out_err.is_cuda):
        raise ValueError("out_err must be a CUDA float32 tensor with one entry per query")
    if row_mask is not None:
        _require_cuda(row_mask, "row_mask")
        if row_mask.dtype != torch.int32 or row_mask.numel() < (n + 31) // 32:
            raise ValueError("row_mask must be an int32 tensor with one bit per row (engine.pack_row_mask)")
    L = _lib.load()
    with torch.cuda.device(dev):
        nbytes = L.b200rag_dense_topk_workspace_bytes(n, dim, b, k, mode)
        ws = _WS.get(dev, nbytes)
        check(L.b200rag_dense_topk_masked(corpus16.data_ptr(), n, dim, code, queries16.data_ptr(), b, k, id_offset,
                                          scores.data_ptr(), ids.data_ptr(), flags.data_ptr(), float(row_norm_bound),
                                          out_err.data_ptr() if out_err is not None else None,
                                          row_mask.data_ptr() if row_mask is not None else None,
                                          ws.data_ptr(), ws.numel(), mode, _stream_ptr(dev)))
    return scores, ids, flags


def merge_topk(cand_scores: torch.Tensor, cand_ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[B, M] candidates (f64 score, i64 id, id<0 = empty) -> top-k by (score desc, id asc)."""
    _require_cuda(cand_scores, "cand_scores")
    _require_cuda(cand_ids, "cand_ids")
    if cand_scores.dtype != torch.float64 or cand_ids.dtype != torch.int64 or cand_scores.shape != cand_ids.shape:
        raise ValueError("merge_topk expects f64 scores and i64 ids of the same [B, M] shape")
    b, m = cand_ids.shape
    dev = cand_ids.device
    scores = torch.empty((b, k), dtype=torch.float64, device=dev)
    ids = torch.empty((b, k), dtype=torch.int64, device=dev)
    ws = _WS.get(dev, 256)
    with torch.cuda.device(dev):
        check(_lib.load().b200rag_merge_topk(cand_scores.data_ptr(), cand_ids.data_ptr(), b, m, k, scores.data_ptr(),
                                             ids.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
    return scores, ids


def merge_gathered(gathered: torch.Tensor, k: int, with_counts: bool = False):
    """All-gather buffer i64 [G, 2, B, k] (per rank a plane of fp64 score bit patterns and a plane of ids) -> global top-k
    (scores f64 [B,k], ids i64 [B,k][, valid counts i32 [B]])."""
    _require_cuda(gathered, "gathered")
    if gathered.dtype != torch.int64 or gathered.dim() != 4 or gathered.shape[1] != 2 or gathered.shape[3] != k:
        raise ValueError("merge_gathered expects an int64 [G, 2, B, k] tensor")
    g, _, b, _ = gathered.shape
    dev = gathered.device
    scores = torch.empty((b, k), dtype=torch.float64, device=dev)
    ids = torch.empty((b, k), dtype=torch.int64, device=dev)
    counts = torch.empty((b,), dtype=torch.int32, device=dev) if with_counts else None
    with torch.cuda.device(dev):
        check(_lib.load().b200rag_merge_gathered(gathered.data_ptr(), g, b, k, scores.data_ptr(), ids.data_ptr(),
                                                 counts.data_ptr() if with_counts else None, _stream_ptr(dev)))
    return (scores, ids, counts) if with_counts else (scores, ids)


@dataclass
class FusedBatch:
    ids: torch.Tensor      # i64 [B, L*K]  fused order, -1 padded
    scores: torch.Tensor   # f64 [B, L*K]
    mask: torch.Tensor     # i32 [B, L*K]  bit l: list l held the id
    first: torch.Tensor    # i32 [B, L*K]  list*K + rank0 of the payload-supplying hit
    n: torch.Tensor        # i32 [B]


def rrf_fuse(list_ids: torch.Tensor, list_len: torch.Tensor, weights: torch.Tensor, rrf_k: int = 60) -> FusedBatch:
    """list_ids i64 [L,B,K], list_len i32 [L,B], weights f64 [B,L] -> FusedBatch (reference retrieval.py:421-491)."""
    for t, nm in ((list_ids, "list_ids"), (list_len, "list_len"), (weights, "weights")):
        _require_cuda(t, nm)
    if list_ids.dtype != torch.int64 or list_len.dtype != torch.int32 or weights.dtype != torch.float64:
        raise ValueError("rrf_fuse expects i64 ids, i32 lengths, f64 weights")
    nl, b, kmax = list_ids.shape
    dev = list_ids.device
    tot = nl * kmax
    out = FusedBatch(torch.empty((b, tot), dtype=torch.int64, device=dev), torch.empty((b, tot), dtype=torch.float64, device=dev),
                     torch.empty((b, tot), dtype=torch.int32, device=dev), torch.empty((b, tot), dtype=torch.int32, device=dev),
                     torch.empty((b,), dtype=torch.int32, device=dev))
    ws = _WS.get(dev, 256)
    with torch.cuda.device(dev):
        check(_lib.load().b200rag_rrf_fuse(list_ids.data_ptr(), list_len.data_ptr(), nl, b, kmax, weights.data_ptr(), rrf_k,
                                           out.ids.data_ptr(), out.scores.data_ptr(), out.mask.data_ptr(),
                                           out.first.data_ptr(), out.n.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
    return out


def fuse_select(fused: "FusedBatch", picks: Optional[torch.Tensor], use_mmr: Optional[torch.Tensor], top_k: torch.Tensor,
                list_scores: torch.Tensor, t_max: int):
    """The tail of the fusion stage (reference retrieval.py:322-333, 485-491, 512-516): per query the MMR picks or the first
    top_k fused entries, with their columns.  fused = rrf_fuse output; picks i32 [B,t_max] / use_mmr i32 [B] or None; top_k
    i32 [B]; list_scores f64 [L,B,K].  Returns (rows i64, scores f64, mask i32, n i32 [B], first_method i32, original f64)."""
    nl, b, kmax = list_scores.shape
    dev = list_scores.device
    tot = fused.ids.shape[1]
    if tot != nl * kmax or list_scores.dtype != torch.float64 or top_k.dtype != torch.int32 or not list_scores.is_contiguous():
        raise ValueError("fuse_select: list_scores must be contiguous f64 [L,B,K] matching the fused width, top_k i32")
    for t in (picks, use_mmr):
        if t is not None and (t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous()):
            raise ValueError("fuse_select: picks / use_mmr must be contiguous CUDA i32 tensors")
    rows = torch.empty((b, t_max), dtype=torch.int64, device=dev)
    scores = torch.empty((b, t_max), dtype=torch.float64, device=dev)
    mask = torch.empty((b, t_max), dtype=torch.int32, device=dev)
    first = torch.empty((b, t_max), dtype=torch.int32, device=dev)
    orig = torch.empty((b, t_max), dtype=torch.float64, device=dev)
    n = torch.empty((b,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().b200rag_fuse_select(fused.ids.data_ptr(), fused.scores.data_ptr(), fused.mask.data_ptr(), fused.first.data_ptr(),
                                              fused.n.data_ptr(), b, tot, picks.data_ptr() if picks is not None else None,
                                              use_mmr.data_ptr() if use_mmr is not None else None, top_k.data_ptr(),
                                              list_scores.data_ptr(), nl, kmax, int(t_max), rows.data_ptr(), scores.data_ptr(),
                                              mask.data_ptr(), first.data_ptr(), orig.data_ptr(), n.data_ptr(), _stream_ptr(dev)))
    return rows, scores, mask, n, first, orig


def mmr_select(cand_doc: torch.Tensor, cand_rel: torch.Tensor, cand_n: torch.Tensor, doc_tok_ptr: torch.Tensor,
               doc_tok_ids: torch.Tensor, vocab_size: int, lam: torch.Tensor, k_sel: torch.Tensor, k_max: int
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Greedy MMR (reference retrieval.py:493-516).  Returns (picks i32 [B,k_max] positions into the candidates, n i32 [B])."""
    for t, nm in ((cand_doc, "cand_doc"), (cand_rel, "cand_rel"), (cand_n, "cand_n"), (doc_tok_ptr, "doc_tok_ptr"),
                  (doc_tok_ids, "doc_tok_ids"), (lam, "lambda"), (k_sel, "k_sel")):
        _require_cuda(t, nm)
    if (cand_doc.dtype, cand_rel.dtype, cand_n.dtype, doc_tok_ptr.dtype, doc_tok_ids.dtype, lam.dtype, k_sel.dtype) != (
            torch.int32, torch.float64, torch.int32, torch.int64, torch.int32, torch.float64, torch.int32):
        raise ValueError("mmr_select: wrong tensor dtypes")
    b, nmax = cand_doc.shape
    dev = cand_doc.device
    picks = torch.empty((b, k_max), dtype=torch.int32, device=dev)
    n = torch.empty((b,), dtype=torch.int32, device=dev)
    ws = _WS.get(dev, _lib.load().b200rag_mmr_select_workspace_bytes(b, nmax, int(vocab_size)))
    with torch.cuda.device(dev):
        check(_lib.load().b200rag_mmr_select(cand_doc.data_ptr(), cand_rel.data_ptr(), cand_n.data_ptr(), b, nmax,
                                             doc_tok_ptr.data_ptr(), doc_tok_ids.data_ptr(), int(vocab_size), lam.data_ptr(),
                                             k_sel.data_ptr(), k_max, picks.data_ptr(), n.data_ptr(), ws.data_ptr(),
                                             ws.numel(), _stream_ptr(dev)))
    return picks, n


def rerank_learned(scores: torch.Tensor, method_mask: torch.Tensor, n: torch.Tensor, k_out: int, base_weight: float = 1.0,
                   method_bonus: float = 0.1, recency_weight: float = 0.0, recency: Optional[torch.Tensor] = None
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """LearnedRanker re-rank of a fused batch (reference ranker.py:109-125 + retrieval.py:518-563): scores f64 [B,T], method
    masks i32 [B,T], valid counts i32 [B] -> (positions i32 [B,k_out] into the input, -1 padded; re-ranked scores f64; counts)."""
    for t, nm in ((scores, "scores"), (method_mask, "method_mask"), (n, "n")):
        _require_cuda(t, nm)
    if (scores.dtype, method_mask.dtype, n.dtype) != (torch.float64, torch.int32, torch.int32) or scores.shape != method_mask.shape:
        raise ValueError("rerank_learned expects f64 scores, i32 masks of the same [B, T] shape and i32 counts")
    if recency is not None and (recency.dtype != torch.float64 or recency.shape != scores.shape or not recency.is_cuda):
        raise ValueError("recency must be a CUDA f64 tensor shaped like scores")
    b, t_max = scores.shape
    dev = scores.device
    pos = torch.empty((b, k_out), dtype=torch.int32, device=dev)
    out = torch.empty((b, k_out), dtype=torch.float64, device=dev)
    cnt = torch.empty((b,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().b200rag_rerank_learned(scores.data_ptr(), method_mask.data_ptr(),
                                                 recency.contiguous().data_ptr() if recency is not None else None, n.data_ptr(), b,
                                                 t_max, float(base_weight), float(method_bonus), float(recency_weight), int(k_out),
                                                 pos.data_ptr(), out.data_ptr(), cnt.data_ptr(), _stream_ptr(dev)))
    return pos, out, cnt


def pairwise_jaccard(docs: torch.Tensor, n: torch.Tensor, doc_tok_ptr: torch.Tensor, doc_tok_ids: torch.Tensor
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Mean pairwise token-set Jaccard per result list (reference evaluation.py:327-344): docs i32 [B, n_max] rows in the token
    CSR, n i32 [B] -> (mean f64 [B], pairs averaged i32 [B]).  diversity = 1 - mean (evaluation.py:315-325)."""
    for t, nm in ((docs, "docs"), (n, "n"), (doc_tok_ptr, "doc_tok_ptr"), (doc_tok_ids, "doc_tok_ids")):
        _require_cuda(t, nm)
    if (docs.dtype, n.dtype, doc_tok_ptr.dtype, doc_tok_ids.dtype) != (torch.int32, torch.int32, torch.int64, torch.int32):
        raise ValueError("pairwise_jaccard: wrong tensor dtypes")
    b, n_max = docs.shape
    dev = docs.device
    mean = torch.empty((b,), dtype=torch.float64, device=dev)
    pairs = torch.empty((b,), dtype=torch.int32, device=dev)
    L = _lib.load()
    ws = _WS.get(dev, L.b200rag_pairwise_jaccard_workspace_bytes(b, n_max))
    with torch.cuda.device(dev):
        check(L.b200rag_pairwise_jaccard(docs.data_ptr(), n.data_ptr(), b, n_max, doc_tok_ptr.data_ptr(), doc_tok_ids.data_ptr(),
                                         mean.data_ptr(), pairs.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
    return mean, pairs


class DenseIndex:
    """A contiguous row shard of 16-bit vectors in HBM with exact inner-product / cosine search.

    metric "COSINE" (the reference's semantic/domain collections, indexing.py:143-180) stores canonically
    L2-normalised rows and normalises queries the same way; "IP" stores the plain rounded values.  Row i of this
    shard has id  id_offset + i.
    """

    def __init__(self, dim: int, dtype="f16", metric: str = "COSINE", device="cuda", id_offset: int = 0,
                 capacity: int = 0):
        if dim % 8 != 0:
            raise ValueError("dim must be a multiple of 8")
        if metric not in ("COSINE", "IP"):
            raise ValueError("metric must be COSINE or IP")
        self.dim, self.metric, self.id_offset = dim, metric, int(id_offset)
        self.code = dtype_code(dtype)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("DenseIndex lives on a CUDA device (b200rag has no CPU path)")
        self._rows = torch.empty((max(capacity, 0), dim), dtype=_TORCH_DTYPE[self.code], device=self.device)
        self.n = 0
        self.last_flags: Optional[torch.Tensor] = None
        # upper bound on the stored rows' L2 norm (error margin of the tensor-core path's completeness proof)
        self.row_norm_bound = 1.001 if metric == "COSINE" else 0.0

    @property
    def rows(self) -> torch.Tensor:
        return self._rows[: self.n]

    def _reserve(self, n_total: int) -> None:
        if n_total > self._rows.shape[0]:
            cap = max(n_total, int(self._rows.shape[0] * 1.5), 1024)
            new = torch.empty((cap, self.dim), dtype=self._rows.dtype, device=self.device)
            new[: self.n] = self._rows[: self.n]
            self._rows = new

    def add(self, rows_f32: torch.Tensor, chunk: int = 1 << 18) -> None:
        """Append fp32 rows (host or device); they are normalised / rounded on the GPU."""
        if rows_f32.dim() != 2 or rows_f32.shape[1] != self.dim:
            raise ValueError(f"rows must be [n, {self.dim}]")
        self._reserve(self.n + rows_f32.shape[0])
        for s in range(0, rows_f32.shape[0], chunk):
            part = rows_f32[s: s + chunk].to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
            prepare_rows(part, self.code, self.metric == "COSINE", out=self._rows[self.n: self.n + part.shape[0]])
            if self.metric != "COSINE" and part.shape[0]:
                nb = float(self._rows[self.n: self.n + part.shape[0]].float().norm(dim=1).max()) * 1.001
                self.row_norm_bound = max(self.row_norm_bound, nb)
            self.n += part.shape[0]

    def add_prepared(self, rows16: torch.Tensor) -> None:
        """Append rows that are already in stored form (16-bit, normalised if COSINE)."""
        if rows16.dtype != self._rows.dtype or rows16.shape[1] != self.dim:
            raise ValueError("prepared rows have the wrong dtype or dim")
        self._reserve(self.n + rows16.shape[0])
        self._rows[self.n: self.n + rows16.shape[0]] = rows16.to(self.device)
        if self.metric != "COSINE" and rows16.shape[0]:
            nb = float(self._rows[self.n: self.n + rows16.shape[0]].float().norm(dim=1).max()) * 1.001
            self.row_norm_bound = max(self.row_norm_bound, nb)
        self.n += rows16.shape[0]

    def prepare_queries(self, queries_f32: torch.Tensor) -> torch.Tensor:
        q = queries_f32.to(self.device, dtype=torch.float32, non_blocking=True)
        if q.dim() == 1:
            q = q[None, :]
        return prepare_rows(q.contiguous(), self.code, self.metric == "COSINE")

    def search(self, queries_f32: torch.Tensor, k: int, mode: int = DENSE_AUTO, out_err: Optional[torch.Tensor] = None,
               row_mask: Optional[torch.Tensor] = None, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """queries fp32 [B, dim] (host or device) -> (scores f64 [B,k], ids i64 [B,k], flags i32 [B]) on device.
        row_mask (pack_row_mask) restricts the search to the allowed rows of this shard."""
        q16 = self.prepare_queries(queries_f32)
        return self.search_prepared(q16, k, mode, out_err, row_mask, out)

    def search_prepared(self, queries16: torch.Tensor, k: int, mode: int = DENSE_AUTO, out_err: Optional[torch.Tensor] = None,
                        row_mask: Optional[torch.Tensor] = None, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        res = dense_topk(self._rows, queries16, k, self.id_offset, mode, n_rows=self.n,
                         row_norm_bound=max(self.row_norm_bound, 1e-30), out_err=out_err, row_mask=row_mask, out=out)
        self.last_flags = res[2]                 # per-query "needed the exact fallback" flags of the last search (device)
        return res


class SparseIndex:
    """Doc-range-blocked postings in HBM (layout: include/b200rag.h, b200rag_sparse_topk).

    Built from a doc-major CSR (rows = documents, columns = term ids, values = fp32 weights) -- the layout the
    reference assembles for its sparse collection (indexing.py:379-404).  The re-blocking runs on the GPU with torch
    sort/scan ops (ingest side, not the hot path).  Blocks are independent of each other, so `append` re-blocks only the
    new documents plus the last, partially filled block; earlier blocks are never touched (incremental ingest, the
    reference's per-batch `collection.insert`, indexing.py:372,419,426).
    """

    def __init__(self, doc_ptr, term_ids, weights, n_terms: int, device="cuda", block_docs: int = 16384,
                 id_offset: int = 0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("SparseIndex lives on a CUDA device (b200rag has no CPU path)")
        if block_docs % 32 or not 0 < block_docs <= 32768:
            raise ValueError("block_docs must be a multiple of 32 in (0, 32768]")
        self.n_terms = int(n_terms)
        self.block_docs = int(block_docs)
        self.id_offset = int(id_offset)
        self.n_docs = 0
        self.nnz = 0
        self._n_blocks = 0                       # block rows in use (0 while the index is empty)
        dev = self.device
        self._post_doc = torch.empty(0, dtype=torch.int16, device=dev)      # capacity buffers, valid prefix = nnz
        self._post_w = torch.empty(0, dtype=torch.float32, device=dev)
        self._ptr = torch.zeros((1, n_terms + 1), dtype=torch.int64, device=dev)   # [block capacity, V + 1]
        self.df = torch.zeros(n_terms, dtype=torch.int64, device=dev)        # document frequency per term (statistics)
        self.blocks_built = 0                    # blocks (re)built so far: observability for the incremental-ingest test
        self.append(doc_ptr, term_ids, weights)

    @property
    def n_blocks(self) -> int:
        return max(1, self._n_blocks)

    @property
    def post_doc(self) -> torch.Tensor:
        return self._post_doc[: self.nnz]

    @property
    def post_w(self) -> torch.Tensor:
        return self._post_w[: self.nnz]

    @property
    def blk_term_ptr(self) -> torch.Tensor:
        return self._ptr[: self.n_blocks]

    @staticmethod
    def _t(a, np_dtype, t_dtype):                # numpy / list / torch (any device) -> torch tensor of the wanted dtype
        return a.to(t_dtype) if torch.is_tensor(a) else torch.as_tensor(np.asarray(a, dtype=np_dtype))

    def append(self, doc_ptr, term_ids, weights) -> None:
        """Append documents (doc-major CSR over the NEW documents only); they get rows n_docs, n_docs + 1, ..."""
        dev, V, bd = self.device, self.n_terms, self.block_docs
        doc_ptr = self._t(doc_ptr, np.int64, torch.int64).cpu()
        n_new = int(doc_ptr.numel() - 1)
        if n_new <= 0:
            return
        t = self._t(term_ids, np.int64, torch.int64).to(dev)
        w = self._t(weights, np.float32, torch.float32).to(dev)
        nnz_new = int(t.numel())
        if int(doc_ptr[-1]) != nnz_new or w.numel() != nnz_new:
            raise ValueError("CSR arrays are inconsistent (doc_ptr[-1] != nnz)")
        if nnz_new and (int(t.min()) < 0 or int(t.max()) >= V):
            raise ValueError("term id out of range")
        counts = (doc_ptr[1:] - doc_ptr[:-1]).to(dev)
        d = self.n_docs + torch.repeat_interleave(torch.arange(n_new, device=dev, dtype=torch.int64), counts, output_size=nnz_new)
        first_blk = self.n_docs // bd              # first block that changes
        keep_nnz = self.nnz
        if self.n_docs % bd:
            # the last block is partially filled: pull its postings back out (term of a posting = its position in the block's
            # pointer row) and re-block them together with the new documents
            row = self._ptr[first_blk]
            keep_nnz = int(row[0])
            cnt_old = row[1:] - row[:-1]
            t_old = torch.repeat_interleave(torch.arange(V, device=dev, dtype=torch.int64), cnt_old, output_size=self.nnz - keep_nnz)
            d_old = first_blk * bd + (self._post_doc[keep_nnz: self.nnz].to(torch.int64) & 0xffff)
            d = torch.cat([d_old, d])
            t_all = torch.cat([t_old, t])
            w_all = torch.cat([self._post_w[keep_nnz: self.nnz], w])
        else:
            t_all, w_all = t, w
        n_total = self.n_docs + n_new
        n_blocks_total = -(-n_total // bd)
        nb = n_blocks_total - first_blk            # blocks (re)built by this call
        blk = d // bd - first_blk
        key = (blk * V + t_all) * bd + (d % bd)
        key, order = torch.sort(key)
        if key.numel() > 1 and bool((key[1:] == key[:-1]).any()):
            raise ValueError("duplicate (document, term) pair in the CSR")
        local = key % bd
        nnz_total = keep_nnz + int(key.numel())
        self._reserve(nnz_total, n_blocks_total)
        # u16 bit patterns held in an int16 tensor (values >= 32768 wrap; the kernel reads uint16_t)
        self._post_doc[keep_nnz: nnz_total] = (((local + 32768) % 65536) - 32768).to(torch.int16)
        self._post_w[keep_nnz: nnz_total] = w_all[order]
        seg = key // bd                            # blk * V + term
        cnt = torch.bincount(seg, minlength=nb * V) if key.numel() else torch.zeros(nb * V, dtype=torch.int64, device=dev)
        start = keep_nnz + torch.cumsum(cnt, 0) - cnt
        self._ptr[first_blk: n_blocks_total, :V] = start.view(nb, V)
        self._ptr[first_blk: n_blocks_total, V] = keep_nnz + torch.cumsum(cnt.view(nb, V).sum(1), 0)
        if nnz_new:
            self.df += torch.bincount(t, minlength=V)
        self.n_docs, self.nnz, self._n_blocks = n_total, nnz_total, n_blocks_total
        self.blocks_built += nb

    def to_doc_major(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """The index as a doc-major CSR again (device tensors: doc_ptr i64 [n_docs+1], term ids i64 ascending per document,
        weights f32) -- checkpointing and compaction read the postings back instead of keeping a host copy of them."""
        dev, V, bd = self.device, self.n_terms, self.block_docs
        ptr = self.blk_term_ptr
        cnt = (ptr[:, 1:] - ptr[:, :-1]).reshape(-1)                              # postings per (block, term)
        seg = torch.repeat_interleave(torch.arange(cnt.numel(), device=dev, dtype=torch.int64), cnt, output_size=self.nnz)
        d = (seg // V) * bd + (self.post_doc.to(torch.int64) & 0xffff)
        key, order = torch.sort(d * V + seg % V)
        doc_ptr = torch.zeros(self.n_docs + 1, dtype=torch.int64, device=dev)
        if self.nnz:
            doc_ptr[1:] = torch.cumsum(torch.bincount(key // V, minlength=self.n_docs), 0)
        return doc_ptr, key % V, self.post_w[order]

    def _reserve(self, nnz: int, n_blocks: int) -> None:
        if nnz > self._post_doc.numel():
            cap = max(nnz, int(self._post_doc.numel() * 1.5))
            for name in ("_post_doc", "_post_w"):
                old = getattr(self, name)
                new = torch.empty(cap, dtype=old.dtype, device=self.device)
                new[: self.nnz] = old[: self.nnz]
                setattr(self, name, new)
        if n_blocks > self._ptr.shape[0]:
            cap = max(n_blocks, int(self._ptr.shape[0] * 1.5))
            new = torch.zeros((cap, self.n_terms + 1), dtype=torch.int64, device=self.device)
            new[: self._n_blocks] = self._ptr[: self._n_blocks]
            self._ptr = new

    def search(self, q_ptr, q_terms, q_vals, k: int, doc_mask: Optional[torch.Tensor] = None):
        """CSR queries (q_ptr i64 [B+1], q_terms i32 ascending per query, q_vals f32) ->
        (scores f32 [B,k], ids i64 [B,k], counts i32 [B]) on device.  doc_mask (pack_row_mask) = metadata filter."""
        dev = self.device
        q_ptr = torch.as_tensor(q_ptr, dtype=torch.int64).to(dev).contiguous()
        q_terms = torch.as_tensor(q_terms, dtype=torch.int32).to(dev).contiguous()
        q_vals = torch.as_tensor(q_vals, dtype=torch.float32).to(dev).contiguous()
        b = q_ptr.numel() - 1
        scores = torch.empty((b, k), dtype=torch.float32, device=dev)
        ids = torch.empty((b, k), dtype=torch.int64, device=dev)
        counts = torch.zeros((b,), dtype=torch.int32, device=dev)
        L = _lib.load()
        with torch.cuda.device(dev):
            nbytes = L.b200rag_sparse_topk_workspace_bytes(self.n_docs, self.block_docs, b, k)
            ws = _WS.get(dev, nbytes)
            check(L.b200rag_sparse_topk_masked(self.blk_term_ptr.data_ptr(), self._post_doc.data_ptr(), self._post_w.data_ptr(),
                                               self.n_docs, self.n_terms, self.block_docs, q_ptr.data_ptr(), q_terms.data_ptr(),
                                               q_vals.data_ptr(), b, k, self.id_offset, scores.data_ptr(), ids.data_ptr(),
                                               counts.data_ptr(), doc_mask.data_ptr() if doc_mask is not None else None,
                                               ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
        return scores, ids, counts

    def query_bytes(self, q_ptr, q_terms) -> int:
        """Algorithmic HBM bytes of a query batch: sum over query terms of df(t) * 6 (u16 doc + f32 weight)."""
        qt = torch.as_tensor(q_terms, dtype=torch.int64).to(self.device)
        return int(self.df[qt].sum().item()) * 6
