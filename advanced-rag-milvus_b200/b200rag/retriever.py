"""B200HybridRetriever -- the reference's Hybrid Retrieval Engine with its arithmetic moved onto the GPU.

Same public surface as reference src/advanced_rag/retrieval.py `HybridRetriever` (:104-648): constructor arguments,
`retrieve` (with the 0.3 s budget, :215-247), `_retrieve_inner` (:249-339), `_search_semantic/_sparse/_domain`
(:341-419), `_fuse_results` (:421-491), `_mmr_diversify` (:493-516), `rerank` (:518-563) and
`_build_filter_expression` (:565-632), so existing callers (AdvancedRAGPipeline.retrieve, pipeline.py:217-309) and
the reference's own tests run unchanged against it.  What differs is where the work happens:

    _fuse_results   weighted Reciprocal Rank Fusion   -> rrf_fuse kernel   (fp64, reference operation order)
    _mmr_diversify  greedy MMR on token-set Jaccard   -> mmr_select kernel (fp64, strict '>' tie rule)
    retrieve_batch  NEW: a whole batch of queries, per-query profile (top_k, weights, MMR flag, lambda), through dense
                    scan + sparse scan + RRF + MMR with every intermediate staying on the device; only the final
                    <= top_k rows per query are turned into result dicts on the host.

The index manager is duck-typed exactly as in the reference: anything with `search` and `_generate_*_embedding`
works for the one-query path; `retrieve_batch` needs a B200IndexManager.
"""
from __future__ import annotations

import asyncio
import logging
import re
from dataclasses import dataclass
from datetime import datetime
from typing import Any, Callable, Dict, List, Optional, Sequence, Set, Tuple

import numpy as np
import torch

from . import engine
from .config import (DOMAIN_WEIGHT, RRF_K, TIMEOUT_SECONDS, LearnedRanker, QueryClassifier, RetrievalConfig,
                     build_default_profiles)

logger = logging.getLogger(__name__)
METHODS = ("semantic", "sparse", "domain")
_OPS = {"$gte": ">=", "$lte": "<=", "$gt": ">", "$lt": "<", "$eq": "==", "$ne": "!="}


def _device_of(manager) -> torch.device:
    dev = getattr(manager, "device", None)
    return torch.device(dev) if dev is not None else torch.device("cuda")


@dataclass
class BatchResult:
    """Device-resident result of a batched retrieve; row r of query b is valid iff r < n[b]."""
    rows: torch.Tensor            # i64 [B,T]  corpus row ids in final order (-1 pads)
    scores: torch.Tensor          # f64 [B,T]  fused scores
    mask: torch.Tensor            # i32 [B,T]  bit l: retrieval method l (semantic, sparse, domain) returned the row
    n: torch.Tensor               # i32 [B]
    first_method: torch.Tensor    # i32 [B,T]  the method whose hit supplies the payload (reference retrieval.py:441-461)
    original_score: torch.Tensor  # f64 [B,T]  that method's own score for the row


class B200HybridRetriever:
    ALLOWED_FILTER_FIELDS: Set[str] = {"doc_id", "chunk_id", "domain_density", "timestamp", "entropy", "redundancy",
                                       "chunk_index", "token_count"}
    ALLOWED_OPERATORS: Set[str] = set(_OPS)

    def __init__(self, index_manager, config: Optional[RetrievalConfig] = None,
                 weight_adapter: Optional[Callable[[str], Tuple[float, float]]] = None,
                 classifier: Optional[QueryClassifier] = None, profiles: Optional[Dict[str, RetrievalConfig]] = None,
                 learned_ranker: Optional[LearnedRanker] = None):
        self.index_manager = index_manager
        base = config or RetrievalConfig()
        self.config = base
        self.weight_adapter = weight_adapter
        self.classifier = classifier or QueryClassifier()
        self.profiles: Dict[str, RetrievalConfig] = profiles or build_default_profiles(base)
        self.reranker = None
        self.learned_ranker = learned_ranker
        self.device = _device_of(index_manager)

    _build_default_profiles = staticmethod(build_default_profiles)

    # ------------------------------------------------------------------------------------------- one query
    async def retrieve(self, query: str, filters: Optional[Dict[str, Any]] = None, use_domain_index: bool = False,
                       domain: Optional[str] = None, profile_hint: Optional[str] = None) -> List[Dict[str, Any]]:
        try:
            return await asyncio.wait_for(self._retrieve_inner(query, filters, use_domain_index, domain, profile_hint),
                                          timeout=float(TIMEOUT_SECONDS))
        except asyncio.TimeoutError:
            logger.warning("B200HybridRetriever.retrieve timed out after %.3f seconds", float(TIMEOUT_SECONDS))
            return []           # degrade to "no results", never raise (reference retrieval.py:242-247)

    def _pick_profile(self, query: str, profile_hint: Optional[str]) -> str:
        try:
            if profile_hint and profile_hint in self.profiles:
                return profile_hint
            if self.classifier:
                return self.classifier.classify(query) or "default"
        except Exception:  # noqa: BLE001 - any classifier failure means "default" (reference :278-279)
            pass
        return "default"

    def _adapt_weights(self, query: str, cfg: RetrievalConfig) -> None:
        """weight_adapter hook with the reference's clamp-to-[0,1] and keep-if-positive-sum rule (:309-320)."""
        if not self.weight_adapter:
            return
        try:
            dw, sw = self.weight_adapter(query)
            dw, sw = max(0.0, min(1.0, float(dw))), max(0.0, min(1.0, float(sw)))
            if dw + sw > 0:
                cfg.dense_weight, cfg.sparse_weight = dw, sw
        except Exception:  # noqa: BLE001
            pass

    async def _retrieve_inner(self, query: str, filters: Optional[Dict[str, Any]] = None, use_domain_index: bool = False,
                              domain: Optional[str] = None, profile_hint: Optional[str] = None) -> List[Dict[str, Any]]:
        profile_name = self._pick_profile(query, profile_hint)
        # per-request config.  The reference assigns it to self.config (:281-284) and reads it back after every await, which
        # races when requests overlap on one retriever; here the request keeps its own reference and
        # self.config is only updated for callers that inspect it afterwards.
        cfg = self.profiles.get(profile_name, self.config)
        self.config = cfg
        semantic_emb = await self._get_semantic_embedding(query)
        sparse_emb = await self._get_sparse_embedding(query)
        filter_expr = self._build_filter_expression(filters) if filters else None
        tasks = [self._search_semantic(semantic_emb, filter_expr, cfg), self._search_sparse(sparse_emb, filter_expr, cfg)]
        if use_domain_index and domain:
            tasks.append(self._search_domain(await self._get_domain_embedding(query, domain), filter_expr, cfg))
        lists = await asyncio.gather(*tasks)
        self._adapt_weights(query, cfg)
        fused = self._fuse_results(semantic_results=lists[0], sparse_results=lists[1],
                                   domain_results=lists[2] if len(lists) > 2 else [], cfg=cfg)
        for r in fused:
            meta = r.get("metadata")
            if isinstance(meta, dict):
                meta.setdefault("retrieval_profile", profile_name)
            else:
                r["retrieval_profile"] = profile_name
        return fused[: cfg.top_k]

    async def _tagged_search(self, embedding, collection: str, top_k: int, filters, params, method: str):
        try:
            hits = await self.index_manager.search(query_embedding=embedding, collection_name=collection, top_k=top_k,
                                                   filters=filters, search_params=params)
        except Exception:  # noqa: BLE001 - a failing index degrades to "no hits" (reference :355-358,387-389,411-413)
            return []
        for h in hits:
            h["method"] = method
            h["original_score"] = h["score"]
        return hits

    async def _search_semantic(self, embedding, filters: Optional[str], cfg: Optional[RetrievalConfig] = None) -> List[Dict[str, Any]]:
        cfg = cfg or self.config
        return await self._tagged_search(embedding, "semantic_index", cfg.top_k * 2, filters, cfg.semantic_search_params, "semantic")

    async def _search_sparse(self, embedding, filters: Optional[str], cfg: Optional[RetrievalConfig] = None) -> List[Dict[str, Any]]:
        cfg = cfg or self.config
        collections = getattr(self.index_manager, "collections", None)
        if collections is not None and "sparse_index" not in collections:
            return []
        return await self._tagged_search(embedding, "sparse_index", cfg.top_k * 2, filters, cfg.sparse_search_params, "sparse")

    async def _search_domain(self, embedding, filters: Optional[str], cfg: Optional[RetrievalConfig] = None) -> List[Dict[str, Any]]:
        cfg = cfg or self.config
        return await self._tagged_search(embedding, "domain_index", cfg.top_k, filters, cfg.semantic_search_params, "domain")

    # ------------------------------------------------------------------------------------------- fusion (GPU)
    def _fuse_results(self, semantic_results: List[Dict], sparse_results: List[Dict],
                      domain_results: Optional[List[Dict]] = None, cfg: Optional[RetrievalConfig] = None) -> List[Dict[str, Any]]:
        """Weighted RRF of up to three ranked hit lists on the GPU; result dicts as the reference builds them."""
        cfg = cfg or self.config
        lists = [semantic_results or [], sparse_results or []] + ([domain_results] if domain_results else [])
        weights = [cfg.dense_weight, cfg.sparse_weight, DOMAIN_WEIGHT][: len(lists)]
        code: Dict[Any, int] = {}
        for lst in lists:
            for h in lst:
                code.setdefault(h["id"], len(code))
        if not code:
            return []
        kmax = max(1, max(len(lst) for lst in lists))
        ids = np.full((len(lists), 1, kmax), -1, dtype=np.int64)
        for li, lst in enumerate(lists):
            ids[li, 0, : len(lst)] = [code[h["id"]] for h in lst]
        lens = np.asarray([[len(lst)] for lst in lists], dtype=np.int32)
        dev = self.device
        fused = engine.rrf_fuse(torch.as_tensor(ids).to(dev), torch.as_tensor(lens).to(dev),
                                torch.tensor([weights], dtype=torch.float64, device=dev), RRF_K)
        n = int(fused.n[0])
        order = fused.ids[0, :n].cpu().tolist()
        scores = fused.scores[0, :n].cpu().tolist()
        masks = fused.mask[0, :n].cpu().tolist()
        # payload: the semantic hit if the id appears there (the last one on duplicates, :441), else the first list
        # that saw it (:449-450, :460-461)
        data: Dict[int, Dict] = {}
        for li, lst in enumerate(lists):
            for h in lst:
                c = code[h["id"]]
                if li == 0 or c not in data:
                    data[c] = h
        now = datetime.utcnow()
        out = []
        for c, sc, mk in zip(order, scores, masks):
            r = data[c]
            r["score"] = sc
            r["retrieval_methods"] = [m for bit, m in enumerate(METHODS) if mk >> bit & 1]
            meta = r.get("metadata")
            if isinstance(meta, dict) and "timestamp" in meta and "recency" not in meta:
                try:
                    age_days = max(0.0, (now - datetime.fromisoformat(str(meta["timestamp"]))).total_seconds() / 86400.0)
                    meta["recency"] = float(1.0 / (1.0 + age_days))
                except Exception:  # noqa: BLE001 - unparsable timestamps are skipped (:481-483)
                    pass
            out.append(r)
        if cfg.enable_mmr and out:
            return self._mmr_diversify(out, cfg.top_k, cfg.mmr_lambda)
        return out

    def _mmr_diversify(self, ranked: List[Dict[str, Any]], k: int, mmr_lambda: float) -> List[Dict[str, Any]]:
        """Greedy MMR over token-set Jaccard on the GPU; `ranked` is in fused order, returns min(k, n) picks."""
        n = len(ranked)
        if n == 0 or k <= 0:
            return []
        vocab: Dict[str, int] = {}
        ptr, toks = [0], []
        for r in ranked:
            s = sorted({vocab.setdefault(t, len(vocab)) for t in (r.get("content") or "").lower().split()})
            toks.extend(s)
            ptr.append(len(toks))
        dev = self.device
        k_sel = min(int(k), n)
        picks, cnt = engine.mmr_select(
            torch.arange(n, dtype=torch.int32, device=dev)[None, :].contiguous(),
            torch.tensor([[float(r["score"]) for r in ranked]], dtype=torch.float64, device=dev),
            torch.tensor([n], dtype=torch.int32, device=dev),
            torch.tensor(ptr, dtype=torch.int64, device=dev),
            torch.tensor(toks if toks else [0], dtype=torch.int32, device=dev), max(1, len(vocab)),
            torch.tensor([float(mmr_lambda)], dtype=torch.float64, device=dev),
            torch.tensor([k_sel], dtype=torch.int32, device=dev), k_sel)
        return [ranked[i] for i in picks[0, : int(cnt[0])].cpu().tolist()]

    # ------------------------------------------------------------------------------------------- rerank (host, tiny)
    async def rerank(self, query: str, results: List[Dict[str, Any]], top_k: Optional[int] = None) -> List[Dict[str, Any]]:
        if not self.config.enable_reranking or not results:
            return results[:top_k] if top_k else results
        top_k = top_k or self.config.rerank_top_k
        if self.learned_ranker and self.config.enable_learned_ranker:
            new = await self.learned_ranker.score(query, results)
        elif self.reranker:
            new = await self.reranker.score([(query, r["content"]) for r in results])
        else:
            new = [r["score"] + np.random.normal(0, 0.01) for r in results]      # the reference's placeholder (:550-553)
        for r, s in zip(results, new):
            r["rerank_score"] = s
            r["original_retrieval_score"] = r["score"]
            r["score"] = s
        results.sort(key=lambda r: r["rerank_score"], reverse=True)
        return results[:top_k]

    # ------------------------------------------------------------------------------------------- filters
    def _build_filter_expression(self, filters: Dict[str, Any]) -> Optional[str]:
        """Whitelisted metadata predicate -> boolean expression string (same grammar and escaping as the reference)."""
        def quote(v: str) -> str:
            return '"' + v.replace("\\", "\\\\").replace('"', '\\"') + '"'

        terms = []
        for field, value in filters.items():
            if field not in self.ALLOWED_FILTER_FIELDS:
                logger.warning("Invalid filter field attempted: %s", field)
                raise ValueError(f"Invalid filter field: {field}")
            if not re.match(r"^[a-zA-Z_][a-zA-Z0-9_]*$", field):
                raise ValueError(f"Invalid field name format: {field}")
            if isinstance(value, dict):
                for op, v in value.items():
                    if op not in self.ALLOWED_OPERATORS:
                        logger.warning("Invalid operator attempted: %s", op)
                        raise ValueError(f"Invalid operator: {op}")
                    if not isinstance(v, (int, float, str, bool)):
                        raise ValueError(f"Invalid value type for {field}: {type(v)}")
                    terms.append(f"{field} {_OPS[op]} {quote(v) if isinstance(v, str) else v}")
            elif isinstance(value, str):
                terms.append(f"{field} == {quote(value)}")
            elif isinstance(value, (int, float, bool)):
                terms.append(f"{field} == {value}")
            else:
                raise ValueError(f"Unsupported value type for {field}: {type(value)}")
        return " and ".join(terms) if terms else None

    async def _get_semantic_embedding(self, text: str):
        return await self.index_manager._generate_semantic_embedding(text)

    async def _get_sparse_embedding(self, text: str):
        return await self.index_manager._generate_sparse_embedding(text)

    async def _get_domain_embedding(self, text: str, domain: str):
        return await self.index_manager._generate_domain_embedding(text, domain)

    # ------------------------------------------------------------------------------------------- batch (device resident)
    def fuse_batch(self, lists: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                   configs: Sequence[RetrievalConfig]) -> "BatchResult":
        """RRF (+ MMR where a query's profile enables it) for a batch, entirely on the device.

        lists: per retrieval method (semantic, sparse[, domain]) a triple (scores f64 [B,K_l], row ids i64 [B,K_l]
        with -1 padding, valid counts i32 [B]); configs: one RetrievalConfig per query.

        On a row-sharded manager every rank holds the same merged candidate lists; the fusion work is then split BY QUERY:
        rank r fuses a contiguous slice of the batch and one all-gather hands every rank the whole result."""
        b = len(configs)
        t_max = max(c.top_k for c in configs)
        m = self.index_manager
        world, rank = getattr(m, "world", 1), getattr(m, "rank", 0)
        if world <= 1 or b < 2 * world:
            return self._fuse_local(lists, configs, t_max)
        import torch.distributed as dist
        per = -(-b // world)
        lo, hi = min(b, rank * per), min(b, (rank + 1) * per)
        dev = self.device
        msg = torch.zeros((per, 4 * t_max + 1), dtype=torch.int64, device=dev)
        if hi > lo:
            r = self._fuse_local([(s[lo:hi], i[lo:hi], c[lo:hi]) for s, i, c in lists], configs[lo:hi], t_max)
            n = hi - lo
            msg[:n, :t_max] = r.rows
            msg[:n, t_max: 2 * t_max] = r.scores.view(torch.int64)
            msg[:n, 2 * t_max: 3 * t_max] = r.original_score.view(torch.int64)
            msg[:n, 3 * t_max: 4 * t_max] = (r.mask.to(torch.int64) << 32) | r.first_method.to(torch.int64)
            msg[:n, 4 * t_max] = r.n.to(torch.int64)
        out = torch.empty((world * per, 4 * t_max + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(out, msg, group=getattr(m, "group", None))
        out = out[:b]
        packed = out[:, 3 * t_max: 4 * t_max]
        return BatchResult(out[:, :t_max].contiguous(), out[:, t_max: 2 * t_max].contiguous().view(torch.float64),
                           (packed >> 32).to(torch.int32), out[:, 4 * t_max].to(torch.int32),
                           (packed & 0xffffffff).to(torch.int32), out[:, 2 * t_max: 3 * t_max].contiguous().view(torch.float64))

    def _config_columns(self, configs: Sequence[RetrievalConfig], n_lists: int):
        """Per-query fusion parameters as device columns: weights f64 [B,L], top_k i32 [B], lambda f64 [B], use_mmr i32 [B].
        One packed host array -> one copy; batches that repeat a configuration pattern (the usual case: one profile for the
        whole batch) reuse the columns of the last call."""
        key = (n_lists, tuple((c.dense_weight, c.sparse_weight, c.top_k, c.mmr_lambda, c.enable_mmr) for c in configs))
        hit = self._cfg_cache.get(key) if hasattr(self, "_cfg_cache") else None
        if hit is None:
            if not hasattr(self, "_cfg_cache") or len(self._cfg_cache) > 16:
                self._cfg_cache = {}
            host = np.asarray([[c.dense_weight, c.sparse_weight, DOMAIN_WEIGHT, c.top_k, c.mmr_lambda, float(c.enable_mmr)]
                               for c in configs], dtype=np.float64)
            cols = torch.from_numpy(host).to(self.device)
            hit = (cols[:, :n_lists].contiguous(), cols[:, 3].to(torch.int32), cols[:, 4].contiguous(),
                   cols[:, 5].to(torch.int32), bool(host[:, 5].any()))
            self._cfg_cache[key] = hit
        return hit

    def _fuse_local(self, lists: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], configs: Sequence[RetrievalConfig],
                    t_max: int) -> "BatchResult":
        dev = self.device
        b = len(configs)
        kmax = max(int(ids.shape[1]) for _, ids, _ in lists)
        if all(int(ids.shape[1]) == kmax for _, ids, _ in lists):
            lid = torch.stack([ids for _, ids, _ in lists])
            lsc = torch.stack([sc for sc, _, _ in lists])
            llen = torch.stack([cnt for _, _, cnt in lists])
        else:
            lid = torch.full((len(lists), b, kmax), -1, dtype=torch.int64, device=dev)
            lsc = torch.zeros((len(lists), b, kmax), dtype=torch.float64, device=dev)
            llen = torch.zeros((len(lists), b), dtype=torch.int32, device=dev)
            for li, (sc, ids, cnt) in enumerate(lists):
                lid[li, :, : ids.shape[1]] = ids
                lsc[li, :, : ids.shape[1]] = sc
                llen[li] = cnt
        w, top, lam, use, any_mmr = self._config_columns(configs, len(lists))
        fused = engine.rrf_fuse(lid, llen, w, RRF_K)
        picks = None
        if any_mmr:
            tok_ptr, tok_ids, vocab = self.index_manager.token_sets()
            k_sel = top * use                                                  # k_sel = 0: the kernel skips the query
            picks, _ = engine.mmr_select(fused.ids.clamp(min=0).to(torch.int32), fused.scores, fused.n,
                                         tok_ptr, tok_ids, vocab, lam, k_sel, t_max)
        rows, scores, mask, n_out, first, orig = engine.fuse_select(fused, picks, use if any_mmr else None, top, lsc, t_max)
        return BatchResult(rows, scores, mask, n_out, first, orig)

    def rerank_batch(self, res: "BatchResult", top_k: Optional[int] = None,
                     recency: Optional[torch.Tensor] = None) -> "BatchResult":
        """`rerank` (reference retrieval.py:518-563) with the LearnedRanker for a whole device-resident batch: one launch of
        b200rag_rerank_learned scores  base_weight*score + method_bonus*|retrieval_methods| + recency_weight*recency  in the
        reference's fp64 operation order and stable-sorts every list; the columns of `res` are re-ordered on the device.
        recency: f64 [B,T] (1 / (1 + age in days), retrieval.py:473-483) or None.  `original_score` keeps what the reference
        calls original_retrieval_score (the fused score before the re-rank)."""
        if not self.config.enable_reranking:
            return res
        if not (self.learned_ranker and self.config.enable_learned_ranker):
            raise ValueError("rerank_batch runs the LearnedRanker; cross-encoder re-ranking goes through rerank()")
        k_out = int(top_k or self.config.rerank_top_k)
        c = self.learned_ranker.config
        pos, new_scores, cnt = engine.rerank_learned(res.scores.contiguous(), res.mask.contiguous(), res.n.contiguous(), k_out,
                                                     c.base_weight, c.method_bonus, c.recency_weight, recency)
        take = pos.to(torch.int64).clamp(min=0)
        valid = pos >= 0
        rows = torch.where(valid, res.rows.gather(1, take), torch.full_like(take, -1))
        mask = torch.where(valid, res.mask.gather(1, take), torch.zeros_like(pos))
        first = torch.where(valid, res.first_method.gather(1, take), torch.zeros_like(pos))
        orig = torch.where(valid, res.scores.gather(1, take), torch.full_like(new_scores, float("-inf")))
        return BatchResult(rows, new_scores, mask, cnt, first, orig)

    def retrieve_batch_embedded(self, semantic: Any, sparse: Sequence[Any], configs: Sequence[RetrievalConfig],
                                domain: Any = None, filter_expr: Optional[str] = None) -> "BatchResult":
        """The hot path for pre-embedded queries: dense top-2k + sparse top-2k (+ domain top-k) -> RRF -> MMR.
        Everything stays on the device."""
        m = self.index_manager
        k2 = 2 * max(c.top_k for c in configs)
        twice = torch.tensor([2 * c.top_k for c in configs], dtype=torch.int32, device=self.device)
        ss, si, sc = m.search_batch_ids(semantic, "semantic_index", k2, filter_expr)
        lists = [(ss, si, torch.minimum(sc, twice))]
        if "sparse_index" in m.collections:
            ps, pi, pc = m.search_batch_ids(sparse, "sparse_index", k2, filter_expr)
            lists.append((ps, pi, torch.minimum(pc, twice)))
        else:
            lists.append((torch.zeros((len(configs), 1), dtype=torch.float64, device=self.device),
                          torch.full((len(configs), 1), -1, dtype=torch.int64, device=self.device),
                          torch.zeros(len(configs), dtype=torch.int32, device=self.device)))
        if domain is not None:
            once = torch.tensor([c.top_k for c in configs], dtype=torch.int32, device=self.device)
            ds, di, dc = m.search_batch_ids(domain, "domain_index", k2 // 2, filter_expr)
            lists.append((ds, di, torch.minimum(dc, once)))
        return self.fuse_batch(lists, configs)

    async def retrieve_batch(self, queries: Sequence[str], filters: Optional[Dict[str, Any]] = None,
                             use_domain_index: bool = False, domain: Optional[str] = None,
                             profile_hints: Optional[Sequence[Optional[str]]] = None) -> List[List[Dict[str, Any]]]:
        """`retrieve` for many queries at once: one result list per query, identical to calling `retrieve` on each."""
        import copy
        m = self.index_manager
        names = [self._pick_profile(q, profile_hints[i] if profile_hints else None) for i, q in enumerate(queries)]
        configs = []
        for q, nm in zip(queries, names):
            cfg = copy.copy(self.profiles.get(nm, self.config))
            self._adapt_weights(q, cfg)
            configs.append(cfg)
        sem = np.stack([np.asarray(await self._get_semantic_embedding(q), dtype=np.float32) for q in queries])
        spa = [await self._get_sparse_embedding(q) for q in queries]
        dom = None
        if use_domain_index and domain:
            dom = np.stack([np.asarray(await self._get_domain_embedding(q, domain), dtype=np.float32) for q in queries])
        expr = self._build_filter_expression(filters) if filters else None
        res_dev = self.retrieve_batch_embedded(sem, spa, configs, dom, expr)
        rows_h, sc_h, mk_h, n_h = (t.cpu().numpy() for t in (res_dev.rows, res_dev.scores, res_dev.mask, res_dev.n))
        first_h, orig_h = res_dev.first_method.cpu().numpy(), res_dev.original_score.cpu().numpy()
        # result dicts for the whole batch from one gather per payload column (indexing.py:534-551 shape), then the tags the
        # reference adds on the way (retrieval.py:361-363, 469-470, 473-483, 332)
        valid = np.arange(rows_h.shape[1])[None, :] < n_h[:, None]
        flat = m.payload.hits(rows_h[valid], sc_h[valid])
        first_f, orig_f, mk_f = first_h[valid].tolist(), orig_h[valid].tolist(), mk_h[valid].tolist()
        now = datetime.utcnow()
        out, pos = [], 0
        for b, nm in enumerate(names):
            res = flat[pos: pos + int(n_h[b])]
            for j, hit in enumerate(res, start=pos):
                hit["method"] = METHODS[first_f[j]]
                hit["original_score"] = orig_f[j]
                mk = mk_f[j]
                hit["retrieval_methods"] = [mm for bit, mm in enumerate(METHODS) if mk >> bit & 1]
                meta = hit["metadata"]
                if meta.get("timestamp") is not None:
                    try:
                        age = max(0.0, (now - datetime.fromisoformat(str(meta["timestamp"]))).total_seconds() / 86400.0)
                        meta["recency"] = float(1.0 / (1.0 + age))
                    except Exception:  # noqa: BLE001
                        pass
                meta.setdefault("retrieval_profile", nm)
            pos += int(n_h[b])
            out.append(res)
        return out
