"""Secondary configurations of BASELINE.json, measured inside bench.py's run so that they land in the driver's record
(they used to exist only as builder logs).  Everything goes through the product's own call path -- the index
manager / retriever of advanced-rag-milvus_b200/b200rag -- with CUDA-event medians in steady state.

    c1        100K chunks x 384-d + BM25 (30K terms), alpha 0.7, top_k 20: dense top-40 + sparse top-40 -> RRF -> top 20, batch 256,
              through B200IndexManager (real payload) + B200HybridRetriever; next to it the reference chain on the host cores
              (oracle/pipeline.py restatement of HybridRetriever.retrieve over the "Milvus mocked" arrays), >= 200
              queries one at a time: p50 latency + QPS, and the two are compared result by result.
    c2        1M x 768 fp16 inner product top-100, batch 1024.
    c4        1M x 1024 bf16 dense + BM25 (1M docs, 100K terms) -> RRF -> MMR 0.7 over <= 1000 candidates, k = 100, batch 256:
              per-stage times, the sparse scan against its HBM roofline (bytes = sum over query terms of df * 6).
    c5_shard  one GPU's 12.5M x 384 fp16 shard of config 5, top-10, batch 4096.
    near_duplicates  1M x 768 corpus in which every row occurs 64 times: cost of the exact-fallback tiers.
Only bench.py imports this module (it uses oracle/ for the CPU leg of c1, which nothing but tests and bench.py may do).
"""
from __future__ import annotations

import os
import statistics
import time

import numpy as np
import torch

HBM_GBS = 6545.0


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    r = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), r


def wall(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    r = None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), r


def c1(dev, reps=10, n_cpu=200):
    from b200rag import bm25, synth
    from b200rag.config import RetrievalConfig
    from b200rag.index_manager import B200IndexManager
    from b200rag.retriever import B200HybridRetriever
    from oracle import pipeline as opipe
    n, d, v, b, tk = 100_000, 384, 30_000, 256, 20
    x = synth.dense_rows(n, d, 0)
    dp, ti, tf = synth.zipf_corpus(n, v, 0)
    w = bm25.bm25_weights(dp, ti, tf, v)
    contents = [" ".join([f"w{t}" for t in ti[dp[r]: dp[r + 1]]]) for r in range(n)]
    t0 = time.time()
    mgr = B200IndexManager(semantic_dim=d, sparse_dim=v, domain_dim=8, device=dev, enable_sparse=True)
    mgr.add([f"c{r:09d}" for r in range(n)], contents, x, (dp, ti, w), None, None)
    torch.cuda.synchronize()
    t_ingest = time.time() - t0
    retr = B200HybridRetriever(mgr, RetrievalConfig(hybrid_alpha=0.7, top_k=tk))
    q = synth.dense_rows(b, d, 1000)
    qp, qt, qv = synth.zipf_queries(b, v, 1, n_terms=8, skip_top=100)
    cfgs = [RetrievalConfig(hybrid_alpha=0.7, top_k=tk)] * b
    q_pin = torch.from_numpy(q).pin_memory()

    def columnar():
        r = retr.retrieve_batch_embedded(q_pin, (qp, qt, qv), cfgs)
        return r.rows.cpu(), r.scores.cpu(), r.n.cpu()

    def with_dicts():
        rows, scores, nn = columnar()
        valid = torch.arange(rows.shape[1])[None, :] < nn[:, None]
        return mgr.payload.hits(rows[valid].numpy(), scores[valid].numpy())

    t_col, (rows, scores, nn) = wall(columnar, reps)
    t_dict, hits = wall(with_dicts, max(3, reps // 3))
    # the reference chain on the host cores, one query at a time (what a CPU deployment of the mocked-Milvus path does)
    corpus = opipe.ArrayCorpus(x, None, dp, ti, w, v, None)
    lat, same = [], True
    for r in range(n_cpu):
        sq = {"indices": qt[qp[r]: qp[r + 1]].tolist(), "values": qv[qp[r]: qp[r + 1]].tolist()}
        t1 = time.perf_counter()
        ids, scs, _ = opipe.retrieve(corpus, q[r], sq, None, tk)
        lat.append((time.perf_counter() - t1) * 1e3)
        m = int(nn[r])
        same &= rows[r, :m].tolist() == ids and scores[r, :m].tolist() == scs
    out = {"workload": f"{n} chunks x {d}-d + BM25 over {v} terms, alpha 0.7, top_k {tk}, RRF; batch {b} through B200IndexManager + B200HybridRetriever",
           "gpu_ms_per_batch_columnar": t_col, "gpu_qps_columnar": b / t_col * 1e3,
           "gpu_ms_per_batch_with_result_dicts": t_dict, "gpu_qps_with_result_dicts": b / t_dict * 1e3, "result_dicts_per_batch": len(hits),
           "ingest_s": t_ingest,
           "cpu_reference_chain": {"kind": "port", "what": "oracle/pipeline.py restatement of HybridRetriever.retrieve over in-memory arrays "
                                                           "(exact C/OpenMP scans), one query per call", "queries": n_cpu,
                                   "p50_ms": statistics.median(lat), "qps": 1e3 / statistics.mean(lat), "cores": os.cpu_count()},
           "bit_exact_vs_cpu_chain": bool(same), "checked_queries": n_cpu}
    del mgr, retr
    torch.cuda.empty_cache()
    return out


def dense_config(dev, n, d, b, k, metric, reps=8):
    from b200rag import engine
    g = torch.Generator(device=dev).manual_seed(1)
    idx = engine.DenseIndex(d, "f16", metric, dev, capacity=n)
    for s in range(0, n, 250_000):
        x = torch.randn(min(250_000, n - s), d, generator=g, device=dev)
        idx.add(x / x.norm(dim=1, keepdim=True) if metric == "IP" else x)
    q = [torch.randn(b, d, generator=g, device=dev) for _ in range(4)]
    if metric == "IP":
        q = [t / t.norm(dim=1, keepdim=True) for t in q]
    it = [0]

    def fn():
        it[0] += 1
        return idx.search(q[it[0] % 4], k)

    t, (s_, i_, f_) = timed(fn, reps)
    flops = 2.0 * b * n * d
    out = {"rows": n, "dim": d, "batch": b, "k": k, "metric": metric, "ms": t, "qps": b / t * 1e3, "tflops": flops / t / 1e9,
           "hbm_frac": n * d * 2 / (t * 1e-3) / 1e9 / HBM_GBS, "flagged": int(f_.sum())}
    del idx
    torch.cuda.empty_cache()
    return out


def c4(dev, reps=8, docs=1_000_000, vocab=100_000, dim=1024, batch=256, depth=500, k=100, manager=None):
    """Stage table of config 4.  manager: an already loaded (possibly row-sharded) manager; None = build one here."""
    from b200rag import _lib, bm25, engine, synth
    from b200rag.config import RetrievalConfig
    from b200rag.index_manager import B200IndexManager
    from b200rag.retriever import B200HybridRetriever
    g = torch.Generator(device=dev).manual_seed(0)
    mgr = manager
    if mgr is None:
        mgr = B200IndexManager(semantic_dim=dim, sparse_dim=vocab, domain_dim=8, device=dev, dtype="bf16", enable_sparse=True)
        doc_ptr, term_ids, tf = synth.zipf_corpus_device(docs, vocab, 0, dev)
        w = bm25.bm25_weights_device(doc_ptr, term_ids, tf, vocab)
        for s in range(0, docs, 250_000):
            e = min(docs, s + 250_000)
            a0, a1 = int(doc_ptr[s]), int(doc_ptr[e])
            mgr.add_vectors(torch.randn(e - s, dim, generator=g, device=dev),
                            ((doc_ptr[s: e + 1] - doc_ptr[s]).cpu(), term_ids[a0:a1], w[a0:a1]))
        # token sets for MMR = the unique terms of each document (content = " ".join(f"w{t}")), i.e. the CSR itself
        mgr.set_token_sets(doc_ptr, term_ids.to(torch.int32), vocab)
        del tf, w
    retr = B200HybridRetriever(mgr, RetrievalConfig(top_k=k))
    cfgs = [RetrievalConfig(top_k=depth // 2, enable_mmr=False)] * batch          # searches at depth 2 * top_k = 500
    cfg_h = [RetrievalConfig(top_k=k, enable_mmr=True, mmr_lambda=0.7)] * batch
    qd = [torch.randn(batch, dim, generator=g, device=dev) for _ in range(4)]
    qs = [synth.zipf_queries(batch, vocab, 100 + i, n_terms=8, skip_top=100) for i in range(4)]
    it = [0]

    def dense_fn():
        it[0] += 1
        return mgr.search_batch_ids(qd[it[0] % 4], "semantic_index", depth)

    def sparse_fn():
        it[0] += 1
        return mgr.search_batch_ids(qs[it[0] % 4], "sparse_index", depth)

    t_dense, (ds, di, dc) = timed(dense_fn, reps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); e1.record(); torch.cuda.synchronize()
    _lib.load().b200rag_profile_next_scan(e0.cuda_event, e1.cuda_event)
    dense_fn(); torch.cuda.synchronize()
    t_dense_scan = e0.elapsed_time(e1)
    t_sparse, (ss, si, sc) = timed(sparse_fn, reps)
    df = mgr._sparse.df if getattr(mgr, "world", 1) == 1 else None
    sp_bytes = statistics.mean(int(df[torch.from_numpy(q[1].astype(np.int64)).to(dev)].sum()) * 6 for q in qs) if df is not None else None
    lists = torch.stack([di, si]).contiguous()
    lens = torch.stack([dc, sc]).contiguous()
    wts = torch.tensor([[0.7, 0.3]] * batch, dtype=torch.float64, device=dev)
    t_rrf, fused = timed(lambda: engine.rrf_fuse(lists, lens, wts), reps)
    lam = torch.full((batch,), 0.7, dtype=torch.float64, device=dev)
    ksel = torch.full((batch,), k, dtype=torch.int32, device=dev)
    tok_ptr, tok_ids, tv = mgr.token_sets()
    cand_doc = fused.ids.clamp(min=0).to(torch.int32).contiguous()
    t_mmr, _ = timed(lambda: engine.mmr_select(cand_doc, fused.scores, fused.n, tok_ptr, tok_ids, tv, lam, ksel, k), max(3, reps // 2))

    def hybrid():
        it[0] += 1
        # depth 500 per list (top_k * 2 with top_k = 250) would change k; the reference fuses top_k*2 lists and keeps top_k:
        # run the chain at the candidate depth the config names, then MMR to k
        a = mgr.search_batch_ids(qd[it[0] % 4], "semantic_index", depth)
        b_ = mgr.search_batch_ids(qs[it[0] % 4], "sparse_index", depth)
        return retr.fuse_batch([a, b_], cfg_h)

    t_all, res = timed(hybrid, max(3, reps // 2))

    a_fix = mgr.search_batch_ids(qd[0], "semantic_index", depth)
    b_fix = mgr.search_batch_ids(qs[0], "sparse_index", depth)
    t_fuse, _ = timed(lambda: retr.fuse_batch([a_fix, b_fix], cfg_h), max(3, reps // 2))
    flops = 2.0 * batch * docs * dim
    out = {"workload": f"{docs} x {dim} bf16 dense top-{depth} + BM25 ({docs} docs, {vocab} terms) top-{depth} -> RRF -> MMR 0.7, k {k}, batch {batch}",
           "dense_ms": t_dense, "dense_full_scan_kernel_ms": t_dense_scan, "dense_tflops": flops / t_dense / 1e9,
           "sparse_ms": t_sparse, "sparse_alg_bytes": sp_bytes,
           "sparse_gbs": sp_bytes / t_sparse / 1e6 if sp_bytes else None,
           "sparse_hbm_frac": sp_bytes / (t_sparse * 1e-3) / 1e9 / HBM_GBS if sp_bytes else None,
           "rrf_ms": t_rrf, "mmr_ms": t_mmr, "fused_candidates_mean": float(fused.n.float().mean()),
           "hybrid_ms": t_all, "hybrid_qps": batch / t_all * 1e3,
           "fuse_batch_ms": t_fuse, "results_per_query": float(res.n.float().mean())}
    if manager is None:
        del mgr
        torch.cuda.empty_cache()
    return out


def near_duplicates(dev, n=1_000_000, d=768, b=1024, k=100, copies=64, reps=5):
    """Every row occurs `copies` times (boilerplate / duplicated chunks): whole tie groups straddle rank k, which the
    completeness proof of the tensor-core path cannot certify from k' = k + 28 candidates.  Reports the flagged fraction and
    the step time next to a duplicate-free corpus of the same shape."""
    from b200rag import engine
    g = torch.Generator(device=dev).manual_seed(3)
    base = torch.randn(n // copies, d, generator=g, device=dev)
    q = torch.randn(b, d, generator=g, device=dev)
    out = {"rows": n, "dim": d, "batch": b, "k": k, "copies": copies}
    for name, rows in (("unique", torch.randn(n, d, generator=g, device=dev)), ("duplicated", base.repeat(copies, 1)[:n])):
        idx = engine.DenseIndex(d, "f16", "COSINE", dev, capacity=n)
        idx.add(rows)
        t, (s_, i_, f_) = timed(lambda: idx.search(q, k), reps)
        sx, ix, _ = idx.search(q[:8], k, engine.DENSE_EXACT)
        out[name] = {"ms": t, "flagged_fraction": float((f_ != 0).float().mean()),            # not proven by the first pass
                     "served_by_tier0_fraction": float(((f_ & 2) != 0).float().mean()),
                     "served_by_exact_fallback_fraction": float(((f_ & 1) != 0).float().mean()),
                     "exact": bool(torch.equal(i_[:8], ix) and torch.equal(s_[:8], sx))}
        del idx
        torch.cuda.empty_cache()
    out["slowdown"] = out["duplicated"]["ms"] / out["unique"]["ms"]
    return out


def _max_over_ranks(ms, dev):
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def c5_sharded(dev, world, reps=5, rows_per_gpu=12_500_000, d=384, b=4096, k=10):
    """BASELINE config 5 on the GPUs of this run: rows_per_gpu x world rows (100M at 8 GPUs) x 384 fp16 cosine top-10, batch
    4096, through the row-sharded manager (local scan -> all-gather of k candidates -> merge on every rank)."""
    from b200rag import distributed as bdist
    n_total = rows_per_gpu * world
    mgr = bdist.ShardedIndexManager(n_total, semantic_dim=d, sparse_dim=8, domain_dim=8, device=dev, dtype="f16", enable_sparse=False)
    g = torch.Generator(device=dev).manual_seed(1000 + mgr.rank)
    row = mgr.start
    while row < mgr.end:
        m = min(250_000, mgr.end - row)
        mgr.add_vectors(torch.randn(m, d, generator=g, device=dev))
        row += m
    gq = torch.Generator(device=dev).manual_seed(5)                      # the same queries on every rank
    q = [torch.randn(b, d, generator=gq, device=dev) for _ in range(4)]
    it = [0]

    def fn():
        it[0] += 1
        return mgr.search_batch_ids(q[it[0] % 4], "semantic_index", k)

    t, (s_, i_, c_) = timed(fn, reps)
    t = _max_over_ranks(t, dev)
    flops = 2.0 * b * n_total * d
    out = {"workload": f"{n_total} x {d} fp16 cosine top-{k}, batch {b}, row-sharded over {world} GPU(s) ({rows_per_gpu} rows = "
                       f"{rows_per_gpu * d * 2 / 1e9:.1f} GB per GPU)", "ms": t, "qps": b / t * 1e3, "tflops_all_gpus": flops / t / 1e9,
           "ceiling_qps_at_burst_peak": world * 1627.3e12 / (2.0 * rows_per_gpu * d), "ids_sorted": bool((s_[:, 1:] <= s_[:, :-1]).all())}
    del mgr
    torch.cuda.empty_cache()
    return out


def c4_sharded(dev, world, reps=6, docs=1_000_000, vocab=100_000, dim=1024):
    """BASELINE config 4 with the corpus sharded over the GPUs of this run: dense rows and postings of this rank's document
    range, global idf / avgdl, token sets replicated; searches merge over the ranks, RRF + MMR are split by query."""
    from b200rag import bm25, distributed as bdist, synth
    mgr = bdist.ShardedIndexManager(docs, semantic_dim=dim, sparse_dim=vocab, domain_dim=8, device=dev, dtype="bf16", enable_sparse=True)
    # every rank generates the whole synthetic tf CSR (same seed) and keeps the postings of its own range
    doc_ptr, term_ids, tf = synth.zipf_corpus_device(docs, vocab, 0, dev)
    a, e = mgr.start, mgr.end
    p0, p1 = int(doc_ptr[a]), int(doc_ptr[e])
    lp, lt, lf = doc_ptr[a: e + 1] - doc_ptr[a], term_ids[p0:p1], tf[p0:p1]
    stats = bdist.bm25_global_stats(lp, lt, lf, vocab)
    w = bm25.bm25_weights_device(lp, lt, lf, vocab, stats=stats)
    g = torch.Generator(device=dev).manual_seed(100 + mgr.rank)
    mgr.add_vectors(torch.randn(e - a, dim, generator=g, device=dev), (lp.cpu(), lt, w))
    mgr.set_token_sets(doc_ptr, term_ids.to(torch.int32), vocab)
    del tf, w
    out = c4(dev, reps=reps, docs=docs, vocab=vocab, dim=dim, manager=mgr)
    for key in ("dense_ms", "sparse_ms", "rrf_ms", "mmr_ms", "hybrid_ms"):
        out[key] = _max_over_ranks(out[key], dev)
    out["hybrid_qps"] = 256 / out["hybrid_ms"] * 1e3
    out["workload"] += f"; corpus row-sharded over {world} GPU(s), fusion split by query"
    del mgr
    torch.cuda.empty_cache()
    return out


def filters_and_ingest(dev, rows=10_000_000, reps=10):
    """SURVEY 8f-1 / 8f-2 at corpus scale.  (1) The predicate kernel over 10M rows: three ANDed terms over a float column, an
    integer column and a dictionary-coded string column -> row bit mask (what a NEW filter expression costs before the masked
    search runs; repeated expressions hit the manager's mask cache).  (2) Appending 100K documents to a 1M-document blocked
    postings index: only the tail block is rebuilt."""
    import ctypes
    from b200rag import _lib, bm25, engine, synth
    g = torch.Generator(device=dev).manual_seed(11)
    ent = torch.rand(rows, generator=g, device=dev, dtype=torch.float64)
    chunk = torch.randint(0, 16, (rows,), generator=g, device=dev, dtype=torch.int64)
    code = torch.randint(0, 5000, (rows,), generator=g, device=dev, dtype=torch.int32)
    terms = []
    for col, kind, op, fv, iv in ((ent, _lib.COL_F64, _lib.OP_GE, 0.5, 0), (chunk, _lib.COL_I64, _lib.OP_LT, 0.0, 12),
                                  (code, _lib.COL_CODE, _lib.OP_NE, 0.0, 7)):
        t = _lib.FilterTerm()
        t.column, t.lut, t.fvalue, t.ivalue, t.kind, t.op, t.lut_size = col.data_ptr(), None, fv, iv, kind, op, 0
        terms.append(t)
    ms, (words, count) = timed(lambda: engine.filter_mask(terms, rows, dev), reps)
    want = int(((ent >= 0.5) & (chunk < 12) & (code != 7)).sum())
    out = {"predicate": {"rows": rows, "terms": 3, "ms": ms, "allowed_rows": int(count.item()), "matches_torch": int(count.item()) == want,
                         "bytes_read": rows * (8 + 8 + 4), "gbs": rows * 20 / (ms * 1e-3) / 1e9}}
    del ent, chunk, code
    docs, add, vocab = 1_000_000, 100_000, 100_000
    dp, ti, tf = synth.zipf_corpus_device(docs + add, vocab, 21, dev)
    w = bm25.bm25_weights_device(dp, ti, tf, vocab)
    a1 = int(dp[docs])
    t0 = time.perf_counter()
    sidx = engine.SparseIndex(dp[: docs + 1].cpu(), ti[:a1], w[:a1], vocab, dev)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    b0 = sidx.blocks_built
    t0 = time.perf_counter()
    sidx.append((dp[docs:] - dp[docs]).cpu(), ti[a1:], w[a1:])
    torch.cuda.synchronize()
    t_app = time.perf_counter() - t0
    out["sparse_append"] = {"docs": docs, "appended": add, "build_s": t_build, "append_s": t_app,
                            "blocks_total": int(sidx._n_blocks), "blocks_rebuilt_by_append": int(sidx.blocks_built - b0)}
    return out
