"""BASELINE config 4 on one B200, stage by stage: dense (1M x 1024 bf16 cosine, top-500) + BM25 (1M docs, 100K-term
Zipf vocabulary, CSR postings, top-500) -> weighted RRF -> MMR lambda 0.7 over the <= 1000 fused candidates, k = 100,
query batch 256.  Every stage is timed with CUDA events in steady state; the sparse scan is reported against its HBM
roofline (bytes = sum over query terms of df(t) * 6, computed exactly from the index).

Also: config 2 (1M x 768 fp16 IP top-100, batch 1024) and one GPU's shard of config 5 (12.5M x 384 fp16 top-10, batch 4096).
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]
from b200rag import bm25, engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=1_000_000)
ap.add_argument("--vocab", type=int, default=100_000)
ap.add_argument("--dim", type=int, default=1024)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--depth", type=int, default=500)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--skip", default="")
ap.add_argument("--block-docs", type=int, default=16384)
args = ap.parse_args()
dev = torch.device("cuda:0")
HBM = 6545e9
out = {}


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), r


if "c4" not in args.skip:
    t0 = time.time()
    g = torch.Generator(device=dev).manual_seed(0)
    dense = engine.DenseIndex(args.dim, "bf16", "COSINE", dev, capacity=args.docs)
    for s in range(0, args.docs, 250_000):
        dense.add(torch.randn(min(250_000, args.docs - s), args.dim, generator=g, device=dev))
    doc_ptr, term_ids, tf = synth.zipf_corpus_device(args.docs, args.vocab, 0, dev)
    w = bm25.bm25_weights_device(doc_ptr, term_ids, tf, args.vocab)
    sparse = engine.SparseIndex(doc_ptr, term_ids, w, args.vocab, dev, block_docs=args.block_docs)
    nnz = int(term_ids.numel())
    # token sets for MMR = the unique terms of each document (content = " ".join(f"w{t}")), i.e. the CSR itself
    tok_ptr, tok_ids = doc_ptr.to(dev), term_ids.to(torch.int32).contiguous()
    del tf, w
    qd = [torch.randn(args.batch, args.dim, generator=g, device=dev) for _ in range(4)]
    qs = [synth.zipf_queries(args.batch, args.vocab, 100 + i, n_terms=8, skip_top=100) for i in range(4)]
    print(f"C4 build: {args.docs} docs, dim {args.dim} bf16, vocab {args.vocab}, nnz {nnz} ({time.time() - t0:.1f} s)", flush=True)
    B, K2, K = args.batch, args.depth, args.k
    it = [0]

    def dense_fn():
        it[0] += 1
        return dense.search(qd[it[0] % 4], K2)

    def sparse_fn():
        it[0] += 1
        qp, qt, qv = qs[it[0] % 4]
        return sparse.search(qp, qt, qv, K2)

    t_dense, (ds, di, _) = timed(dense_fn, args.reps)
    from b200rag import _lib
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); e1.record(); torch.cuda.synchronize()
    _lib.load().b200rag_profile_next_scan(e0.cuda_event, e1.cuda_event)
    dense_fn(); torch.cuda.synchronize()
    t_dense_scan = e0.elapsed_time(e1)
    t_sparse, (ss, si, sc) = timed(sparse_fn, args.reps)
    sp_bytes = statistics.mean(sparse.query_bytes(q[0], q[1]) for q in qs)
    lists = torch.stack([di, si]).contiguous()
    lens = torch.stack([torch.full((B,), K2, dtype=torch.int32, device=dev), sc]).contiguous()
    wts = torch.tensor([[0.7, 0.3]] * B, dtype=torch.float64, device=dev)
    t_rrf, fused = timed(lambda: engine.rrf_fuse(lists, lens, wts), args.reps)
    lam = torch.full((B,), 0.7, dtype=torch.float64, device=dev)
    ksel = torch.full((B,), K, dtype=torch.int32, device=dev)
    cand_doc = fused.ids.clamp(min=0).to(torch.int32).contiguous()
    t_mmr, (picks, pn) = timed(lambda: engine.mmr_select(cand_doc, fused.scores, fused.n, tok_ptr, tok_ids, args.vocab, lam, ksel, K),
                               max(3, args.reps // 4))

    def hybrid():
        it[0] += 1
        a = dense.search(qd[it[0] % 4], K2)
        qp, qt, qv = qs[it[0] % 4]
        b = sparse.search(qp, qt, qv, K2)
        li = torch.stack([a[1], b[1]]).contiguous()
        ln = torch.stack([torch.full((B,), K2, dtype=torch.int32, device=dev), b[2]]).contiguous()
        f = engine.rrf_fuse(li, ln, wts)
        return engine.mmr_select(f.ids.clamp(min=0).to(torch.int32).contiguous(), f.scores, f.n, tok_ptr, tok_ids, args.vocab,
                                 lam, ksel, K)

    t_all, _ = timed(hybrid, max(3, args.reps // 4))
    flops = 2.0 * B * args.docs * args.dim
    dbytes = args.docs * args.dim * 2
    out["c4"] = {"dense_ms": t_dense, "dense_full_scan_kernel_ms": t_dense_scan, "dense_tflops": flops / t_dense / 1e9, "dense_hbm_frac": dbytes / (t_dense * 1e-3) / HBM,
                 "sparse_ms": t_sparse, "sparse_alg_bytes": sp_bytes, "sparse_gbs": sp_bytes / t_sparse / 1e6,
                 "sparse_hbm_frac": sp_bytes / (t_sparse * 1e-3) / HBM, "rrf_ms": t_rrf, "mmr_ms": t_mmr,
                 "fused_candidates_mean": float(fused.n.float().mean()), "hybrid_ms": t_all, "hybrid_qps": B / t_all * 1e3}
    print(json.dumps(out["c4"]), flush=True)
    del dense, sparse, doc_ptr, term_ids, tok_ptr, tok_ids
    torch.cuda.empty_cache()

if "c1" not in args.skip:
    # BASELINE config 1: 100K chunks, 384-d, BM25 over a 30K-term vocabulary, alpha 0.7, top_k 20 (dense top-40 + sparse top-40
    # -> RRF -> top 20), on the GPU in batches of 256 and on the host cores through the oracle's restatement of the chain.
    from oracle import pipeline as opipe
    n1, d1, v1, b1, tk = 100_000, 384, 30_000, 256, 20
    x = synth.dense_rows(n1, d1, 0)
    dp, ti, tf = synth.zipf_corpus(n1, v1, 0)
    w = bm25.bm25_weights(dp, ti, tf, v1)
    dense = engine.DenseIndex(d1, "f16", "COSINE", dev)
    dense.add(torch.from_numpy(x))
    sparse = engine.SparseIndex(dp, ti, w, v1, dev, block_docs=args.block_docs)
    q = synth.dense_rows(b1, d1, 1000)
    qp, qt, qv = synth.zipf_queries(b1, v1, 1, n_terms=8, skip_top=100)
    qd = torch.from_numpy(q).to(dev)
    wts = torch.tensor([[0.7, 0.3]] * b1, dtype=torch.float64, device=dev)

    def c1_step():
        ds, di, _ = dense.search(qd, 2 * tk)
        ss, si, sc = sparse.search(qp, qt, qv, 2 * tk)
        li = torch.stack([di, si]).contiguous()
        ln = torch.stack([torch.full((b1,), 2 * tk, dtype=torch.int32, device=dev), sc]).contiguous()
        return engine.rrf_fuse(li, ln, wts)

    t_c1, fused = timed(c1_step, args.reps)
    corpus = opipe.ArrayCorpus(x, None, dp, ti, w, v1, None)
    n_cpu = 40
    t0 = time.time()
    same = True
    for r in range(n_cpu):
        sq = {"indices": qt[qp[r]:qp[r + 1]].tolist(), "values": qv[qp[r]:qp[r + 1]].tolist()}
        ids, scs, _ = opipe.retrieve(corpus, q[r], sq, None, tk)
        same &= fused.ids[r, :tk].cpu().tolist() == ids and fused.scores[r, :tk].cpu().tolist() == scs
    t_cpu = (time.time() - t0) / n_cpu
    out["c1"] = {"gpu_ms_per_batch": t_c1, "gpu_qps": b1 / t_c1 * 1e3, "cpu_port_ms_per_query": t_cpu * 1e3, "cpu_port_qps": 1.0 / t_cpu,
                 "cpu_cores": os.cpu_count(), "bit_exact_vs_oracle_pipeline": bool(same), "checked_queries": n_cpu}
    print("c1", json.dumps(out["c1"]), flush=True)
    del dense, sparse
    torch.cuda.empty_cache()

for name, (n, d, b, k, metric) in (("c2", (1_000_000, 768, 1024, 100, "IP")), ("c5_shard", (12_500_000, 384, 4096, 10, "COSINE"))):
    if name in args.skip:
        continue
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

    t, (s_, i_, f_) = timed(fn, args.reps)
    flops = 2.0 * b * n * d
    out[name] = {"rows": n, "dim": d, "batch": b, "k": k, "ms": t, "qps": b / t * 1e3, "tflops": flops / t / 1e9,
                 "hbm_frac": n * d * 2 / (t * 1e-3) / HBM, "flagged": int(f_.sum())}
    print(name, json.dumps(out[name]), flush=True)
    del idx
    torch.cuda.empty_cache()
