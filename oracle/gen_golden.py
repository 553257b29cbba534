"""Generate tests/golden/fusion_golden.json by EXECUTING THE REFERENCE'S OWN CODE -- test infrastructure.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden
The reference package is imported unmodified through oracle/ref_import.py; the functions executed are
HybridRetriever._fuse_results / _mmr_diversify / rerank / _build_filter_expression, QueryClassifier.classify
and HybridRetriever._build_default_profiles (reference src/advanced_rag/retrieval.py:33-67,142-213,421-563,
573-632) and LearnedRanker.score (ranker.py:109-125).  Floats are stored as C99 hex strings so the
comparison is bit exact.
"""
from __future__ import annotations

import copy
import json
import os
import random

from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "fusion_golden.json")


def _doc(i: int, rng: random.Random, vocab: int, n_tok: int, with_meta: bool = True):
    toks = [f"w{rng.randrange(vocab)}" for _ in range(n_tok)]
    # mixed case + repeated tokens exercise .lower() and set semantics
    if toks and rng.random() < 0.3:
        toks[0] = toks[0].upper()
    d = {"id": f"c{i:06d}", "content": " ".join(toks), "score": rng.random()}
    if with_meta:
        d["metadata"] = {"doc_id": f"d{i // 4}", "chunk_index": i % 4}
    return d


def _ranked(pool, rng: random.Random, n: int):
    picks = rng.sample(range(len(pool)), n)
    out = []
    for r, i in enumerate(picks):
        d = copy.deepcopy(pool[i])
        d["score"] = 1.0 - 0.01 * r
        out.append(d)
    return out


def main() -> None:
    ar = ref_import.load()
    from advanced_rag.retrieval import HybridRetriever, RetrievalConfig, QueryClassifier
    from advanced_rag.ranker import LearnedRanker

    cases = {"rrf": [], "mmr": [], "rerank": [], "filters": [], "classifier": [], "profiles": [],
             "reference_version": ar.__version__}

    # ---------------- RRF (with and without domain list, ties, empties, duplicates) -------------
    rng = random.Random(1234)
    specs = [
        # (pool, n_sem, n_sparse, n_domain, dense_w, sparse_w)
        (50, 10, 10, 0, 0.7, 0.3),
        (50, 40, 40, 0, 0.7, 0.3),
        (60, 40, 40, 20, 0.7, 0.3),
        (30, 20, 20, 10, 0.5, 0.5),      # equal weights -> many exact ties
        (30, 25, 25, 25, 0.2, 0.2),      # all three weights equal (domain is hard-coded 0.2)
        (10, 0, 5, 0, 0.7, 0.3),         # empty semantic list
        (10, 5, 0, 0, 0.7, 0.3),         # empty sparse list
        (10, 0, 0, 0, 0.7, 0.3),         # everything empty
        (200, 100, 100, 50, 0.1, 0.9),
        (1200, 500, 500, 0, 0.7, 0.3),   # C4-like candidate depth
        (80, 40, 40, 40, 1.0, 0.0),      # zero sparse weight
    ]
    for pool_n, ns, np_, nd, dw, sw in specs:
        pool = [_doc(i, rng, 50, 6) for i in range(pool_n)]
        sem = _ranked(pool, rng, ns)
        spa = _ranked(pool, rng, np_)
        dom = _ranked(pool, rng, nd)
        cfg = RetrievalConfig(dense_weight=dw, sparse_weight=sw, top_k=20)
        r = HybridRetriever(index_manager=None, config=cfg)
        fused = r._fuse_results(copy.deepcopy(sem), copy.deepcopy(spa), copy.deepcopy(dom))
        cases["rrf"].append({
            "semantic": [d["id"] for d in sem], "sparse": [d["id"] for d in spa],
            "domain": [d["id"] for d in dom], "dense_weight": dw, "sparse_weight": sw,
            "out_ids": [d["id"] for d in fused],
            "out_scores_hex": [float(d["score"]).hex() for d in fused],
            "out_methods": [sorted(d["retrieval_methods"]) for d in fused],
        })
    # a list that repeats an id (the reference adds the contribution twice)
    sem = [{"id": "A", "content": "x", "score": 1.0}, {"id": "B", "content": "y", "score": 0.9},
           {"id": "A", "content": "x", "score": 0.8}]
    spa = [{"id": "B", "content": "y", "score": 1.0}]
    r = HybridRetriever(index_manager=None, config=RetrievalConfig())
    fused = r._fuse_results(copy.deepcopy(sem), copy.deepcopy(spa), [])
    cases["rrf"].append({
        "semantic": ["A", "B", "A"], "sparse": ["B"], "domain": [], "dense_weight": 0.7, "sparse_weight": 0.3,
        "out_ids": [d["id"] for d in fused], "out_scores_hex": [float(d["score"]).hex() for d in fused],
        "out_methods": [sorted(d["retrieval_methods"]) for d in fused]})

    # the reference's own test vectors (test_extended.py:81-130 and :189-213)
    r = HybridRetriever(index_manager=None)
    fused = r._fuse_results(
        semantic_results=[{"id": "A", "content": "x", "score": 0.9}, {"id": "B", "content": "y", "score": 0.8}],
        sparse_results=[{"id": "A", "content": "x", "score": 0.7}, {"id": "C", "content": "z", "score": 0.6}],
        domain_results=[])
    cases["rrf"].append({
        "semantic": ["A", "B"], "sparse": ["A", "C"], "domain": [], "dense_weight": 0.7, "sparse_weight": 0.3,
        "out_ids": [d["id"] for d in fused], "out_scores_hex": [float(d["score"]).hex() for d in fused],
        "out_methods": [sorted(d["retrieval_methods"]) for d in fused], "source": "test_extended.py:103-111"})

    # ---------------- fuse + MMR --------------------------------------------------------------
    mmr_specs = [
        # (pool, n_sem, n_sparse, n_dom, vocab, tokens/doc, top_k, lambda)
        (40, 20, 20, 0, 30, 8, 10, 0.7),
        (40, 20, 20, 10, 30, 8, 20, 0.5),
        (60, 40, 40, 0, 12, 5, 20, 0.8),     # tiny vocab -> many identical Jaccards -> ties
        (30, 20, 20, 0, 30, 0, 5, 0.7),      # empty contents -> sim 0
        (100, 60, 60, 0, 200, 30, 30, 0.6),
        (20, 10, 10, 0, 30, 8, 50, 0.7),     # k larger than candidate count
        (150, 100, 100, 0, 400, 40, 40, 0.0),  # lambda 0: pure diversity
        (50, 30, 30, 0, 60, 10, 10, 1.0),    # lambda 1: pure relevance
    ]
    for pool_n, ns, np_, nd, vocab, ntok, top_k, lam in mmr_specs:
        pool = [_doc(i, rng, vocab, ntok) for i in range(pool_n)]
        sem = _ranked(pool, rng, ns)
        spa = _ranked(pool, rng, np_)
        dom = _ranked(pool, rng, nd)
        cfg = RetrievalConfig(top_k=top_k, enable_mmr=True, mmr_lambda=lam)
        r = HybridRetriever(index_manager=None, config=cfg)
        out = r._fuse_results(copy.deepcopy(sem), copy.deepcopy(spa), copy.deepcopy(dom))
        cases["mmr"].append({
            "contents": {d["id"]: d["content"] for d in pool},
            "semantic": [d["id"] for d in sem], "sparse": [d["id"] for d in spa],
            "domain": [d["id"] for d in dom], "dense_weight": 0.7, "sparse_weight": 0.3,
            "top_k": top_k, "mmr_lambda": lam,
            "out_ids": [d["id"] for d in out],
            "out_scores_hex": [float(d["score"]).hex() for d in out],
        })
    # reference test vector test_extended.py:189-213
    cfg = RetrievalConfig(hybrid_alpha=0.7, top_k=3, enable_mmr=True, mmr_lambda=0.6)
    r = HybridRetriever(index_manager=None, config=cfg)
    sem = [{"id": "A", "content": "alpha alpha content one", "score": 0.95},
           {"id": "B", "content": "bravo content two", "score": 0.85},
           {"id": "C", "content": "alpha content three", "score": 0.80}]
    spa = [{"id": "A", "content": "alpha alpha content one", "score": 0.75},
           {"id": "D", "content": "delta unique different", "score": 0.70},
           {"id": "E", "content": "echo also different", "score": 0.65}]
    out = r._fuse_results(copy.deepcopy(sem), copy.deepcopy(spa), [])
    cases["mmr"].append({
        "contents": {d["id"]: d["content"] for d in sem + spa},
        "semantic": ["A", "B", "C"], "sparse": ["A", "D", "E"], "domain": [],
        "dense_weight": 0.7, "sparse_weight": 0.3, "top_k": 3, "mmr_lambda": 0.6,
        "out_ids": [d["id"] for d in out], "out_scores_hex": [float(d["score"]).hex() for d in out],
        "source": "test_extended.py:189-213"})

    # ---------------- rerank with the learned ranker (deterministic branch) --------------------
    for n, top_k in [(20, 5), (8, 10), (40, 5)]:
        pool = [_doc(i, rng, 50, 6) for i in range(60)]
        sem = _ranked(pool, rng, n)
        spa = _ranked(pool, rng, n)
        cfg = RetrievalConfig(top_k=20, enable_learned_ranker=True)
        r = HybridRetriever(index_manager=None, config=cfg, learned_ranker=LearnedRanker())
        fused = r._fuse_results(copy.deepcopy(sem), copy.deepcopy(spa), [])[:cfg.top_k]
        fused_in = [(d["id"], float(d["score"]).hex(), len(d["retrieval_methods"])) for d in fused]
        rer = ref_import.run(r.rerank("q", fused, top_k=top_k))
        cases["rerank"].append({
            "in_ids": [x[0] for x in fused_in], "in_scores_hex": [x[1] for x in fused_in],
            "in_n_methods": [x[2] for x in fused_in], "top_k": top_k,
            "out_ids": [d["id"] for d in rer], "out_scores_hex": [float(d["score"]).hex() for d in rer]})

    # ---------------- evaluation: mean pairwise token-set Jaccard (reference evaluation.py:327-344) ----------
    from advanced_rag.evaluation import RAGEvaluator
    ev = RAGEvaluator()
    cases["pairwise_similarity"] = []
    for n_res, vocab, max_len in [(5, 12, 8), (20, 40, 25), (2, 5, 4), (1, 5, 4), (30, 300, 60), (100, 60, 30), (17, 9, 6)]:
        res = [_doc(i, rng, vocab, rng.randint(0, max_len)) for i in range(n_res)]
        if n_res >= 5:
            res[2]["content"] = ""                       # an empty token set: its pairs are skipped (:338)
            res[4]["content"] = res[3]["content"].upper()  # .lower() makes them identical
        val = ev._calculate_pairwise_similarity(res)
        cases["pairwise_similarity"].append({"contents": [d["content"] for d in res], "mean_hex": float(val).hex(),
                                             "diversity_hex": float(ev._calculate_diversity(res)).hex()})

    # ---------------- filter expressions (row S3) ------------------------------------------------
    r = HybridRetriever(index_manager=None)
    for f in [
        {"doc_id": 'doc"123', "entropy": {"$gte": 0.2}},
        {"redundancy": {"$lt": 0.5, "$gt": 0.1, "$eq": 0.2, "$ne": 0.3}, "chunk_index": 1},
        {},
        {"timestamp": {"$gte": "2024-01-01"}, "token_count": {"$lte": 512}},
        {"chunk_id": "a\\b"},
        {"domain_density": 0.5},
    ]:
        cases["filters"].append({"filters": f, "expr": r._build_filter_expression(f)})
    for bad in [{"evil": 1}, {"doc_id": {"$regex": "x"}}, {"doc_id": [1, 2]}, {"entropy": {"$gte": [1]}}]:
        try:
            r._build_filter_expression(bad)
            err = None
        except ValueError as e:
            err = str(e)
        cases["filters"].append({"filters": bad, "error": err})

    # ---------------- classifier + profiles (row R3) --------------------------------------------
    qc = QueryClassifier()
    for q in ["What is vector search?", "I see an error: connection failed", "", "x" * 250,
              "Please provide a summary or overview of RAG.", "how do I fix a stack trace", "hello world",
              "Compare the two approaches in depth " * 8, "tl;dr of the doc?", "why?"]:
        cases["classifier"].append({"query": q, "label": qc.classify(q)})
    for base in [dict(top_k=20, rerank_top_k=5), dict(top_k=50, rerank_top_k=20), dict(top_k=5, rerank_top_k=5),
                 dict(top_k=150, rerank_top_k=120)]:
        r = HybridRetriever(index_manager=None, config=RetrievalConfig(**base))
        prof = {name: {"top_k": c.top_k, "rerank_top_k": c.rerank_top_k, "enable_mmr": c.enable_mmr,
                       "mmr_lambda": c.mmr_lambda, "enable_reranking": c.enable_reranking}
                for name, c in r.profiles.items()}
        cases["profiles"].append({"base": base, "profiles": prof})

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(cases, f, indent=1, sort_keys=True)
    print("wrote", OUT, {k: len(v) for k, v in cases.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
