"""Host-side mirror of the reference interface (b200rag/config.py, retriever.py, index_manager.py): everything that
runs without a GPU, checked against golden vectors produced by EXECUTING the reference (tests/golden/*.json)."""
import asyncio

import numpy as np
import pytest

from util import load_golden


def _run(coro):
    loop = asyncio.new_event_loop()
    try:
        return loop.run_until_complete(coro)
    finally:
        loop.close()


@pytest.fixture(scope="module")
def golden():
    return load_golden()


def _retriever(**cfg):
    from b200rag.config import RetrievalConfig
    from b200rag.retriever import B200HybridRetriever
    return B200HybridRetriever(index_manager=None, config=RetrievalConfig(**cfg))


def test_query_classifier_matches_reference(golden):
    from b200rag.config import QueryClassifier
    qc = QueryClassifier()
    assert len(golden["classifier"]) >= 10
    for c in golden["classifier"]:
        assert qc.classify(c["query"]) == c["label"], c["query"]
    assert qc.classify(None) == "default"


def test_default_profiles_match_reference(golden):
    from b200rag.config import RetrievalConfig, build_default_profiles
    for case in golden["profiles"]:
        base = RetrievalConfig(**case["base"])
        prof = build_default_profiles(base)
        assert prof["default"] is base
        assert set(prof) == set(case["profiles"])
        for name, want in case["profiles"].items():
            got = prof[name]
            assert {k: getattr(got, k) for k in want} == want, (case["base"], name)
    # the reference test's inequalities (test_extended.py:133-186)
    prof = build_default_profiles(RetrievalConfig(top_k=20))
    assert prof["faq"].top_k <= 10 and prof["troubleshooting"].top_k >= 30 and prof["summary"].top_k >= 40
    assert prof["troubleshooting"].enable_mmr and prof["analysis"].enable_mmr and not prof["summary"].enable_reranking


def test_retrieval_config_defaults_match_reference():
    from b200rag.config import RetrievalConfig
    c = RetrievalConfig()
    assert (c.hybrid_alpha, c.top_k, c.rerank_top_k, c.dense_weight, c.sparse_weight) == (0.7, 20, 5, 0.7, 0.3)
    assert (c.enable_reranking, c.enable_mmr, c.mmr_lambda, c.enable_learned_ranker) == (True, False, 0.7, False)
    assert c.semantic_search_params == {"metric_type": "COSINE", "params": {"ef": 64}}
    assert c.sparse_search_params == {"metric_type": "IP", "params": {"drop_ratio_search": 0.2}}


def test_filter_expressions_match_reference(golden):
    r = _retriever()
    n_err = 0
    for case in golden["filters"]:
        if "error" in case:
            n_err += 1
            with pytest.raises(ValueError) as ei:
                r._build_filter_expression(case["filters"])
            assert str(ei.value) == case["error"]
        else:
            # the fixture was dumped with sorted keys, the reference iterates in dict order: compare the terms
            got = r._build_filter_expression(case["filters"])
            assert (got is None) == (case["expr"] is None)
            if got is not None:
                assert sorted(got.split(" and ")) == sorted(case["expr"].split(" and ")), case["filters"]
    assert n_err >= 4


def _store(n=50):
    from b200rag.index_manager import PayloadStore
    store = PayloadStore()
    meta = [{"doc_id": 'doc"123' if i % 2 else "a\\b", "chunk_index": i % 4, "entropy": i / 50.0,
             "redundancy": 0.2 if i % 5 == 0 else 0.3, "domain_density": 0.5 if i < 10 else 0.25,
             "timestamp": f"2024-0{1 + i % 9}-01", "token_count": 100 + 10 * i} for i in range(n)]
    store.commit([f"c{i}" for i in range(n)], ["x"] * n, store.convert(n, meta))
    return store, meta


def test_filter_expression_round_trips_through_the_evaluator(golden):
    """The strings _build_filter_expression emits are what B200IndexManager.search receives as `filters`: the manager's
    parser must accept every one of them and select the right rows (host restatement of the predicate semantics; the GPU
    kernel is checked against it in tests/test_gpu_reference_behaviours.py)."""
    from b200rag.index_manager import _parse_filter, eval_filter_host as _eval_filter
    store, _ = _store()
    r = _retriever()
    for case in golden["filters"]:
        if "error" in case or not case["expr"]:
            continue
        _parse_filter(case["expr"])
    m = _eval_filter(store, r._build_filter_expression({"doc_id": 'doc"123', "entropy": {"$gte": 0.2}}))
    assert m.tolist() == [i % 2 == 1 and i / 50.0 >= 0.2 for i in range(50)]
    m = _eval_filter(store, r._build_filter_expression({"chunk_id": "c7"}))
    assert np.flatnonzero(m).tolist() == [7]
    m = _eval_filter(store, r._build_filter_expression({"timestamp": {"$gte": "2024-05-01"}, "token_count": {"$lte": 300}}))
    assert m.tolist() == [f"2024-0{1 + i % 9}-01" >= "2024-05-01" and 100 + 10 * i <= 300 for i in range(50)]
    m = _eval_filter(store, r._build_filter_expression({"doc_id": "a\\b", "domain_density": 0.5}))
    assert m.tolist() == [i % 2 == 0 and i < 10 for i in range(50)]
    # a literal of the wrong type matches nothing; != never matches a missing value
    assert not _eval_filter(store, 'entropy == "0.5"').any() and not _eval_filter(store, "doc_id == 3").any()
    assert _eval_filter(store, "chunk_index >= 1.5").tolist() == [i % 4 >= 2 for i in range(50)]
    with pytest.raises(ValueError):
        _eval_filter(store, "evil == 1")
    with pytest.raises(ValueError):
        _eval_filter(store, "entropy >= 0.2 or entropy < 0.1")


def test_payload_store_is_columnar_and_hits_are_fresh_dicts():
    """Reference hit shape (indexing.py:534-551) from typed columns: values come back as the schema types them
    (indexing.py:191-225), missing values as None, and every call builds fresh dicts (retrieval.py:361-363 mutates them)."""
    from b200rag.index_manager import HitLists, PayloadStore, SearchArrays
    store, meta = _store(20)
    extra = store.convert(2, [None, {"doc_id": 7, "entropy": None, "chunk_index": 3}])
    store.commit(["n0", "n1"], ["", "y z"], extra)
    assert len(store) == 22 and store.num["entropy"].view.dtype == np.float64 and store.codes["doc_id"].view.dtype == np.int32
    rows = np.asarray([3, 21, 20, 0], dtype=np.int64)
    hits = store.hits(rows, np.asarray([0.5, 0.25, 0.125, 0.0]))
    assert [h["id"] for h in hits] == ["c3", "n1", "n0", "c0"] and [h["score"] for h in hits] == [0.5, 0.25, 0.125, 0.0]
    assert hits[0]["metadata"] == {k: meta[3][k] for k in ("doc_id", "chunk_index", "entropy", "redundancy", "domain_density", "timestamp")}
    assert hits[1]["metadata"] == {"doc_id": "7", "chunk_index": 3, "entropy": None, "redundancy": None, "domain_density": None,
                                   "timestamp": None}
    assert all(v is None for v in hits[2]["metadata"].values()) and hits[2]["content"] == ""
    assert isinstance(hits[0]["metadata"]["chunk_index"], int) and isinstance(hits[0]["metadata"]["entropy"], float)
    again = store.hits(rows, np.zeros(4))
    assert again[0] is not hits[0] and again[0]["metadata"] is not hits[0]["metadata"]
    assert store.hit(3, 0.5) == hits[0]
    # lazy List[List[dict]] view over a columnar result; counts and -1 padding are honoured
    arr = SearchArrays(np.asarray([[3, 21, -1], [0, -1, -1], [-1, -1, -1]]), np.asarray([[.5, .25, 0], [1., 0, 0], [0, 0, 0.]]),
                       np.asarray([2, 1, 0], dtype=np.int32), store)
    lazy = arr.hits()
    assert isinstance(lazy, HitLists) and len(lazy) == 3
    assert [h["id"] for h in lazy[0]] == ["c3", "n1"] and lazy[2] == [] and lazy[-2][0]["id"] == "c0"
    assert lazy[0][0] is not lazy[0][0]
    assert [[h["id"] for h in q] for q in lazy.materialize()] == [["c3", "n1"], ["c0"], []] == arr.chunk_ids()
    assert [[h["id"] for h in q] for q in lazy] == arr.chunk_ids()
    # a value the schema type cannot hold is refused before anything is stored
    with pytest.raises(ValueError):
        store.convert(1, [{"entropy": "high"}])
    with pytest.raises(ValueError):
        store.convert(2, [{}])
    assert len(store) == 22 and len(store.dict_values["doc_id"]) == 3
    store.keep(np.asarray([0, 21]))
    assert store.ids == ["c0", "n1"] and store.codes["doc_id"].view.tolist() == [store.dict_code["doc_id"]["a\\b"], 2]


def test_csr_take_selects_rows():
    from b200rag.index_manager import _csr_take
    ptr = np.asarray([0, 2, 2, 5, 6], np.int64)
    a, b = np.arange(6, dtype=np.int32), np.arange(6, dtype=np.float32) * 0.5
    p2, (a2, b2) = _csr_take(ptr, [a, b], np.asarray([2, 0, 1]))
    assert p2.tolist() == [0, 3, 5, 5] and a2.tolist() == [2, 3, 4, 0, 1] and b2.tolist() == [1.0, 1.5, 2.0, 0.0, 0.5]
    p3, (a3,) = _csr_take(ptr, [a], np.zeros(0, np.int64))
    assert p3.tolist() == [0] and a3.size == 0


def test_learned_rerank_matches_reference(golden):
    from b200rag.config import LearnedRanker, RetrievalConfig
    from b200rag.retriever import B200HybridRetriever
    for case in golden["rerank"]:
        hits = [{"id": i, "content": "", "score": float.fromhex(s), "retrieval_methods": ["semantic", "sparse"][:m]}
                for i, s, m in zip(case["in_ids"], case["in_scores_hex"], case["in_n_methods"])]
        r = B200HybridRetriever(None, RetrievalConfig(top_k=20, enable_learned_ranker=True), learned_ranker=LearnedRanker())
        out = _run(r.rerank("q", hits, top_k=case["top_k"]))
        assert [h["id"] for h in out] == case["out_ids"]
        assert [float(h["score"]).hex() for h in out] == case["out_scores_hex"]
        assert all(h["score"] == h["rerank_score"] and "original_retrieval_score" in h for h in out)


def test_rerank_branches_like_the_reference_tests():
    """reference test_extended.py:238-273: placeholder branch truncates, external reranker orders by its scores,
    disabled reranking slices only."""
    from b200rag.config import RetrievalConfig
    from b200rag.retriever import B200HybridRetriever
    res = [{"id": "A", "content": "a", "score": 0.5}, {"id": "B", "content": "b", "score": 0.4}, {"id": "C", "content": "c", "score": 0.3}]
    r = B200HybridRetriever(None, RetrievalConfig())
    out = _run(r.rerank("q", [dict(x) for x in res], top_k=2))
    assert len(out) == 2 and all("rerank_score" in x for x in out)

    class Ext:
        async def score(self, pairs):
            return [0.1, 0.9, 0.5][: len(pairs)]
    r.reranker = Ext()
    out = _run(r.rerank("q", [dict(x) for x in res], top_k=2))
    assert [x["id"] for x in out] == ["B", "C"]
    r2 = B200HybridRetriever(None, RetrievalConfig(enable_reranking=False))
    assert [x["id"] for x in _run(r2.rerank("q", [dict(x) for x in res], top_k=2))] == ["A", "B"]
    assert _run(r2.rerank("q", [])) == []


def test_learned_ranker_formula_and_feedback():
    from b200rag.config import LearnedRanker, LearnedRankerConfig
    lr = LearnedRanker(LearnedRankerConfig(base_weight=2.0, method_bonus=0.5, recency_weight=0.25))
    hits = [{"id": "a", "score": 0.1, "retrieval_methods": ["semantic", "sparse"], "metadata": {"recency": 0.5}},
            {"id": "b", "score": 0.2}]
    assert _run(lr.score("q", hits)) == [2.0 * 0.1 + 0.5 * 2.0 + 0.25 * 0.5, 2.0 * 0.2]
    lr.update_from_feedback("q", hits, [{"id": "a", "label": 1.0}, {"id": "zz", "label": 0.0}])
    assert len(lr.training_examples) == 1 and lr.training_examples[0].label == 1.0


def test_retriever_without_a_gpu_fails_loudly_in_fusion():
    """No CPU fallback: the fusion entry points need the CUDA library and a device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    r = _retriever()
    with pytest.raises(Exception):
        r._fuse_results([{"id": "A", "content": "x", "score": 1.0}], [{"id": "B", "content": "y", "score": 1.0}], [])
    from b200rag.index_manager import B200IndexManager
    with pytest.raises(Exception):
        B200IndexManager(semantic_dim=8, domain_dim=8, device="cpu")


class _StubManager:
    """search_batch stand-in for the micro-batcher: echoes (collection, top_k, query checksum) per query."""

    def __init__(self, fail=False):
        self.calls, self.fail = [], fail

    def search_batch(self, batch, collection, top_k, filters):
        self.calls.append((collection, top_k, filters, len(batch)))
        if self.fail:
            raise RuntimeError("index down")
        if collection == "sparse_index":
            return [[{"id": f"s{sum(q['indices'])}", "score": float(top_k)}] for q in batch]
        return [[{"id": f"d{int(row.sum())}", "score": float(top_k)}] for row in batch]


def test_micro_batcher_groups_concurrent_requests_and_routes_results():
    from b200rag.index_manager import _MicroBatcher
    stub = _StubManager()
    mb = _MicroBatcher(stub, max_batch=64, max_wait_s=0.02)

    async def storm():
        dense = [mb.submit(np.full(4, i, dtype=np.float32), "semantic_index", 10, None) for i in range(12)]
        sparse = [mb.submit({"indices": [i, i + 1], "values": [1.0, 1.0]}, "sparse_index", 10, None) for i in range(5)]
        other_k = [mb.submit(np.full(4, 7, dtype=np.float32), "semantic_index", 20, 'doc_id == "x"')]
        return await asyncio.gather(*dense, *sparse, *other_k)

    out = _run(storm())
    assert [h[0]["id"] for h in out[:12]] == [f"d{4 * i}" for i in range(12)]
    assert [h[0]["id"] for h in out[12:17]] == [f"s{2 * i + 1}" for i in range(5)]
    assert out[17][0] == {"id": "d28", "score": 20.0}
    assert sorted(stub.calls) == sorted([("semantic_index", 10, None, 12), ("sparse_index", 10, None, 5),
                                         ("semantic_index", 20, 'doc_id == "x"', 1)])
    assert mb.batches == 3
    mb.close()


def test_micro_batcher_flushes_at_max_batch_and_propagates_failures():
    from b200rag.index_manager import _MicroBatcher
    stub = _StubManager()
    mb = _MicroBatcher(stub, max_batch=4, max_wait_s=5.0)          # only the size trigger can fire within the test

    async def four():
        return await asyncio.wait_for(asyncio.gather(*[mb.submit(np.zeros(4, np.float32), "semantic_index", 5, None)
                                                       for _ in range(8)]), timeout=2.0)

    assert len(_run(four())) == 8 and [c[3] for c in stub.calls] == [4, 4]
    mb.close()
    bad = _MicroBatcher(_StubManager(fail=True), max_batch=2, max_wait_s=5.0)

    async def failing():
        return await asyncio.gather(*[bad.submit(np.zeros(4, np.float32), "semantic_index", 5, None) for _ in range(2)],
                                    return_exceptions=True)

    res = _run(failing())
    assert all(isinstance(e, RuntimeError) and "index down" in str(e) for e in res)
    bad.close()


# ------------------------------------------------------------------------------------------------ BM25 encoder
def test_bm25_weights_follow_the_stated_formula():
    """w = idf * tf * (k1 + 1) / (tf + k1 * (1 - b + b * len / avgdl)), idf = ln(1 + (N - df + .5) / (df + .5)), fp64 in this
    order, one rounding to fp32 (b200rag/bm25.py docstring) -- against a scalar restatement."""
    import math
    from b200rag import bm25
    texts = ["the cat sat on the mat", "The dog sat", "a cat and a dog and a bird", "", "mat mat mat"]
    enc = bm25.Bm25Encoder()
    dp, ti, tf = enc.fit_transform(texts)
    w = bm25.bm25_weights(dp, ti, tf, len(enc.vocab))
    n = len(texts)
    lens = [sum(tf[dp[d]:dp[d + 1]]) for d in range(n)]
    avgdl = sum(lens) / n
    df = np.bincount(ti, minlength=len(enc.vocab))
    for d in range(n):
        for p in range(dp[d], dp[d + 1]):
            idf = math.log(1.0 + (n - float(df[ti[p]]) + 0.5) / (float(df[ti[p]]) + 0.5))
            norm = 1.2 * (1.0 - 0.75 + 0.75 * float(lens[d]) / avgdl)
            ref = np.float32(idf * float(tf[p]) * (1.2 + 1.0) / (float(tf[p]) + norm))
            assert w[p] == ref, (d, p)
    assert dp.tolist() == [0, 5, 8, 13, 13, 14] and w.dtype == np.float32


def test_bm25_encoder_tokenises_like_the_reference_mmr():
    """Tokeniser = text.lower().split() (reference retrieval.py:497), ids ascending per document, query value 1.0 per
    unique known term, unknown terms dropped; token_sets (the MMR side) shares the vocabulary."""
    from b200rag import bm25
    enc = bm25.Bm25Encoder()
    dp, ti, tf = enc.fit_transform(["Alpha beta  ALPHA\tgamma", "beta delta"])
    assert enc.vocab == {"alpha": 0, "beta": 1, "gamma": 2, "delta": 3}
    assert ti.tolist() == [0, 1, 2, 1, 3] and tf.tolist() == [2, 1, 1, 1, 1]
    q = enc.encode_sparse("delta ALPHA alpha unknown")
    assert q == {"indices": [0, 3], "values": [1.0, 1.0]}
    tp, tids = enc.token_sets(["gamma gamma beta", "new token"])
    assert tp.tolist() == [0, 2, 4] and tids.tolist() == [1, 2, 4, 5] and tids.dtype == np.int32
    assert bm25.tokenize(None) == [] and bm25.tokenize("  ") == []
