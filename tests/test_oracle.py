"""CPU tests: the oracle against the golden vectors produced by the reference's own code, and against
independent numpy arithmetic.  No GPU, no product code."""
import numpy as np
import pytest

from oracle import fusion, ref_import
from util import METHOD_NAMES, golden_lists, load_golden


def test_rrf_restatement_matches_reference_golden():
    g = load_golden()
    assert len(g["rrf"]) >= 10
    for c in g["rrf"]:
        lists, w = golden_lists(c)
        ids, sc, mask = fusion.rrf_fuse(lists, w)
        assert ids == c["out_ids"]
        assert [s.hex() for s in sc] == c["out_scores_hex"]          # bit exact fp64
        assert [sorted(METHOD_NAMES[i] for i in range(3) if m >> i & 1) for m in mask] == c["out_methods"]


def test_mmr_restatement_matches_reference_golden():
    g = load_golden()
    assert len(g["mmr"]) >= 8
    for c in g["mmr"]:
        lists, w = golden_lists(c)
        ids, sc, _ = fusion.rrf_fuse(lists, w)
        toks = [fusion.tokens(c["contents"][i]) for i in ids]
        pick = fusion.mmr_select(sc, toks, c["top_k"], c["mmr_lambda"])
        assert [ids[i] for i in pick] == c["out_ids"]
        assert [sc[i].hex() for i in pick] == c["out_scores_hex"]


def test_learned_rank_restatement_matches_reference_golden():
    for c in load_golden()["rerank"]:
        sc = [float.fromhex(h) for h in c["in_scores_hex"]]
        order, rs = fusion.learned_rank(sc, c["in_n_methods"], [0.0] * len(sc), c["top_k"])
        assert [c["in_ids"][i] for i in order] == c["out_ids"]
        assert [r.hex() for r in rs] == c["out_scores_hex"]


def test_pairwise_similarity_restatement_matches_reference_golden():
    """reference evaluation.py:327-344 (RAGEvaluator._calculate_pairwise_similarity / _calculate_diversity), golden produced by
    executing the reference (oracle/gen_golden.py)."""
    from oracle import fusion
    g = load_golden()
    assert len(g["pairwise_similarity"]) >= 6
    for case in g["pairwise_similarity"]:
        mean, pairs = fusion.pairwise_similarity([fusion.tokens(c) for c in case["contents"]])
        assert float(mean).hex() == case["mean_hex"], len(case["contents"])
        if len(case["contents"]) >= 2:
            assert float(1.0 - mean).hex() == case["diversity_hex"]


def test_reference_known_answers():
    """Known answers quoted in SURVEY.md section 8c (probe of the reference's _fuse_results)."""
    ids, sc, _ = fusion.rrf_fuse([["A", "B"], ["A", "C"]], [0.7, 0.3])
    assert ids == ["A", "B", "C"]
    assert sc[0] == (1.0 / 61) * 0.7 + (1.0 / 61) * 0.3 and sc[1] == (1.0 / 62) * 0.7 and sc[2] == (1.0 / 62) * 0.3
    # equal weights: ties keep first-seen order (semantic list before sparse list), not id order
    ids, _, _ = fusion.rrf_fuse([["z9", "a1"], ["m5", "b2"]], [0.5, 0.5])
    assert ids == ["z9", "m5", "a1", "b2"] or ids == ["z9", "a1", "m5", "b2"]
    s = dict(zip(*fusion.rrf_fuse([["z9", "a1"], ["m5", "b2"]], [0.5, 0.5])[:2]))
    assert s["z9"] == s["m5"] and ids.index("z9") < ids.index("m5")


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_restatement_against_live_reference_random():
    """Property test: random inputs through the reference's own functions vs the restatement."""
    import copy
    import random
    ref_import.load()
    from advanced_rag.retrieval import HybridRetriever, RetrievalConfig
    rng = random.Random(7)
    for trial in range(25):
        n_pool = rng.randint(1, 80)
        pool = [{"id": f"d{i}", "content": " ".join(f"t{rng.randrange(12)}" for _ in range(rng.randint(0, 7))),
                 "score": rng.random()} for i in range(n_pool)]
        pick = lambda: [copy.deepcopy(pool[i]) for i in rng.sample(range(n_pool), rng.randint(0, n_pool))]
        sem, spa, dom = pick(), pick(), (pick() if trial % 2 else [])
        lam, k = rng.choice([0.0, 0.3, 0.5, 0.7, 1.0]), rng.randint(1, 30)
        cfg = RetrievalConfig(top_k=k, enable_mmr=True, mmr_lambda=lam, dense_weight=rng.random(), sparse_weight=rng.random())
        out = HybridRetriever(index_manager=None, config=cfg)._fuse_results(
            copy.deepcopy(sem), copy.deepcopy(spa), copy.deepcopy(dom))
        lists = [[d["id"] for d in sem], [d["id"] for d in spa]] + ([[d["id"] for d in dom]] if dom else [])
        w = [cfg.dense_weight, cfg.sparse_weight, 0.2][: len(lists)]
        ids, sc, _ = fusion.rrf_fuse(lists, w)
        content = {d["id"]: d["content"] for d in pool}
        picks = fusion.mmr_select(sc, [fusion.tokens(content[i]) for i in ids], k, lam) if ids else []
        assert [ids[i] for i in picks] == [d["id"] for d in out]
        assert [sc[i].hex() for i in picks] == [float(d["score"]).hex() for d in out]


def test_conversions_against_numpy(oracle_lib):
    o = oracle_lib
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s for s in (1e-8, 1e-4, 1.0, 1e3, 7e4)])
    x = np.concatenate([x, np.array([0.0, -0.0, 65504.0, 65520.0, 1e-7, 6e-8, 2.98e-8, np.inf, -np.inf], np.float32)])
    h = o.round_f32(x, o.F16)
    with np.errstate(over="ignore"):
        assert np.array_equal(h, x.astype(np.float16).view(np.uint16))
    assert np.array_equal(o.bits_to_f32(h, o.F16).view(np.uint32), h.view(np.float16).astype(np.float32).view(np.uint32))
    # bf16: round-to-nearest-even on the fp32 bit pattern
    u = x.view(np.uint32).astype(np.uint64)
    finite = np.isfinite(x)
    rne = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    b = o.round_f32(x, o.BF16)
    assert np.array_equal(b[finite], rne[finite])
    assert np.array_equal(o.bits_to_f32(b, o.BF16).view(np.uint32) >> 16, b.astype(np.uint32))


def test_dense_topk_against_numpy_fp64(oracle_lib):
    o = oracle_lib
    rng = np.random.default_rng(1)
    for dt in (o.F16, o.BF16):
        for n, d, b, k in [(1000, 64, 7, 10), (5000, 384, 4, 40), (37, 8, 3, 50), (1, 16, 2, 1)]:
            xb = o.normalize_rows(rng.standard_normal((n, d)).astype(np.float32), dt)
            qb = o.normalize_rows(rng.standard_normal((b, d)).astype(np.float32), dt)
            s, ids = o.dense_topk(xb, qb, k, dt, id_offset=100)
            xf = o.bits_to_f32(xb, dt).astype(np.float64)
            qf = o.bits_to_f32(qb, dt).astype(np.float64)
            S = qf @ xf.T
            kk = min(k, n)
            ref = np.lexsort((np.broadcast_to(np.arange(n), S.shape), -S), axis=1)[:, :kk]
            assert np.array_equal(ids[:, :kk], ref + 100)
            assert np.allclose(s[:, :kk], np.take_along_axis(S, ref, 1), rtol=0, atol=1e-14)
            assert np.all(ids[:, kk:] == -1) and np.all(np.isneginf(s[:, kk:]))


def test_dense_ties_break_by_id(oracle_lib):
    o = oracle_lib
    rng = np.random.default_rng(2)
    base = rng.standard_normal((5, 32)).astype(np.float32)
    x = np.concatenate([base] * 40)                      # every row appears 40 times
    xb = o.normalize_rows(x, o.F16)
    qb = o.normalize_rows(base[:2], o.F16)
    s, ids = o.dense_topk(xb, qb, 60, o.F16)
    for r in range(2):
        assert list(ids[r, :40]) == [r + 5 * j for j in range(40)]     # 40 exact ties, ids ascending
        assert len(set(s[r, :40])) == 1


def test_sparse_topk_against_scipy(oracle_lib):
    import scipy.sparse as sp
    from b200rag import bm25, synth
    o = oracle_lib
    n_docs, vocab = 3000, 500
    dp, ti, tf = synth.zipf_corpus(n_docs, vocab, 3, mean_len=40)
    w = bm25.bm25_weights(dp, ti, tf, vocab)
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
    qp, qt, qv = synth.zipf_queries(16, vocab, 4, n_terms=6, skip_top=10)
    s, ids, cnt = o.sparse_topk(tp, pd, pw, n_docs, qp, qt, qv, 25)
    M = sp.csr_matrix((w.astype(np.float64), ti, dp), shape=(n_docs, vocab))
    for q in range(16):
        qvec = np.zeros(vocab)
        qvec[qt[qp[q]:qp[q + 1]]] = 1.0
        sc = M @ qvec
        touched = np.flatnonzero((M != 0) @ qvec)
        assert cnt[q] == min(25, touched.size)
        got = ids[q, :cnt[q]]
        assert set(got) <= set(touched)
        assert np.allclose(s[q, :cnt[q]], sc[got], rtol=2e-6)
        # no untouched / worse document was missed (up to fp32 rounding at the boundary)
        rest = np.setdiff1d(touched, got)
        if rest.size:
            assert sc[rest].max() <= s[q, cnt[q] - 1] * (1 + 4e-6)


def test_sparse_edge_cases(oracle_lib):
    o = oracle_lib
    # 3 docs, 4 terms; term 3 has no postings
    tp = np.array([0, 2, 3, 4, 4], np.int64)
    pd = np.array([0, 2, 1, 2], np.int32)
    pw = np.array([1.0, 1.0, 2.0, 0.5], np.float32)
    qp = np.array([0, 0, 1, 3, 4], np.int64)            # empty query, [3], [0,2], [1]
    qt = np.array([3, 0, 2, 1], np.int32)
    qv = np.ones(4, np.float32)
    s, ids, cnt = o.sparse_topk(tp, pd, pw, 3, qp, qt, qv, 2)
    assert list(cnt) == [0, 0, 2, 1]
    assert list(ids[2]) == [2, 0] and list(s[2]) == [1.5, 1.0]
    assert list(ids[3]) == [1, -1]


def test_merge_topk(oracle_lib):
    o = oracle_lib
    sc = np.array([[0.5, 0.9, 0.5, -1.0, 0.9]], np.float64)
    ids = np.array([[7, 3, 2, -1, 1]], np.int64)
    s, i = o.merge_topk(sc, ids, 6)
    assert list(i[0]) == [1, 3, 2, 7, -1, -1]


def test_pipeline_restatement_matches_reference_e2e_golden():
    """oracle/pipeline.py (the chain the GPU tests check against) == the unmodified reference HybridRetriever.retrieve
    over the in-memory index, on the committed golden queries (tests/golden/e2e_golden.json)."""
    import json
    import os
    from oracle import e2e_corpus, inmem_index, pipeline
    with open(os.path.join(os.path.dirname(__file__), "golden", "e2e_golden.json")) as f:
        golden = json.load(f)
    c = e2e_corpus.build()
    gen = inmem_index.HashEmbeddingGenerator(e2e_corpus.SEM_DIM, e2e_corpus.DOM_DIM, c["vocab"])
    corpus = pipeline.ArrayCorpus(c["semantic"], c["domain"], c["sp_ptr"], c["sp_idx"], c["sp_val"], e2e_corpus.VOCAB,
                                  c["contents"])
    assert len(golden["cases"]) == len(e2e_corpus.queries())
    n_mmr = 0
    for case, (text, kw) in zip(golden["cases"], e2e_corpus.queries()):
        assert case["query"] == text
        prof = pipeline.PROFILES_TOPK20[case["profile"]]
        n_mmr += prof["enable_mmr"]
        dom_q = gen.encode_domain(text, kw["domain"]) if kw.get("use_domain_index") else None
        ids, scores, masks = pipeline.retrieve(corpus, gen.encode_semantic(text), gen.encode_sparse(text), dom_q, **prof)
        assert [c["ids"][i] for i in ids] == case["ids"], text
        assert [float(s).hex() for s in scores] == case["scores_hex"], text
        names = ["semantic", "sparse", "domain"]
        assert [sorted(n for b, n in enumerate(names) if m >> b & 1) for m in masks] == case["methods"], text
    assert n_mmr >= 5
