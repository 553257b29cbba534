"""B200IndexManager at the boundary (reference src/advanced_rag/indexing.py:264-551, 692-695): columnar results,
GPU-evaluated metadata predicates, incremental ingest, tombstone deletes, validation before mutation, concurrent callers.
Everything is checked against the CPU oracle or the host restatement of the predicate semantics."""
import asyncio
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(coro):
    loop = asyncio.new_event_loop()
    try:
        return loop.run_until_complete(coro)
    finally:
        loop.close()


def _corpus(n, dim=64, vocab=500, seed=0):
    from b200rag import bm25, synth
    rng = np.random.default_rng(seed)
    x = synth.dense_rows(n, dim, seed)
    dp, ti, tf = synth.zipf_corpus(n, vocab, seed, mean_len=20)
    w = bm25.bm25_weights(dp, ti, tf, vocab)
    contents = [" ".join(f"w{t}" for t in ti[dp[d]: dp[d + 1]]) for d in range(n)]
    meta = [{"doc_id": f"d{d // 3}", "chunk_index": d % 3, "entropy": float(rng.random()), "redundancy": float(d % 7) / 7.0,
             "domain_density": None if d % 11 == 0 else float(d % 5) / 5.0, "timestamp": f"2024-{1 + d % 12:02d}-01T00:00:00",
             "token_count": int(dp[d + 1] - dp[d])} for d in range(n)]
    return x, (dp, ti, w), contents, meta


def _manager(n=5000, dim=64, vocab=500, **kw):
    from b200rag.index_manager import B200IndexManager
    x, csr, contents, meta = _corpus(n, dim, vocab)
    m = B200IndexManager(semantic_dim=dim, sparse_dim=vocab, domain_dim=32, device=DEV, enable_sparse=True, **kw)
    m.add([f"c{d:06d}" for d in range(n)], contents, x, csr, None, meta)
    return m, x, csr, contents, meta


def test_columnar_arrays_equal_dict_results_and_oracle(oracle_lib):
    """search_batch_arrays (the columnar plugin call) == search_batch (reference dict shape) == the oracle, for the dense and
    the sparse collection; CSR-triple and dict sparse queries are the same query."""
    from b200rag import synth
    o = oracle_lib
    m, x, (dp, ti, w), contents, meta = _manager()
    q = synth.dense_rows(33, 64, 99)
    arr = m.search_batch_arrays(q, "semantic_index", 12)
    ref_s, ref_i = o.dense_topk(o.normalize_rows(x, o.F16), o.normalize_rows(q, o.F16), 12, o.F16)
    assert np.array_equal(arr.rows, ref_i) and np.array_equal(arr.scores, ref_s) and (arr.counts == 12).all()
    lists = m.search_batch(q, "semantic_index", 12)
    assert [[h["id"] for h in hits] for hits in lists] == arr.chunk_ids() == [[f"c{r:06d}" for r in row] for row in ref_i]
    assert lists[5][3]["metadata"] == {k: meta[ref_i[5, 3]][k] for k in ("doc_id", "chunk_index", "entropy", "redundancy",
                                                                        "domain_density", "timestamp")}
    assert lists[5][3]["content"] == contents[ref_i[5, 3]] and lists[5][3]["score"] == ref_s[5, 3]
    lazy = arr.hits()
    assert lazy[7] == lists[7] and lazy[7][0] is not lazy[7][0]
    # pinned host queries and device queries are accepted as they are
    a2 = m.search_batch_arrays(torch.from_numpy(q).pin_memory(), "semantic_index", 12)
    a3 = m.search_batch_arrays(torch.from_numpy(q).to(DEV), "semantic_index", 12)
    assert np.array_equal(a2.rows, ref_i) and np.array_equal(a3.scores, ref_s)
    # sparse
    qp, qt, qv = synth.zipf_queries(20, 500, 5, n_terms=6, skip_top=10)
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, 500)
    rs, ri, rc = o.sparse_topk(tp, pd, pw, 5000, qp, qt, qv, 15)
    sa = m.search_batch_arrays((qp, qt, qv), "sparse_index", 15)
    assert np.array_equal(sa.counts, rc) and np.array_equal(sa.rows, ri)
    assert np.array_equal(sa.scores.astype(np.float32).view(np.uint32), rs.view(np.uint32))
    dicts = [{"indices": qt[qp[j]: qp[j + 1]].tolist(), "values": qv[qp[j]: qp[j + 1]].tolist()} for j in range(20)]
    sb = m.search_batch_arrays(dicts, "sparse_index", 15)
    assert np.array_equal(sb.rows, sa.rows) and np.array_equal(sb.scores, sa.scores)
    with pytest.raises(ValueError):
        m.search_batch_arrays((qp, qt[::-1].copy(), qv), "sparse_index", 5)


def test_gpu_predicate_kernel_matches_host_semantics(oracle_lib):
    """b200rag_filter_mask over the typed device columns == the numpy restatement, for every operator, field type, missing
    values, type mismatches and chunk_id; and the masked search is the exact top-k of the allowed rows."""
    from b200rag.index_manager import eval_filter_host
    o = oracle_lib
    m, x, _, _, meta = _manager()
    exprs = ['entropy >= 0.5', 'entropy < 0.25 and redundancy != 0.0', 'domain_density == 0.2', 'domain_density != 0.2',
             'chunk_index == 2', 'chunk_index >= 0.5', 'token_count <= 18 and token_count > 12', 'doc_id == "d17"',
             'doc_id != "d17"', 'doc_id == "nope"', 'doc_id != "nope"', 'timestamp >= "2024-06-01" and timestamp < "2024-09"',
             'chunk_id == "c000123"', 'chunk_id > "c004990"', 'entropy == "x"', 'doc_id == 5',
             'timestamp <= "2024-03-01T00:00:00" and entropy > 0.9 and chunk_index != 1']
    for e in exprs:
        want = eval_filter_host(m.payload, e)
        cnt, words = m._filter_words(e)
        got = np.unpackbits(words.cpu().numpy().view(np.uint8), bitorder="little")[: m.n_slots].astype(bool)
        assert np.array_equal(got, want), e
        assert cnt == int(want.sum()), e
    # cached: same tensor object the second time
    assert m._filter_words(exprs[0])[1] is m._filter_words(exprs[0])[1]
    q = np.random.default_rng(1).standard_normal((3, 64)).astype(np.float32)
    keep = eval_filter_host(m.payload, exprs[1])
    rows = np.flatnonzero(keep)
    arr = m.search_batch_arrays(q, "semantic_index", 10, filters=exprs[1])
    s, i = o.dense_topk(o.normalize_rows(x[rows], o.F16), o.normalize_rows(q, o.F16), 10, o.F16)
    assert np.array_equal(arr.rows, rows[i]) and np.array_equal(arr.scores, s)
    none = m.search_batch_arrays(q, "semantic_index", 10, filters='doc_id == "nope"')
    assert (none.counts == 0).all() and (none.rows == -1).all() and m.search_batch(q, "semantic_index", 10, 'doc_id == "nope"') == [[], [], []]
    with pytest.raises(ValueError):
        m.search_batch_arrays(q, "semantic_index", 10, filters="evil == 1")


def test_filter_at_ten_million_rows_is_milliseconds():
    """a NEW filter expression at 10M rows must cost < 5 ms once the columns are resident (it used to be a
    Python loop over every row)."""
    from b200rag import _lib, engine
    n = 10_000_000
    g = torch.Generator(device=DEV).manual_seed(0)
    ent = torch.rand(n, generator=g, device=DEV, dtype=torch.float64)
    idx = torch.randint(0, 4, (n,), generator=g, device=DEV, dtype=torch.int64)
    codes = torch.randint(0, 1000, (n,), generator=g, device=DEV, dtype=torch.int32)
    lut = (torch.arange(1000, device=DEV) % 3 == 0).to(torch.uint8)

    def terms():
        t0, t1, t2 = _lib.FilterTerm(), _lib.FilterTerm(), _lib.FilterTerm()
        t0.column, t0.kind, t0.op, t0.fvalue = ent.data_ptr(), _lib.COL_F64, _lib.OP_GE, 0.5
        t1.column, t1.kind, t1.op, t1.ivalue = idx.data_ptr(), _lib.COL_I64, _lib.OP_NE, 1
        t2.column, t2.kind, t2.op, t2.lut, t2.lut_size = codes.data_ptr(), _lib.COL_CODE, _lib.OP_GE, lut.data_ptr(), 1000
        return [t0, t1, t2]

    engine.filter_mask(terms(), n, DEV)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    words, count = engine.filter_mask(terms(), n, DEV)
    b.record()
    torch.cuda.synchronize()
    want = (ent >= 0.5) & (idx != 1) & (lut[codes.long()] != 0)
    assert int(count) == int(want.sum())
    assert torch.equal(words, engine.pack_row_mask(want))
    assert a.elapsed_time(b) < 5.0, a.elapsed_time(b)


def test_incremental_ingest_only_rebuilds_the_tail_block(oracle_lib):
    """appending documents re-blocks the new rows and the last partial block only; results equal a
    one-shot build and the oracle."""
    from b200rag import bm25, engine, synth
    o = oracle_lib
    n, vocab, bd = 10_000, 800, 1024
    dp, ti, tf = synth.zipf_corpus(n, vocab, 3, mean_len=30)
    w = bm25.bm25_weights(dp, ti, tf, vocab)
    whole = engine.SparseIndex(dp, ti, w, vocab, DEV, block_docs=bd)
    cuts = [0, 3000, 3001, 4096, 9000, n]
    inc = engine.SparseIndex(dp[:1], ti[:0], w[:0], vocab, DEV, block_docs=bd)
    built = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        before = inc.blocks_built
        inc.append(dp[a: b + 1] - dp[a], ti[dp[a]: dp[b]], w[dp[a]: dp[b]])
        built.append(inc.blocks_built - before)
    assert built == [3, 1, 2, 5, 2]                          # never the whole index again (10 blocks)
    assert inc.n_docs == n and inc.nnz == whole.nnz
    assert torch.equal(inc.blk_term_ptr, whole.blk_term_ptr) and torch.equal(inc.post_doc, whole.post_doc)
    assert torch.equal(inc.post_w, whole.post_w) and torch.equal(inc.df, whole.df)
    d2, t2, w2 = inc.to_doc_major()
    assert np.array_equal(d2.cpu().numpy(), dp) and np.array_equal(t2.cpu().numpy(), ti) and np.array_equal(w2.cpu().numpy(), w)
    qp, qt, qv = synth.zipf_queries(16, vocab, 4, n_terms=8, skip_top=10)
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
    rs, ri, rc = o.sparse_topk(tp, pd, pw, n, qp, qt, qv, 20)
    s, i, c = inc.search(qp, qt, qv, 20)
    assert np.array_equal(i.cpu().numpy(), ri) and np.array_equal(s.cpu().numpy().view(np.uint32), rs.view(np.uint32))
    with pytest.raises(ValueError):
        inc.append(np.asarray([0, 2]), np.asarray([5, 5]), np.asarray([1.0, 1.0], np.float32))      # duplicate (doc, term)


def test_manager_add_in_batches_tombstones_and_compaction(oracle_lib, tmp_path):
    from b200rag.index_manager import B200IndexManager
    from b200rag import synth
    o = oracle_lib
    n, dim, vocab = 6000, 64, 500
    x, (dp, ti, w), contents, meta = _corpus(n, dim, vocab)
    ids = [f"c{d:06d}" for d in range(n)]
    m = B200IndexManager(semantic_dim=dim, sparse_dim=vocab, domain_dim=32, device=DEV, enable_sparse=True, sparse_block_docs=1024)
    for a, b in ((0, 2500), (2500, 2501), (2501, n)):
        rows = [{"indices": ti[dp[d]: dp[d + 1]].tolist(), "values": w[dp[d]: dp[d + 1]].tolist()} for d in range(a, b)]
        m.add(ids[a:b], contents[a:b], x[a:b], rows, None, meta[a:b])
    assert m.num_rows == n and m._sparse.n_docs == n
    qp, qt, qv = synth.zipf_queries(8, vocab, 5, n_terms=6, skip_top=10)
    q = synth.dense_rows(4, dim, 77)
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
    # delete ~30 % (tombstones: row ids stay, the kernels skip the rows), then > 50 % (compaction)
    for expr, keep_fn in (("redundancy >= 0.5", lambda d: not (float(d % 7) / 7.0 >= 0.5)),
                          ("chunk_index == 0", lambda d: not (float(d % 7) / 7.0 >= 0.5) and d % 3 != 0)):
        before = m.num_rows
        n_del = _run(m.delete_by_filter("semantic_index", expr))
        keep = np.asarray([keep_fn(d) for d in range(n)])
        assert n_del == before - int(keep.sum()) and m.num_rows == int(keep.sum())
        rows = np.flatnonzero(keep)
        s, i = o.dense_topk(o.normalize_rows(x[rows], o.F16), o.normalize_rows(q, o.F16), 9, o.F16)
        hits = m.search_batch(q, "semantic_index", 9)
        assert [[h["id"] for h in hh] for hh in hits] == [[ids[rows[j]] for j in row] for row in i]
        assert [[h["score"] for h in hh] for hh in hits] == [list(row) for row in s]
        rs, ri, rc = o.sparse_topk(tp, pd, pw, n, qp, qt, qv, n)
        sp = m.search_batch((qp, qt, qv), "sparse_index", 7)
        for j in range(8):
            want = [int(ri[j, t]) for t in range(rc[j]) if keep[ri[j, t]]][:7]
            assert [h["id"] for h in sp[j]] == [ids[r] for r in want], (expr, j)
    assert m.n_slots == m.num_rows < n // 2                   # the second delete crossed 50 % dead rows: compacted
    # a filter combined with deleted rows, after compaction
    hits = m.search_batch(q[:1], "semantic_index", 5, filters='doc_id == "d100"')
    assert {h["metadata"]["doc_id"] for h in hits[0]} <= {"d100"}
    # checkpoint: nothing pickled, reload answers identically
    path = str(tmp_path / "idx.b200rag")
    m.save(path)
    with np.load(path, allow_pickle=False) as z:
        assert "sem" in z.files and z["sem"].dtype == np.int16
    m2 = B200IndexManager.load(path, device=DEV, sparse_block_docs=1024)
    a1, a2 = m.search_batch_arrays(q, "semantic_index", 9), m2.search_batch_arrays(q, "semantic_index", 9)
    assert np.array_equal(a1.rows, a2.rows) and np.array_equal(a1.scores, a2.scores) and a1.chunk_ids() == a2.chunk_ids()
    s1, s2 = m.search_batch((qp, qt, qv), "sparse_index", 7), m2.search_batch((qp, qt, qv), "sparse_index", 7)
    assert s1 == s2
    assert torch.equal(m.token_sets()[1], m2.token_sets()[1]) and m.token_sets()[2] == m2.token_sets()[2]


def test_add_validates_before_mutating():
    """a bad row must leave every index untouched (it used to leave the dense index n rows ahead)."""
    m, x, _, contents, meta = _manager(n=300)
    state = (m._sem.n, m._sparse.n_docs, len(m.payload), m._tok_ptr.n, m._live.n, len(m._tok_vocab))
    ok_rows = [{"indices": [1, 2], "values": [1.0, 2.0]}] * 3
    bad = [
        dict(sparse=[{"indices": [1], "values": [1.0]}, {"indices": [10**6], "values": [1.0]}, None]),      # index out of range
        dict(sparse=[{"indices": [1, 2], "values": [1.0]}, None, None]),                                   # ragged entry
        dict(sparse=ok_rows, metadata=[{"entropy": "high"}, {}, {}]),                                       # untypable value
        dict(sparse=ok_rows, metadata=[{}]),                                                               # wrong length
        dict(sparse=ok_rows, domain=np.zeros((3, 32), np.float32)),                                         # domain for some rows only
        dict(sparse=ok_rows, semantic=np.zeros((3, 63), np.float32)),                                       # wrong dim
    ]
    for kw in bad:
        args = dict(ids=["n0", "n1", "n2"], contents=["a b", "c", ""], semantic=np.ones((3, 64), np.float32), sparse=None,
                    domain=None, metadata=None)
        args.update(kw)
        with pytest.raises(ValueError):
            m.add(**args)
        assert (m._sem.n, m._sparse.n_docs, len(m.payload), m._tok_ptr.n, m._live.n, len(m._tok_vocab)) == state, kw
    m.add(["n0", "n1", "n2"], ["a b", "c", ""], np.ones((3, 64), np.float32), ok_rows, None, None)
    assert m.num_rows == 303 and m._sparse.n_docs == 303
    hit = _run(m.search(np.ones(64, np.float32), "semantic_index", top_k=3))
    assert {h["id"] for h in hit} == {"n0", "n1", "n2"} and hit[0]["metadata"]["doc_id"] is None


def test_concurrent_callers_get_their_own_results(oracle_lib):
    """two host threads searching at once (the micro-batcher's worker and a direct caller) used to share one
    scratch buffer.  Each thread's results must equal the oracle's for ITS queries."""
    from b200rag import synth
    o = oracle_lib
    m, x, _, _, _ = _manager(n=20000)
    xb = o.normalize_rows(x, o.F16)
    errors = []

    def worker(seed):
        try:
            for it in range(6):
                q = synth.dense_rows(64, 64, 1000 * seed + it)
                arr = m.search_batch_arrays(q, "semantic_index", 10)
                s, i = o.dense_topk(xb, o.normalize_rows(q, o.F16), 10, o.F16)
                if not (np.array_equal(arr.rows, i) and np.array_equal(arr.scores, s)):
                    errors.append((seed, it))
        except Exception as e:  # noqa: BLE001
            errors.append((seed, repr(e)))

    ts = [threading.Thread(target=worker, args=(s,)) for s in (1, 2, 3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("vocab", [600_000, 3_000_000])
def test_mmr_large_vocabularies_match_oracle(vocab):
    """vocabularies beyond 589,824 tokens made the general MMR kernel write past its shared memory, beyond
    ~1.8M it refused the batch.  600K: bitset in shared memory, no token cache; 3M: bitset in the caller's workspace."""
    from b200rag import engine
    from oracle import fusion
    rng = np.random.default_rng(vocab)
    b, n_max, k = 3, 300, 25
    n_docs = 2000
    lens = rng.integers(0, 60, size=n_docs)
    dp = np.zeros(n_docs + 1, np.int64)
    np.cumsum(lens, out=dp[1:])
    # a shared pool of frequent tokens (so that documents overlap) plus rare tokens from the whole id range
    pool = rng.choice(vocab, size=400, replace=False)
    toks = []
    for d in range(n_docs):
        k_common = int(lens[d] * 0.7)
        s = set(rng.choice(pool, size=k_common, replace=False).tolist()) if k_common else set()
        while len(s) < lens[d]:
            s.add(int(rng.integers(vocab)))
        toks.append(np.sort(np.fromiter(s, dtype=np.int64)))
    ti = np.concatenate(toks).astype(np.int32)
    assert int(ti.max()) > 589_824
    cand = np.stack([rng.choice(n_docs, size=n_max, replace=False) for _ in range(b)]).astype(np.int32)
    n = np.asarray([n_max, 17, 211], np.int32)
    rel = np.sort(rng.random((b, n_max)) * 0.016, axis=1)[:, ::-1].copy()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    lam, ks = [0.7, 0.5, 0.8], [k, k, 9]
    picks, pn = engine.mmr_select(t(cand), t(rel), t(n), t(dp), t(ti), vocab, t(np.asarray(lam)), t(np.asarray(ks, np.int32)), k)
    for q in range(b):
        sets = [frozenset(ti[dp[d]: dp[d + 1]].tolist()) for d in cand[q, : n[q]]]
        ref = fusion.mmr_select(list(rel[q, : n[q]]), sets, ks[q], lam[q])
        assert picks[q, : int(pn[q])].cpu().tolist() == ref, q


def test_graph_replay_returns_the_same_results_as_eager_search(oracle_lib):
    """search_batch_arrays replays a captured CUDA graph from the third call of a (collection, batch, k, filter) shape on: every
    replay must equal the oracle for ITS queries, with and without a filter, and an insert must drop the stale graphs."""
    from b200rag import synth
    from b200rag.index_manager import eval_filter_host
    o = oracle_lib
    m, x, _, contents, meta = _manager(n=30000)
    assert m.use_graphs
    xb = o.normalize_rows(x, o.F16)
    expr = "entropy >= 0.5"
    rows = np.flatnonzero(eval_filter_host(m.payload, expr))
    for it in range(7):
        q = synth.dense_rows(40, 64, 500 + it)
        qb = o.normalize_rows(q, o.F16)
        arr = m.search_batch_arrays(torch.from_numpy(q).pin_memory(), "semantic_index", 10)
        s, i = o.dense_topk(xb, qb, 10, o.F16)
        assert np.array_equal(arr.rows, i) and np.array_equal(arr.scores, s), it
        fa = m.search_batch_arrays(q, "semantic_index", 10, filters=expr)
        s, i = o.dense_topk(xb[rows], qb, 10, o.F16)
        assert np.array_equal(fa.rows, rows[i]) and np.array_equal(fa.scores, s), it
        one = m.search_batch_arrays(q[0], "semantic_index", 5)
        s1, i1 = o.dense_topk(xb, qb[:1], 5, o.F16)
        assert np.array_equal(one.rows, i1) and np.array_equal(one.scores, s1), it
    assert len(m._graphs) == 3                                       # the three shapes above
    m.add(["new0"], ["fresh text"], 10.0 * x[:1], None, None, None)
    assert len(m._graphs) == 0
    for it in range(4):
        arr = m.search_batch_arrays(x[:1], "semantic_index", 2)
        assert arr.chunk_ids()[0] == ["c000000", "new0"] and arr.scores[0, 0] == arr.scores[0, 1]
