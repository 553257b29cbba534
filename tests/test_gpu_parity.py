"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on identical seeded inputs.
Integer / id results must be bit exact; fp64 / fp32 scores must be bit exact too (same canonical arithmetic)."""
import numpy as np
import pytest
import torch

from util import METHOD_NAMES, golden_lists, id_map, load_golden, to_bits_view

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def eng():
    from b200rag import engine
    sm, major, minor = engine.device_info()
    assert major >= 10, f"expected an sm_100 device, got sm_{major}{minor}"
    return engine


def _codes(o, eng):
    return [(o.F16, eng.F16, torch.float16), (o.BF16, eng.BF16, torch.bfloat16)]


# ------------------------------------------------------------------------------------------------ row preparation
def test_prepare_rows_bit_exact(eng, oracle_lib):
    o = oracle_lib
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 96)).astype(np.float32)
    x[3] = 0.0                                  # all-zero row stays zero
    x[4] *= 1e-6
    x[5] *= 3e4
    for oc, ec, td in _codes(o, eng):
        for normalize in (True, False):
            got = eng.prepare_rows(torch.from_numpy(x).to(DEV), ec, normalize)
            assert got.dtype == td
            ref = o.normalize_rows(x, oc) if normalize else o.round_f32(x, oc)
            assert np.array_equal(to_bits_view(got), ref), (oc, normalize)


# ------------------------------------------------------------------------------------------------ dense, exact mode
@pytest.mark.parametrize("n,d,b,k", [(1, 8, 1, 1), (37, 16, 3, 50), (1000, 64, 5, 10), (4096, 384, 9, 40),
                                     (20000, 768, 4, 100), (100000, 384, 8, 40)])
def test_dense_exact_mode_matches_oracle(eng, oracle_lib, n, d, b, k):
    o = oracle_lib
    rng = np.random.default_rng(n + d)
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((b, d)).astype(np.float32)
    for oc, ec, td in _codes(o, eng):
        xb, qb = o.normalize_rows(x, oc), o.normalize_rows(q, oc)
        ref_s, ref_i = o.dense_topk(xb, qb, k, oc, id_offset=1000)
        idx = eng.DenseIndex(d, ec, "COSINE", DEV, id_offset=1000)
        idx.add(torch.from_numpy(x))
        assert np.array_equal(to_bits_view(idx.rows), xb)
        s, i, f = idx.search(torch.from_numpy(q), k, mode=eng.DENSE_EXACT)
        assert np.array_equal(i.cpu().numpy(), ref_i)
        assert np.array_equal(s.cpu().numpy().view(np.uint64), ref_s.view(np.uint64))   # bit exact fp64
        assert int(f.sum()) == 0


def test_dense_exact_ties_and_ip_metric(eng, oracle_lib):
    o = oracle_lib
    rng = np.random.default_rng(5)
    base = rng.standard_normal((7, 64)).astype(np.float32)
    x = np.concatenate([base] * 300)                       # 300 exact copies of each row -> massive ties
    q = base[:3] * 2.5
    xb, qb = o.round_f32(x, o.F16), o.round_f32(q, o.F16)
    ref_s, ref_i = o.dense_topk(xb, qb, 100, o.F16)
    idx = eng.DenseIndex(64, "f16", "IP", DEV)
    idx.add(torch.from_numpy(x))
    s, i, _ = idx.search(torch.from_numpy(q), 100, mode=eng.DENSE_EXACT)
    assert np.array_equal(i.cpu().numpy(), ref_i)
    assert np.array_equal(s.cpu().numpy(), ref_s)
    assert list(ref_i[0, :5]) == [0, 7, 14, 21, 28]        # ids ascending inside the tie


def test_dense_empty_corpus_and_empty_batch(eng):
    idx = eng.DenseIndex(16, "f16", "COSINE", DEV)
    s, i, _ = idx.search(torch.zeros(2, 16), 3, mode=eng.DENSE_EXACT)
    assert (i.cpu() == -1).all() and torch.isneginf(s.cpu()).all()
    idx.add(torch.randn(10, 16))
    s, i, _ = idx.search(torch.zeros(0, 16), 3, mode=eng.DENSE_EXACT)
    assert s.shape == (0, 3)


# ------------------------------------------------------------------------------------------------ merge
def test_merge_topk_matches_oracle(eng, oracle_lib):
    o = oracle_lib
    rng = np.random.default_rng(9)
    for b, m, k in [(1, 5, 6), (17, 800, 100), (4, 3000, 1000), (3, 1, 1)]:
        sc = np.round(rng.standard_normal((b, m)), 1)                # coarse -> many ties
        ids = np.stack([rng.permutation(m * 3)[:m] for _ in range(b)]).astype(np.int64) + (1 << 33)
        ids[rng.random((b, m)) < 0.1] = -1
        ref_s, ref_i = o.merge_topk(sc, ids, k)
        s, i = eng.merge_topk(torch.from_numpy(sc).to(DEV), torch.from_numpy(ids).to(DEV), k)
        assert np.array_equal(i.cpu().numpy(), ref_i)
        assert np.array_equal(s.cpu().numpy(), ref_s)


# ------------------------------------------------------------------------------------------------ sparse
def _sparse_case(n_docs, vocab, n_q, seed, mean_len=60, n_terms=8, skip_top=20):
    from b200rag import bm25, synth
    dp, ti, tf = synth.zipf_corpus(n_docs, vocab, seed, mean_len=mean_len)
    w = bm25.bm25_weights(dp, ti, tf, vocab)
    qp, qt, qv = synth.zipf_queries(n_q, vocab, seed + 1, n_terms=n_terms, skip_top=skip_top)
    return dp, ti, w, qp, qt, qv


@pytest.mark.parametrize("mask_kernel", [False, True])
@pytest.mark.parametrize("n_docs,vocab,n_q,k,block", [(3000, 500, 16, 25, 1024), (100000, 30000, 32, 40, 32768),
                                                       (70000, 2000, 8, 1000, 32768), (50, 40, 4, 10, 32)])
def test_sparse_matches_oracle(eng, oracle_lib, n_docs, vocab, n_q, k, block, mask_kernel):
    from b200rag import _lib, synth
    o = oracle_lib
    _lib.set_option("sparse_flags", 11 if mask_kernel else -1)       # 11: the experimental term-mask kernel (sparse_mask.cu)
    dp, ti, w, qp, qt, qv = _sparse_case(n_docs, vocab, n_q, seed=n_docs)
    qv = (qv * np.linspace(0.5, 2.0, qv.size)).astype(np.float32)       # general sparse IP, not only 1.0
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
    ref_s, ref_i, ref_c = o.sparse_topk(tp, pd, pw, n_docs, qp, qt, qv, k, id_offset=7)
    idx = eng.SparseIndex(dp, ti, w, vocab, DEV, block_docs=block, id_offset=7)
    try:
        s, i, c = idx.search(qp, qt, qv, k)
    finally:
        _lib.set_option("sparse_flags", -1)
    assert np.array_equal(c.cpu().numpy(), ref_c)
    assert np.array_equal(i.cpu().numpy(), ref_i)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), ref_s.view(np.uint32))      # bit exact fp32


@pytest.mark.parametrize("signed", [False, True])
def test_sparse_long_walk_both_collect_modes(eng, oracle_lib, signed):
    """Enough queries that a CTA walks ALL blocks of its query (no slicing): after the first blocks the running k-th best
    is positive and the collect switches from the touched-bitmap walk to the dense threshold scan.  With signed query
    values many queries keep a non-positive threshold and stay on the bitmap path.  Also with a document filter."""
    from b200rag import synth
    o = oracle_lib
    n_docs, vocab, n_q, k = 40000, 1500, 320, 12
    dp, ti, w, qp, qt, qv = _sparse_case(n_docs, vocab, n_q, seed=77, mean_len=40)
    rng = np.random.default_rng(5)
    qv = (qv * rng.uniform(0.25, 2.0, qv.size)).astype(np.float32)
    if signed:
        qv = (qv * np.where(rng.random(qv.size) < 0.6, -1.0, 1.0)).astype(np.float32)
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
    idx = eng.SparseIndex(dp, ti, w, vocab, DEV, block_docs=2048)
    ref_s, ref_i, ref_c = o.sparse_topk(tp, pd, pw, n_docs, qp, qt, qv, k)
    s, i, c = idx.search(qp, qt, qv, k)
    assert np.array_equal(c.cpu().numpy(), ref_c)
    assert np.array_equal(i.cpu().numpy(), ref_i)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), ref_s.view(np.uint32))
    # the same walk with a filter: exact top-k of the allowed documents
    allowed = rng.random(n_docs) < 0.3
    rs, ri, rc = o.sparse_topk(tp, pd, pw, n_docs, qp[:41], qt, qv, n_docs)
    s, i, c = idx.search(qp, qt, qv, k, doc_mask=eng.pack_row_mask(torch.from_numpy(allowed).to(DEV)))
    s, i, c = s.cpu().numpy(), i.cpu().numpy(), c.cpu().numpy()
    for q in range(40):
        keep = [j for j in range(rc[q]) if allowed[ri[q, j]]][:k]
        assert c[q] == len(keep), q
        assert list(i[q, : c[q]]) == [int(ri[q, j]) for j in keep], q
        assert list(s[q, : c[q]]) == [rs[q, j] for j in keep], q


def test_sparse_full_size_collect_paths_agree(eng):
    """BASELINE config 4's sparse side at full size (1M documents, 100K-term Zipf vocabulary, 256 queries of 8 terms, top-500;
    too large for the CPU oracle inside a test): the product kernel's collect paths, the experimental term-mask kernel, and
    different slicings (1, 5 slices per query) must return bit-identical lists; the lists are sorted by (score desc, id asc);
    and the scores agree with an fp64 recomputation from the CSR.  (The oracle itself checks the same walk at 16384-document
    blocks x 9 blocks in test_sparse_16384_blocks_long_walk_matches_oracle.)"""
    from b200rag import _lib
    from b200rag import bm25, synth
    n_docs, vocab, n_q, k = 1_000_000, 100_000, 256, 500
    doc_ptr, term_ids, tf = synth.zipf_corpus_device(n_docs, vocab, 0, DEV)
    w = bm25.bm25_weights_device(doc_ptr, term_ids, tf, vocab)
    idx = eng.SparseIndex(doc_ptr, term_ids, w, vocab, DEV)
    qp, qt, qv = synth.zipf_queries(n_q, vocab, 100, n_terms=8, skip_top=100)
    qv = (qv * np.random.default_rng(1).uniform(0.5, 2.0, qv.size)).astype(np.float32)
    got = {}
    try:
        for flags in ("0", "1", "2", "3"):
            # 0: bitmap walk + one candidate per thread and round, 11: the experimental term-mask kernel (sparse_mask.cu)
            _lib.set_option("sparse_flags", {"0": -1, "1": 0, "2": 11, "3": 11}[flags])
            _lib.set_option("sparse_slices", {"0": -1, "1": -1, "2": 1, "3": 5}[flags])
            s, i, c = idx.search(qp, qt, qv, k)
            torch.cuda.synchronize()
            got[flags] = (s.cpu().numpy(), i.cpu().numpy(), c.cpu().numpy())
    finally:
        _lib.set_option("sparse_flags", -1)
        _lib.set_option("sparse_slices", -1)
    s0, i0, c0 = got["0"]
    for flags in ("1", "2", "3"):
        s, i, c = got[flags]
        assert np.array_equal(c, c0) and np.array_equal(i, i0), flags
        assert np.array_equal(s.view(np.uint32), s0.view(np.uint32)), flags
    assert (c0 == k).all()
    assert (s0[:, :-1] >= s0[:, 1:]).all()
    ties = s0[:, :-1] == s0[:, 1:]
    assert (i0[:, :-1][ties] < i0[:, 1:][ties]).all()
    # fp64 scores of the first queries from the doc-major CSR
    d_of = torch.repeat_interleave(torch.arange(n_docs, device=DEV), doc_ptr[1:] - doc_ptr[:-1])
    for q in range(4):
        wq = torch.zeros(vocab, dtype=torch.float64, device=DEV)
        wq[torch.from_numpy(qt[qp[q]:qp[q + 1]].astype(np.int64)).to(DEV)] = torch.from_numpy(qv[qp[q]:qp[q + 1]].astype(np.float64)).to(DEV)
        full = torch.zeros(n_docs, dtype=torch.float64, device=DEV).index_add_(0, d_of, wq[term_ids] * w.double())
        ref_s, _ = torch.topk(full, k)
        assert np.allclose(full[torch.from_numpy(i0[q]).to(DEV)].cpu().numpy(), s0[q], rtol=1e-5)
        assert np.allclose(ref_s.cpu().numpy(), s0[q], rtol=1e-5)


@pytest.mark.parametrize("slices,k", [(1, 40), (-1, 40), (3, 500)])
def test_sparse_16384_blocks_long_walk_matches_oracle(eng, oracle_lib, slices, k):
    """The production block size: 16384-document blocks x 10 blocks, walked by one CTA per query (slices = 1), by the
    library's own slicing and by three slices that share thresholds; top-40 (config 1) and top-500 (config 4) -- bit exact
    against the oracle, with and without a document filter."""
    from b200rag import _lib, synth
    o = oracle_lib
    n_docs, vocab, n_q = 150_000, 20_000, 96
    dp, ti, w, qp, qt, qv = _sparse_case(n_docs, vocab, n_q, seed=31, mean_len=100, skip_top=50)
    qv = (qv * np.random.default_rng(2).uniform(0.5, 2.0, qv.size)).astype(np.float32)
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
    ref_s, ref_i, ref_c = o.sparse_topk(tp, pd, pw, n_docs, qp, qt, qv, k)
    idx = eng.SparseIndex(dp, ti, w, vocab, DEV, block_docs=16384)
    assert idx.n_blocks == 10
    _lib.set_option("sparse_slices", slices)
    try:
        s, i, c = idx.search(qp, qt, qv, k)
        allowed = np.random.default_rng(3).random(n_docs) < 0.5
        sm, im, cm = idx.search(qp, qt, qv, k, doc_mask=eng.pack_row_mask(torch.from_numpy(allowed).to(DEV)))
    finally:
        _lib.set_option("sparse_slices", -1)
    assert np.array_equal(c.cpu().numpy(), ref_c)
    assert np.array_equal(i.cpu().numpy(), ref_i)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), ref_s.view(np.uint32))
    rs, ri, rc = o.sparse_topk(tp, pd, pw, n_docs, qp[:9], qt, qv, n_docs)
    sm, im, cm = sm.cpu().numpy(), im.cpu().numpy(), cm.cpu().numpy()
    for q in range(8):
        keep = [j for j in range(rc[q]) if allowed[ri[q, j]]][:k]
        assert cm[q] == len(keep), q
        assert list(im[q, : cm[q]]) == [int(ri[q, j]) for j in keep], q
        assert list(sm[q, : cm[q]]) == [rs[q, j] for j in keep], q


@pytest.mark.parametrize("block", [1024, 16384])
def test_sparse_many_terms_per_query(eng, oracle_lib, block):
    """Queries with more than eight terms (several term groups per block) and with very few, enough queries that a CTA walks
    all blocks of its query."""
    from b200rag import bm25, synth
    o = oracle_lib
    n_docs, vocab, k = 30000, 800, 17
    dp, ti, tf = synth.zipf_corpus(n_docs, vocab, 21, mean_len=30)
    w = bm25.bm25_weights(dp, ti, tf, vocab)
    rng = np.random.default_rng(8)
    terms, ptr = [], [0]
    for q in range(310):
        nt = int(rng.choice([1, 2, 8, 9, 16, 20, 27]))
        terms.append(np.sort(rng.choice(vocab, size=nt, replace=False)))
        ptr.append(ptr[-1] + nt)
    qp = np.asarray(ptr, np.int64)
    qt = np.concatenate(terms).astype(np.int32)
    qv = rng.uniform(0.2, 2.0, qt.size).astype(np.float32)
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
    ref_s, ref_i, ref_c = o.sparse_topk(tp, pd, pw, n_docs, qp, qt, qv, k)
    idx = eng.SparseIndex(dp, ti, w, vocab, DEV, block_docs=block)
    s, i, c = idx.search(qp, qt, qv, k)
    assert np.array_equal(c.cpu().numpy(), ref_c)
    assert np.array_equal(i.cpu().numpy(), ref_i)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), ref_s.view(np.uint32))


def test_sparse_edge_cases(eng, oracle_lib):
    o = oracle_lib
    # doc-major CSR: doc0 {t0:1}, doc1 {t1:2}, doc2 {t0:1, t2:.5}; term 3 unused
    dp = np.array([0, 1, 2, 4], np.int64)
    ti = np.array([0, 1, 0, 2], np.int64)
    w = np.array([1.0, 2.0, 1.0, 0.5], np.float32)
    idx = eng.SparseIndex(dp, ti, w, 4, DEV, block_docs=32)
    qp = np.array([0, 0, 1, 3, 4], np.int64)
    qt = np.array([3, 0, 2, 1], np.int32)
    s, i, c = idx.search(qp, qt, np.ones(4, np.float32), 2)
    assert c.cpu().tolist() == [0, 0, 2, 1]
    assert i.cpu().tolist()[2] == [2, 0] and s.cpu().tolist()[2] == [1.5, 1.0]
    assert i.cpu().tolist()[3] == [1, -1]
    assert i.cpu().tolist()[0] == [-1, -1]


# ------------------------------------------------------------------------------------------------ RRF / MMR
def _run_rrf(eng, lists_per_query, weights_per_query, k_max):
    nl = max(len(l) for l in lists_per_query)
    b = len(lists_per_query)
    ids = np.full((nl, b, k_max), -1, np.int64)
    lens = np.zeros((nl, b), np.int32)
    wts = np.zeros((b, nl), np.float64)
    for q, (lists, w) in enumerate(zip(lists_per_query, weights_per_query)):
        for l, lst in enumerate(lists):
            ids[l, q, :len(lst)] = lst
            lens[l, q] = len(lst)
            wts[q, l] = w[l]
    out = eng.rrf_fuse(torch.from_numpy(ids).to(DEV), torch.from_numpy(lens).to(DEV), torch.from_numpy(wts).to(DEV))
    return out


def test_rrf_matches_reference_golden(eng):
    g = load_golden()
    for c in g["rrf"]:
        lists, w = golden_lists(c)
        if len(lists) < 3:
            w = w + [0.2]
            lists = lists + [[]]
        fwd, back = id_map(lists)
        ilists = [[fwd[x] for x in l] for l in lists]
        k_max = max(1, max(len(l) for l in lists))
        out = _run_rrf(eng, [ilists], [w], k_max)
        n = int(out.n[0])
        assert [back[int(x)] for x in out.ids[0, :n].cpu()] == c["out_ids"]
        assert [float(x).hex() for x in out.scores[0, :n].cpu()] == c["out_scores_hex"]
        masks = out.mask[0, :n].cpu().tolist()
        assert [sorted(METHOD_NAMES[i] for i in range(3) if m >> i & 1) for m in masks] == c["out_methods"]


def test_rrf_batch_matches_oracle_random(eng):
    from oracle import fusion
    rng = np.random.default_rng(11)
    b, k_max = 64, 500
    lists_pq, w_pq = [], []
    for q in range(b):
        pool = rng.permutation(2000)[: rng.integers(1, 1200)] + (1 << 20)
        lists = [list(rng.permutation(pool)[: rng.integers(0, min(k_max, pool.size) + 1)]) for _ in range(2)]
        lists_pq.append(lists)
        w_pq.append([float(rng.choice([0.7, 0.5, 0.1])), float(rng.choice([0.3, 0.5, 0.9]))])
    out = _run_rrf(eng, lists_pq, w_pq, k_max)
    for q in range(b):
        ids, sc, mask = fusion.rrf_fuse(lists_pq[q], w_pq[q])
        n = int(out.n[q])
        assert n == len(ids)
        assert out.ids[q, :n].cpu().tolist() == [int(x) for x in ids]
        assert out.scores[q, :n].cpu().numpy().tobytes() == np.asarray(sc, np.float64).tobytes()
        assert out.mask[q, :n].cpu().tolist() == mask
        assert (out.ids[q, n:].cpu() == -1).all()


def _mmr_inputs(contents_per_query, rel_per_query, lam, k, n_max):
    from b200rag.bm25 import Bm25Encoder
    enc = Bm25Encoder()
    b = len(contents_per_query)
    all_texts = [t for cs in contents_per_query for t in cs]
    ptr, tok = enc.token_sets(all_texts)
    cand_doc = np.zeros((b, n_max), np.int32)
    rel = np.zeros((b, n_max), np.float64)
    n = np.zeros(b, np.int32)
    base = 0
    for q, cs in enumerate(contents_per_query):
        cand_doc[q, :len(cs)] = np.arange(base, base + len(cs))
        rel[q, :len(cs)] = rel_per_query[q]
        n[q] = len(cs)
        base += len(cs)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return (t(cand_doc), t(rel), t(n), t(ptr), t(tok if tok.size else np.zeros(1, np.int32)), max(1, len(enc.vocab)),
            t(np.asarray(lam, np.float64)), t(np.asarray(k, np.int32)))


def test_mmr_matches_reference_golden(eng):
    from oracle import fusion
    g = load_golden()
    for c in g["mmr"]:
        lists, w = golden_lists(c)
        ids, sc, _ = fusion.rrf_fuse(lists, w)                 # restatement already pinned to the reference
        contents = [c["contents"][i] for i in ids]
        n_max = max(1, len(ids))
        args = _mmr_inputs([contents], [sc], [c["mmr_lambda"]], [c["top_k"]], n_max)
        picks, n = eng.mmr_select(*args, k_max=max(1, c["top_k"]))
        got = [ids[int(p)] for p in picks[0, : int(n[0])].cpu()]
        assert got == c["out_ids"], (c["top_k"], c["mmr_lambda"])


def test_mmr_batch_matches_oracle_random(eng):
    from oracle import fusion
    rng = np.random.default_rng(13)
    b, n_max, k_max = 32, 200, 40
    contents_pq, rel_pq, lam, k = [], [], [], []
    for q in range(b):
        n = int(rng.integers(1, n_max + 1))
        vocab = int(rng.choice([8, 40, 400]))
        contents_pq.append([" ".join(f"w{rng.integers(vocab)}" for _ in range(rng.integers(0, 30))) for _ in range(n)])
        rel_pq.append(np.sort(rng.random(n) * 0.02)[::-1].copy())
        lam.append(float(rng.choice([0.0, 0.5, 0.7, 0.8, 1.0])))
        k.append(int(rng.integers(1, k_max + 1)))
    args = _mmr_inputs(contents_pq, rel_pq, lam, k, n_max)
    picks, n = eng.mmr_select(*args, k_max=k_max)
    for q in range(b):
        ref = fusion.mmr_select(list(rel_pq[q]), [fusion.tokens(t) for t in contents_pq[q]], k[q], lam[q])
        assert picks[q, : int(n[q])].cpu().tolist() == ref, q


@pytest.mark.parametrize("mmr_path", [-1, 3, 2, 1])
def test_mmr_config4_scale_matches_oracle(eng, mmr_path):
    """BASELINE config-4 shape for the diversification: ~1000 candidates with ~90 unique tokens each out of a 100K-term
    Zipf vocabulary (token ids beyond 65535 exercise the 16-bit token cache), lambda 0.7, k = 100.  All three MMR kernels
    (-1: the product dispatch; 3: inverted candidate lists ONLY -- a query it could not hold would come back with n = -2;
    2: thread-per-candidate bitset probes; 1: warp-per-candidate general path)
    against the oracle restatement of the reference."""
    from b200rag import synth
    from b200rag import _lib
    from oracle import fusion
    _lib.set_option("mmr_path", mmr_path)
    vocab, n_docs, b, n_max, k = 100_000, 4000, 3, 1000, 100
    dp, ti, _ = synth.zipf_corpus(n_docs, vocab, 3)
    rng = np.random.default_rng(17)
    cand = np.stack([rng.choice(n_docs, size=n_max, replace=False) for _ in range(b)]).astype(np.int32)
    n = np.asarray([n_max, 997, 640], np.int32)
    rel = np.sort(rng.random((b, n_max)) * 0.016, axis=1)[:, ::-1].copy()
    rel[1, 5] = rel[1, 4]                                   # an exact relevance tie: the earlier candidate must win
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    try:
        picks, pn = eng.mmr_select(t(cand), t(rel), t(n), t(dp), t(ti.astype(np.int32)), vocab,
                                   t(np.asarray([0.7, 0.5, 0.8])), t(np.asarray([k, k, 37], np.int32)), k)
    finally:
        _lib.set_option("mmr_path", -1)
    assert int(ti.max()) > 65535
    for q in range(b):
        sets = [frozenset(ti[dp[d]: dp[d + 1]].tolist()) for d in cand[q, : n[q]]]
        ref = fusion.mmr_select(list(rel[q, : n[q]]), sets, [k, k, 37][q], [0.7, 0.5, 0.8][q])
        assert picks[q, : int(pn[q])].cpu().tolist() == ref, q


@pytest.mark.parametrize("vocab", [700_000, 2_500_000])
@pytest.mark.parametrize("mmr_path", [-1, 1])
def test_mmr_large_vocabulary_matches_oracle(eng, vocab, mmr_path):
    """Corpus-wide whitespace vocabularies of 1M+ chunk corpora: above 589 824 tokens the general kernel runs without its
    16-bit token cache (round 1 wrote past its shared memory there), above ~1.8M the per-pick bitset moves to the
    workspace in global memory.  The heavy/light kernel does not take vocabularies this large; the dispatch must fall through."""
    from b200rag import _lib
    from oracle import fusion
    rng = np.random.default_rng(31)
    n_docs, b, n_max, k = 500, 3, 200, 25
    docs = [np.sort(rng.choice(vocab, size=int(rng.integers(5, 60)), replace=False)) for _ in range(n_docs)]
    common = np.sort(rng.choice(vocab, size=30, replace=False))
    docs = [np.union1d(d, common[rng.random(30) < 0.5]) for d in docs]        # shared tokens: non-trivial similarities
    dp = np.zeros(n_docs + 1, np.int64)
    dp[1:] = np.cumsum([len(d) for d in docs])
    ti = np.concatenate(docs).astype(np.int32)
    cand = np.stack([rng.choice(n_docs, size=n_max, replace=False) for _ in range(b)]).astype(np.int32)
    n = np.asarray([n_max, 150, 33], np.int32)
    rel = np.sort(rng.random((b, n_max)) * 0.016, axis=1)[:, ::-1].copy()
    lam, ks = [0.7, 0.3, 0.5], [k, k, 10]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    _lib.set_option("mmr_path", mmr_path)
    try:
        picks, pn = eng.mmr_select(t(cand), t(rel), t(n), t(dp), t(ti), vocab, t(np.asarray(lam)), t(np.asarray(ks, np.int32)), k)
    finally:
        _lib.set_option("mmr_path", -1)
    for q in range(b):
        sets = [frozenset(docs[d].tolist()) for d in cand[q, : n[q]]]
        ref = fusion.mmr_select(list(rel[q, : n[q]]), sets, ks[q], lam[q])
        assert picks[q, : int(pn[q])].cpu().tolist() == ref, q


def test_mmr_batches_beyond_one_wave_agree_with_the_bitset_kernels(eng):
    """More than 1024 queries: the heavy/light kernel is launched in waves over one workspace region; every query must
    come back exactly as the round-1 bitset kernels (themselves oracle-checked above) return it."""
    from b200rag import _lib, synth
    vocab, n_docs, b, n_max, k = 5000, 3000, 1100, 48, 12
    dp, ti, _ = synth.zipf_corpus(n_docs, vocab, 9, mean_len=25)
    rng = np.random.default_rng(5)
    cand = rng.integers(0, n_docs, size=(b, n_max)).astype(np.int32)
    n = rng.integers(1, n_max + 1, size=b).astype(np.int32)
    rel = np.sort(rng.random((b, n_max)) * 0.016, axis=1)[:, ::-1].copy()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    args = (t(cand), t(rel), t(n), t(dp), t(ti.astype(np.int32)), vocab, t(rng.choice([0.3, 0.7, 1.0], size=b)),
            t(rng.integers(0, k + 1, size=b).astype(np.int32)), k)
    out = {}
    for path in (3, 2):
        _lib.set_option("mmr_path", path)
        try:
            picks, pn = eng.mmr_select(*args)
        finally:
            _lib.set_option("mmr_path", -1)
        out[path] = (picks.cpu().numpy(), pn.cpu().numpy())
    assert (out[3][1] >= 0).all()
    assert np.array_equal(out[3][1], out[2][1]) and np.array_equal(out[3][0], out[2][0])


def test_fuse_select_matches_the_tensor_restatement(eng):
    """b200rag_fuse_select (the tail of the fusion stage as one kernel) against the chain of tensor operations it replaced
    (reference retrieval.py:322-333, 441-461, 485-491, 512-516): mixed MMR / plain queries, picks padded with -1, fewer fused
    entries than top_k, an empty query."""
    rng = np.random.default_rng(41)
    nl, b, kmax, t_max = 3, 37, 24, 15
    tot = nl * kmax
    lsc = rng.random((nl, b, kmax))
    n_f = rng.integers(0, tot + 1, size=b).astype(np.int32)
    n_f[3] = 0
    ids = np.full((b, tot), -1, np.int64)
    sc = np.zeros((b, tot))
    mask = np.zeros((b, tot), np.int32)
    first = np.full((b, tot), -1, np.int32)
    for q in range(b):
        ids[q, : n_f[q]] = rng.permutation(10_000)[: n_f[q]]
        sc[q, : n_f[q]] = np.sort(rng.random(n_f[q]))[::-1]
        mask[q, : n_f[q]] = rng.integers(1, 8, size=n_f[q])
        first[q, : n_f[q]] = rng.integers(0, tot, size=n_f[q])
    top = rng.integers(1, t_max + 1, size=b).astype(np.int32)
    use = (rng.random(b) < 0.5).astype(np.int32)
    picks = np.full((b, t_max), -1, np.int32)
    for q in range(b):
        m = min(int(n_f[q]), int(top[q]))
        if m:
            picks[q, :m] = rng.permutation(int(n_f[q]))[:m]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    fused = eng.FusedBatch(t(ids), t(sc), t(mask), t(first), t(n_f))
    rows, scores, msk, n_out, fm, orig = eng.fuse_select(fused, t(picks), t(use), t(top), t(lsc), t_max)
    # the tensor-operation chain (what B200HybridRetriever._fuse_local did before the kernel existed)
    pos = np.tile(np.arange(t_max), (b, 1))
    pos = np.where(use[:, None] != 0, np.maximum(picks, 0), pos)
    pos = np.minimum(pos, tot - 1)
    n_ref = np.minimum(n_f, top)
    valid = np.arange(t_max)[None, :] < n_ref[:, None]
    take = lambda a: np.take_along_axis(a, pos, 1)
    assert np.array_equal(n_out.cpu().numpy(), n_ref)
    assert np.array_equal(rows.cpu().numpy(), np.where(valid, take(ids), -1))
    assert np.array_equal(scores.cpu().numpy(), np.where(valid, take(sc), -np.inf))
    assert np.array_equal(msk.cpu().numpy(), np.where(valid, take(mask), 0))
    f = np.maximum(take(first), 0)
    assert np.array_equal(fm.cpu().numpy(), f // kmax)
    flat = lsc.transpose(1, 0, 2).reshape(b, -1)
    assert np.array_equal(orig.cpu().numpy(), np.take_along_axis(flat, f.astype(np.int64), 1))
    # no MMR anywhere: picks / use_mmr may be omitted
    rows2, *_ = eng.fuse_select(fused, None, None, t(top), t(lsc), t_max)
    pos0 = np.minimum(np.tile(np.arange(t_max), (b, 1)), tot - 1)
    assert np.array_equal(rows2.cpu().numpy(), np.where(valid, np.take_along_axis(ids, pos0, 1), -1))


def test_mmr_heavy_cap_and_long_documents(eng):
    """The heavy/light MMR kernel outside its comfort zone: a 3000-token vocabulary where the sample calls far more than 384
    tokens heavy (the cap moves the rest to the light lists), two documents of 1500 and 1100 tokens (several rounds of the per-pick
    walk, buckets longer than a warp; the average stays under the workspace's 256 tokens per candidate), duplicates of one document (similarity 1.0, exact score ties) and empty documents."""
    from b200rag import _lib
    from oracle import fusion
    rng = np.random.default_rng(23)
    vocab, n_docs, b, n_max, k = 3000, 400, 4, 300, 40
    docs = [np.sort(rng.choice(vocab, size=int(rng.integers(60, 300)), replace=False)) for _ in range(n_docs)]
    docs[20], docs[21] = (np.sort(rng.choice(vocab, size=m, replace=False)) for m in (1500, 1100))
    docs[7] = docs[3].copy()
    docs[11] = docs[3].copy()
    docs[5] = np.zeros(0, np.int64)
    dp = np.zeros(n_docs + 1, np.int64)
    dp[1:] = np.cumsum([len(d) for d in docs])
    ti = np.concatenate(docs).astype(np.int32)
    cand = np.stack([rng.choice(n_docs, size=n_max, replace=False) for _ in range(b)]).astype(np.int32)
    cand[0, :6] = [3, 7, 11, 5, 20, 21]
    n = np.asarray([n_max, 299, 17, 1], np.int32)
    rel = np.sort(rng.random((b, n_max)) * 0.016, axis=1)[:, ::-1].copy()
    rel[0, 1] = rel[0, 2]
    lam, ks = [0.7, 0.0, 0.5, 0.9], [k, k, 17, 5]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    _lib.set_option("mmr_path", 3)
    try:
        picks, pn = eng.mmr_select(t(cand), t(rel), t(n), t(dp), t(ti), vocab, t(np.asarray(lam)), t(np.asarray(ks, np.int32)), k)
    finally:
        _lib.set_option("mmr_path", -1)
    for q in range(b):
        sets = [frozenset(docs[d].tolist()) for d in cand[q, : n[q]]]
        ref = fusion.mmr_select(list(rel[q, : n[q]]), sets, ks[q], lam[q])
        assert int(pn[q]) == len(ref) and picks[q, : int(pn[q])].cpu().tolist() == ref, (q, int(pn[q]))     # (-2 = not served here)


def test_sparse_filtered_matches_oracle(eng, oracle_lib):
    """doc_mask applied in the candidate collection of the sparse scan: exact top-k of the allowed documents."""
    from b200rag import synth
    o = oracle_lib
    n_docs, vocab, n_q, k = 60000, 3000, 24, 30
    dp, ti, w, qp, qt, qv = _sparse_case(n_docs, vocab, n_q, seed=9)
    rng = np.random.default_rng(4)
    allowed = rng.random(n_docs) < 0.2
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
    big = n_docs
    rs, ri, rc = o.sparse_topk(tp, pd, pw, n_docs, qp, qt, qv, big)
    sidx = eng.SparseIndex(dp, ti, w, vocab, DEV)
    mask = eng.pack_row_mask(torch.from_numpy(allowed).to(DEV))
    s, i, c = sidx.search(qp, qt, qv, k, doc_mask=mask)
    s, i, c = s.cpu().numpy(), i.cpu().numpy(), c.cpu().numpy()
    for q in range(n_q):
        assert rc[q] <= big                                  # k = n_docs: the oracle list is complete
        keep = [j for j in range(rc[q]) if allowed[ri[q, j]]][:k]
        assert c[q] == len(keep)
        assert list(i[q, : c[q]]) == [int(ri[q, j]) for j in keep]
        assert list(s[q, : c[q]]) == [rs[q, j] for j in keep]


# ------------------------------------------------------------------------------------------------ rerank + diversity (8f-4)
def test_learned_rerank_kernel_matches_reference_golden_and_oracle(eng):
    """b200rag_rerank_learned == the reference's rerank with the LearnedRanker (golden from retrieval.py:518-563 +
    ranker.py:109-125) and == the oracle restatement on random batches with ties, recency and ragged lengths."""
    from oracle import fusion
    g = load_golden()
    for case in g["rerank"]:
        sc = np.asarray([float.fromhex(h) for h in case["in_scores_hex"]])
        mk = np.asarray([(1 << m) - 1 for m in case["in_n_methods"]], np.int32)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
        pos, out, cnt = eng.rerank_learned(t(sc[None, :]), t(mk[None, :]), t(np.asarray([sc.size], np.int32)), case["top_k"])
        m = int(cnt[0])
        assert [case["in_ids"][i] for i in pos[0, :m].cpu().tolist()] == case["out_ids"]
        assert [float(v).hex() for v in out[0, :m].cpu().tolist()] == case["out_scores_hex"]
    rng = np.random.default_rng(3)
    b, t_max, k_out = 40, 100, 7
    sc = np.round(rng.random((b, t_max)) * 0.02, 3)                 # rounded: plenty of exact ties
    mk = rng.integers(1, 8, (b, t_max)).astype(np.int32)
    rec = rng.random((b, t_max))
    n = rng.integers(0, t_max + 1, b).astype(np.int32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    for bw, mb, rw, use_rec in ((1.0, 0.1, 0.0, False), (2.0, 0.5, 0.25, True)):
        pos, out, cnt = eng.rerank_learned(t(sc), t(mk), t(n), k_out, bw, mb, rw, t(rec) if use_rec else None)
        for q in range(b):
            nm = [bin(int(v)).count("1") for v in mk[q, : n[q]]]
            order, rs = fusion.learned_rank(list(sc[q, : n[q]]), nm, list(rec[q, : n[q]]) if use_rec else [0.0] * int(n[q]), k_out, bw, mb, rw)
            m = int(cnt[q])
            assert m == len(order) and pos[q, :m].cpu().tolist() == order and out[q, :m].cpu().tolist() == rs, q
            assert (pos[q, m:] == -1).all()


def test_pairwise_jaccard_kernel_matches_reference_golden_and_oracle(eng):
    """b200rag_pairwise_jaccard == RAGEvaluator._calculate_pairwise_similarity (golden from evaluation.py:327-344, numpy mean
    included) and == the oracle restatement on random result lists."""
    from oracle import fusion
    g = load_golden()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)

    def run(lists):                                # lists: per query a list of contents
        vocab, ptr, ids = {}, [0], []
        docs = np.zeros((len(lists), max(1, max(len(l) for l in lists))), np.int32)
        row = 0
        for q, l in enumerate(lists):
            for j, c in enumerate(l):
                s = sorted({vocab.setdefault(tok, len(vocab)) for tok in fusion.tokens(c)})
                ids.extend(s)
                ptr.append(len(ids))
                docs[q, j] = row
                row += 1
        if row == 0:
            ptr.append(0)
        mean, pairs = eng.pairwise_jaccard(t(docs), t(np.asarray([len(l) for l in lists], np.int32)), t(np.asarray(ptr, np.int64)),
                                           t(np.asarray(ids if ids else [0], np.int32)))
        return mean.cpu().tolist(), pairs.cpu().tolist()

    lists = [c["contents"] for c in g["pairwise_similarity"]]
    mean, pairs = run(lists)
    assert [float(v).hex() for v in mean] == [c["mean_hex"] for c in g["pairwise_similarity"]]
    rng = np.random.default_rng(8)
    lists = [[" ".join(f"w{rng.integers(30)}" for _ in range(rng.integers(0, 20))) for _ in range(rng.integers(0, 40))] for _ in range(25)]
    mean, pairs = run(lists)
    for q, l in enumerate(lists):
        want, np_ = fusion.pairwise_similarity([fusion.tokens(c) for c in l])
        assert mean[q] == want and pairs[q] == np_, q
