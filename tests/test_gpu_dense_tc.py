"""GPU parity tests of the tensor-core scan (tcgen05 + TMA + fused top-k filter + exact re-score) against the CPU
oracle: ids and fp64 scores must be bit exact, the completeness proof must hold, and the error bound the proof
relies on must hold with margin."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def eng():
    from b200rag import engine
    assert engine.device_info()[1] >= 10
    return engine


def _case(eng, o, n, d, b, k, dt, seed, metric="COSINE", mode=None):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((b, d)).astype(np.float32)
    oc = o.F16 if dt == "f16" else o.BF16
    if metric == "COSINE":
        xb, qb = o.normalize_rows(x, oc), o.normalize_rows(q, oc)
    else:
        xb, qb = o.round_f32(x, oc), o.round_f32(q, oc)
    ref_s, ref_i = o.dense_topk(xb, qb, k, oc, id_offset=5)
    idx = eng.DenseIndex(d, dt, metric, DEV, id_offset=5)
    idx.add(torch.from_numpy(x))
    err = torch.zeros(b, dtype=torch.float32, device=DEV)
    s, i, f = idx.search(torch.from_numpy(q), k, mode=eng.DENSE_TENSOR if mode is None else mode, out_err=err)
    return (s.cpu().numpy(), i.cpu().numpy(), f.cpu().numpy(), err.cpu().numpy(), ref_s, ref_i,
            o.bits_to_f32(qb, oc).astype(np.float64), idx.row_norm_bound)


@pytest.mark.parametrize("n,d,b,k,dt", [
    (256, 64, 128, 10, "f16"),        # one tile, one k-block
    (300, 64, 5, 10, "f16"),          # partial second tile, partial query block
    (100, 64, 3, 100, "f16"),         # fewer rows than k'
    (1000, 768, 128, 100, "f16"),
    (5000, 384, 300, 40, "f16"),      # 3 query blocks, last one partial
    (40000, 200, 64, 1, "f16"),       # dim not a multiple of 64 (TMA zero fill), k = 1
    (20000, 1024, 64, 100, "bf16"),   # config-4-like dtype/dim
    (100000, 768, 256, 100, "f16"),
    (60000, 384, 1024, 10, "f16"),    # 8 query blocks, config-5-like k
    (30000, 128, 16, 500, "f16"),     # deep candidate list (config-4-like K)
    (6000, 1536, 70, 20, "f16"),      # the reference's default semantic_dim (indexing.py:61-77)
    (3000, 4096, 130, 10, "bf16"),    # very wide vectors: 64 k-blocks, fewer staged rows in the re-score
    (2000, 8, 9, 5, "f16"),           # narrowest vectors the ABI accepts
])
def test_tensor_path_matches_oracle(eng, oracle_lib, n, d, b, k, dt):
    s, i, f, err, ref_s, ref_i, qf, rnb = _case(eng, oracle_lib, n, d, b, k, dt, seed=n + d + b)
    assert np.array_equal(i, ref_i)
    assert np.array_equal(s.view(np.uint64), ref_s.view(np.uint64))     # canonical fp64 scores, bit exact
    assert f.sum() == 0                                                  # every query proven complete
    # the proof's error margin eps = 2*dim*2^-23*|q|*row_norm_bound must dominate the observed tensor-core error
    eps = 2.0 * d * 2.0 ** -23 * np.sqrt((qf ** 2).sum(1)) * rnb
    assert np.all(err <= eps / 4), (err.max(), eps.min())


@pytest.mark.parametrize("finish_version", [1, 2, 3])
@pytest.mark.parametrize("n,d,b,k,dt", [(100, 64, 3, 100, "f16"), (20000, 1024, 64, 100, "bf16"), (30000, 128, 16, 500, "f16"),
                                         (3000, 4096, 130, 10, "bf16"), (60000, 384, 1024, 10, "f16")])
def test_every_finish_generation_matches_oracle(eng, oracle_lib, finish_version, n, d, b, k, dt):
    """The three finish-stage generations (1: block-wide streaming top-k, 2: warp-select monolithic, 3: split select /
    wide re-score / rank) are selected by k' in the product path; forced one by one they must all return the oracle's answer."""
    from b200rag import _lib
    _lib.set_option("finish_version", finish_version)
    try:
        s, i, f, err, ref_s, ref_i, qf, rnb = _case(eng, oracle_lib, n, d, b, k, dt, seed=n + d + b + 1)
    finally:
        _lib.set_option("finish_version", -1)
    assert np.array_equal(i, ref_i)
    assert np.array_equal(s.view(np.uint64), ref_s.view(np.uint64))
    assert f.sum() == 0
    eps = 2.0 * d * 2.0 ** -23 * np.sqrt((qf ** 2).sum(1)) * rnb
    assert np.all(err <= eps / 4), (err.max(), eps.min())


def test_tensor_path_ip_metric_unnormalised(eng, oracle_lib):
    s, i, f, err, ref_s, ref_i, qf, rnb = _case(eng, oracle_lib, 50000, 256, 40, 20, "f16", seed=3, metric="IP")
    assert np.array_equal(i, ref_i) and np.array_equal(s, ref_s)
    eps = 2.0 * 256 * 2.0 ** -23 * np.sqrt((qf ** 2).sum(1)) * rnb
    assert np.all(err <= eps / 4)


def test_massive_ties_are_flagged_and_auto_mode_falls_back(eng, oracle_lib):
    o = oracle_lib
    rng = np.random.default_rng(21)
    base = rng.standard_normal((9, 64)).astype(np.float32)
    x = np.concatenate([base] * 400)                           # 400 exact copies of every row
    q = base[:4]
    xb, qb = o.normalize_rows(x, o.F16), o.normalize_rows(q, o.F16)
    ref_s, ref_i = o.dense_topk(xb, qb, 100, o.F16)
    idx = eng.DenseIndex(64, "f16", "COSINE", DEV)
    idx.add(torch.from_numpy(x))
    s, i, f = idx.search(torch.from_numpy(q), 100, mode=eng.DENSE_TENSOR)
    assert f.cpu().numpy().all(), "400-way ties cannot be proven complete with k'=128 candidates"
    s, i, f = idx.search(torch.from_numpy(q), 100, mode=eng.DENSE_AUTO)
    assert np.array_equal(i.cpu().numpy(), ref_i) and np.array_equal(s.cpu().numpy(), ref_s)
    f = f.cpu().numpy()
    # AUTO: the flagged queries went through the tier-0 re-scan (bit 1), whose k' = 640 candidates hold the whole 400-way tie
    # group, so the proof succeeded there (bit 0 cleared) and the CUDA-core scan was not needed
    assert ((f & 2) == 2).all() and ((f & 1) == 0).all()


@pytest.mark.parametrize("copies,tier0", [(8, True), (64, True), (64, False), (700, True)])
def test_near_duplicate_corpus_is_exact_with_and_without_tier0(eng, oracle_lib, copies, tier0):
    """corpora with duplicated chunks.  Every row occurs `copies` times; tie groups straddle rank k.  The
    result must equal the oracle whichever path resolves a query: first pass (8 copies fit k' = k + 28), tier-0 re-scan
    (64 copies), or the exact CUDA-core scan (tier 0 switched off, or a 700-way tie that exceeds even k' = 640)."""
    from b200rag import _lib
    o = oracle_lib
    rng = np.random.default_rng(copies)
    n_base, d, b, k = 40_000 // copies, 128, 160, 100
    base = rng.standard_normal((n_base, d)).astype(np.float32)
    x = np.tile(base, (copies, 1))                              # copy c of row r has id c * n_base + r
    q = rng.standard_normal((b, d)).astype(np.float32)
    xb, qb = o.normalize_rows(x, o.F16), o.normalize_rows(q, o.F16)
    ref_s, ref_i = o.dense_topk(xb, qb, k, o.F16)
    idx = eng.DenseIndex(d, "f16", "COSINE", DEV)
    idx.add(torch.from_numpy(x))
    _lib.set_option("no_tier0", 0 if tier0 else 1)
    try:
        s, i, f = idx.search(torch.from_numpy(q), k, mode=eng.DENSE_AUTO)
    finally:
        _lib.set_option("no_tier0", -1)
    assert np.array_equal(i.cpu().numpy(), ref_i) and np.array_equal(s.cpu().numpy(), ref_s)
    f = f.cpu().numpy()
    if copies == 8:
        assert (f == 0).all()
    elif copies == 64 and tier0:
        assert ((f & 2) == 2).any() and ((f & 1) == 0).all()     # resolved by the re-scan
    elif copies == 64:
        assert ((f & 1) == 1).any() and ((f & 2) == 0).all()     # straight to the exact scan
    else:
        assert ((f & 3) == 3).any()                              # re-scanned AND still unproven: exact scan


def test_auto_equals_exact_mode_at_one_million_rows(eng):
    """Size-independent property at BASELINE config-2 scale (1M x 768, top-100): the tensor path and the CUDA-core
    exact path (itself oracle-checked at small sizes) return identical ids and scores; lists are sorted."""
    g = torch.Generator(device=DEV).manual_seed(0)
    idx = eng.DenseIndex(768, "f16", "COSINE", DEV, capacity=1_000_000)
    for _ in range(4):
        idx.add(torch.randn(250_000, 768, generator=g, device=DEV))
    q = torch.randn(16, 768, generator=g, device=DEV)
    s1, i1, f1 = idx.search(q, 100, mode=eng.DENSE_AUTO)
    s2, i2, _ = idx.search(q, 100, mode=eng.DENSE_EXACT)
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    assert int(f1.sum()) == 0
    assert bool((s1[:, 1:] <= s1[:, :-1]).all())
    assert bool(((s1[:, 1:] < s1[:, :-1]) | (i1[:, 1:] > i1[:, :-1])).all())


@pytest.mark.parametrize("b", [129, 257, 520])
def test_ragged_batches_through_the_pair_kernel(eng, oracle_lib, b):
    """Batches that do not fill whole 128-query blocks / block pairs (zero-padded query blocks, padded pair)."""
    s, i, f, err, ref_s, ref_i, qf, rnb = _case(eng, oracle_lib, 70000, 256, b, 20, "f16", seed=b, mode=eng.DENSE_AUTO)
    assert np.array_equal(i, ref_i) and np.array_equal(s, ref_s)
    assert f.sum() == 0


def test_more_flagged_queries_than_the_first_fallback_tier(eng, oracle_lib):
    """48 queries that all hit 300-way ties: the device-gated exact fallback must cover tier A (32 slots) and tier B."""
    o = oracle_lib
    rng = np.random.default_rng(5)
    base = rng.standard_normal((60, 64)).astype(np.float32)
    x = np.concatenate([base] * 300)
    q = base[:48]
    xb, qb = o.normalize_rows(x, o.F16), o.normalize_rows(q, o.F16)
    ref_s, ref_i = o.dense_topk(xb, qb, 100, o.F16)
    idx = eng.DenseIndex(64, "f16", "COSINE", DEV)
    idx.add(torch.from_numpy(x))
    from b200rag import _lib
    _lib.set_option("no_tier0", 1)                               # straight to the exact scan: this test is about ITS two tiers
    try:
        s, i, f = idx.search(torch.from_numpy(q), 100, mode=eng.DENSE_AUTO)
    finally:
        _lib.set_option("no_tier0", -1)
    assert int(f.sum()) == 48
    assert np.array_equal(i.cpu().numpy(), ref_i) and np.array_equal(s.cpu().numpy(), ref_s)
    # with the tier-0 re-scan in place the 300-way tie groups fit its k' = 640 candidates: no query reaches the exact scan
    s, i, f = idx.search(torch.from_numpy(q), 100, mode=eng.DENSE_AUTO)
    assert (f.cpu().numpy() == 2).all()
    assert np.array_equal(i.cpu().numpy(), ref_i) and np.array_equal(s.cpu().numpy(), ref_s)


def test_full_size_properties_10m_rows(eng):
    """BASELINE config 3 at full size (10M x 768 fp16 cosine, top-100, batch 1024) through size-independent properties:
    the tensor path equals the CUDA-core exact path on a query subset, every list is strictly ordered by (score desc,
    id asc), no query is flagged, and searching two half shards + the merge kernel reproduces the whole-index result."""
    n, d, b, k = 10_000_000, 768, 1024, 100
    g = torch.Generator(device=DEV).manual_seed(7)
    idx = eng.DenseIndex(d, "f16", "COSINE", DEV, capacity=n)
    for _ in range(n // 250_000):
        idx.add(torch.randn(250_000, d, generator=g, device=DEV))
    q = torch.randn(b, d, generator=g, device=DEV)
    s, i, f = idx.search(q, k, mode=eng.DENSE_AUTO)
    assert int(f.sum()) == 0
    assert bool(((s[:, 1:] < s[:, :-1]) | ((s[:, 1:] == s[:, :-1]) & (i[:, 1:] > i[:, :-1]))).all())
    assert int(i.min()) >= 0 and int(i.max()) < n
    sub = torch.tensor([0, 1, 127, 128, 511, 640, 1000, 1023], device=DEV)
    s2, i2, _ = idx.search(q[sub], k, mode=eng.DENSE_EXACT)
    assert torch.equal(i[sub], i2) and torch.equal(s[sub], s2)
    # two half shards (views of the same rows) + merge == whole
    halves = []
    q16 = idx.prepare_queries(q)
    for lo, hi in ((0, n // 2), (n // 2, n)):
        hs, hi_ids, _ = eng.dense_topk(idx.rows[lo:hi], q16, k, id_offset=lo)
        halves.append((hs, hi_ids))
    ms, mi = eng.merge_topk(torch.cat([h[0] for h in halves], 1).contiguous(), torch.cat([h[1] for h in halves], 1).contiguous(), k)
    assert torch.equal(mi, i) and torch.equal(ms, s)
    # the same through the all-gather layout the multi-GPU path uses
    from b200rag import distributed as bdist
    gathered = torch.stack([bdist.pack_candidates(h[0], h[1]) for h in halves]).contiguous()
    gs, gi = eng.merge_gathered(gathered, k)
    assert torch.equal(gi, i) and torch.equal(gs, s)


def test_random_shapes_against_oracle(eng, oracle_lib):
    """Seeded fuzz over shapes, dtypes, metrics and modes (AUTO path): ids and fp64 scores bit-exact every time."""
    rng = np.random.default_rng(2024)
    for case in range(14):
        d = int(rng.integers(1, 66)) * 8
        n = int(rng.choice([1, 7, 255, 256, 257, int(rng.integers(300, 30000))]))
        b = int(rng.choice([1, 2, 127, 128, 129, int(rng.integers(1, 400))]))
        k = int(rng.integers(1, 150))
        dt = str(rng.choice(["f16", "bf16"]))
        metric = str(rng.choice(["COSINE", "IP"]))
        s, i, f, err, ref_s, ref_i, qf, rnb = _case(eng, oracle_lib, n, d, b, k, dt, seed=1000 + case, metric=metric,
                                                    mode=eng.DENSE_AUTO)
        assert np.array_equal(i, ref_i), (case, n, d, b, k, dt, metric)
        assert np.array_equal(s, ref_s), (case, n, d, b, k, dt, metric)


def test_auto_mode_serves_k_beyond_the_tensor_path_limit(eng, oracle_lib):
    """k = 1500 exceeds the candidate-buffer limit of the tensor-core scan; AUTO must route to the exact scan by itself."""
    s, i, f, err, ref_s, ref_i, qf, rnb = _case(eng, oracle_lib, 9000, 64, 5, 1500, "f16", seed=77, mode=eng.DENSE_AUTO)
    assert np.array_equal(i, ref_i) and np.array_equal(s, ref_s) and f.sum() == 0


@pytest.mark.parametrize("keep_frac", [0.5, 0.05, 0.0004])
def test_filtered_scan_equals_oracle_on_the_allowed_rows(eng, oracle_lib, keep_frac):
    """Metadata filter applied inside the kernels (b200rag_dense_topk_masked): the result is the exact top-k of the allowed
    rows, for mild, selective and nearly-empty masks (the last one leaves fewer allowed rows than k), in AUTO and EXACT mode."""
    o = oracle_lib
    n, d, b, k = 300_000, 128, 130, 50
    rng = np.random.default_rng(31)
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((b, d)).astype(np.float32)
    allowed = rng.random(n) < keep_frac
    rows = np.flatnonzero(allowed)
    xb, qb = o.normalize_rows(x, o.F16), o.normalize_rows(q, o.F16)
    kk = min(k, rows.size)
    ref_s, ref_local = o.dense_topk(xb[rows], qb, kk, o.F16)
    ref_i = rows[ref_local]
    idx = eng.DenseIndex(d, "f16", "COSINE", DEV)
    idx.add(torch.from_numpy(x))
    mask = eng.pack_row_mask(torch.from_numpy(allowed).to(DEV))
    for mode in (eng.DENSE_AUTO, eng.DENSE_EXACT):
        s, i, f = idx.search(torch.from_numpy(q), k, mode=mode, row_mask=mask)
        s, i = s.cpu().numpy(), i.cpu().numpy()
        assert np.array_equal(i[:, :kk], ref_i), (keep_frac, mode)
        assert np.array_equal(s[:, :kk], ref_s), (keep_frac, mode)
        assert (i[:, kk:] == -1).all() and np.isneginf(s[:, kk:]).all()


def test_approximate_mode_recall_and_tolerance(eng, oracle_lib):
    """B200RAG_DENSE_APPROX ranks by the tensor-core fp32 scores (no fp64 re-score, no proof): north star -- fused scores
    within 1e-3 relative in fp16/bf16, recall@k reported against the exact scan.  On random data it loses at most near-ties."""
    o = oracle_lib
    for (n, d, b, k, dt) in ((50_000, 384, 200, 20, "f16"), (30_000, 1024, 130, 100, "bf16")):
        rng = np.random.default_rng(n)
        x = rng.standard_normal((n, d)).astype(np.float32)
        q = rng.standard_normal((b, d)).astype(np.float32)
        code = o.F16 if dt == "f16" else o.BF16
        ref_s, ref_i = o.dense_topk(o.normalize_rows(x, code), o.normalize_rows(q, code), k, code)
        idx = eng.DenseIndex(d, dt, "COSINE", DEV)
        idx.add(torch.from_numpy(x))
        s, i, f = idx.search(torch.from_numpy(q), k, mode=eng.DENSE_APPROX)
        s, i = s.cpu().numpy(), i.cpu().numpy()
        recall = np.mean([len(set(i[r]) & set(ref_i[r])) / k for r in range(b)])
        assert recall >= 0.995, recall
        assert int(f.sum()) == 0 and (np.diff(s, axis=1) <= 0).all()
        same = i == ref_i
        assert np.abs(s[same] - ref_s[same]).max() <= 1e-3 * np.abs(ref_s[same]).max()      # tolerance stated by the north star
