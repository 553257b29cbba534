"""The multi-GPU exchange (b200rag/distributed.py) on CPU: two gloo ranks, each holding a row shard, all-gather of the
per-rank top-k and merge.  The local search and the merge are stood in for by the CPU oracle here (the CUDA kernels are
covered by the -m gpu tests); what is under test is the sharding arithmetic, the message packing and the collective:
the merged result must be bit-identical to a single-shard search of the whole corpus."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, d: int, b: int, k: int, out_dir: str):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "advanced-rag-milvus_b200")]
    from b200rag import distributed as bdist
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        x = rng.standard_normal((n, d)).astype(np.float32)
        x[5] = x[n - 3]                                      # a cross-shard exact tie: must resolve to the lower id
        q = rng.standard_normal((b, d)).astype(np.float32)
        xb, qb = oracle.normalize_rows(x, oracle.F16), oracle.normalize_rows(q, oracle.F16)
        start, end = bdist.shard_range(n, rank, world)
        s, i = oracle.dense_topk(xb[start:end], qb, k, oracle.F16, id_offset=start)

        def merge(cs, ci, kk):
            ms, mi = oracle.merge_topk(cs.numpy(), ci.numpy(), kk)
            return torch.from_numpy(ms), torch.from_numpy(mi)

        ms, mi = bdist.gather_and_merge(torch.from_numpy(s), torch.from_numpy(i), k, merge)
        np.save(os.path.join(out_dir, f"s{rank}.npy"), ms.numpy())
        np.save(os.path.join(out_dir, f"i{rank}.npy"), mi.numpy())
        # ---- sparse: document-range shards (cut at postings-block boundaries) with GLOBAL idf / avgdl, same exchange
        from b200rag import bm25, synth
        vocab, nq, ks, align = 300, 7, 15, 64
        dp, ti, tf = synth.zipf_corpus(n, vocab, 3, mean_len=12)
        qp, qt, qv = synth.zipf_queries(nq, vocab, 4, n_terms=5, skip_top=5)
        a, e = bdist.shard_range(n, rank, world, align=align)
        assert a % align == 0 and (e % align == 0 or e == n)
        lp, lt, lf = dp[a: e + 1] - dp[a], ti[dp[a]: dp[e]], tf[dp[a]: dp[e]]
        stats = bdist.bm25_global_stats(lp, lt, lf, vocab)
        assert stats[0] == n and int(stats[2]) == int(tf.sum())
        w = bm25.bm25_weights(lp, lt, lf, vocab, stats=stats)
        tp, pd, pw = synth.doc_major_to_term_major(lp, lt, w, vocab)
        ss, si, sc = oracle.sparse_topk(tp, pd, pw, e - a, qp, qt, qv, ks, id_offset=a)
        si = np.where(np.arange(ks)[None, :] < sc[:, None], si, -1)
        ms, mi = bdist.gather_and_merge(torch.from_numpy(ss.astype(np.float64)), torch.from_numpy(si), ks, merge)
        np.save(os.path.join(out_dir, f"ss{rank}.npy"), ms.numpy())
        np.save(os.path.join(out_dir, f"si{rank}.npy"), mi.numpy())
        np.save(os.path.join(out_dir, f"w{rank}.npy"), w)
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_rows_exactly_once():
    from b200rag.distributed import shard_range
    for n in (0, 1, 7, 8, 1000, 10_000_000):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(w - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_aligned_shard_ranges_cut_at_block_boundaries():
    from b200rag.distributed import shard_range
    for n in (1, 16383, 16384, 100_000, 1_000_000, 100_000_000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w, align=16384) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(w - 1))
            assert all(s % 16384 == 0 or s == n for s, _ in spans)
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 2 * 16384


def test_pack_unpack_round_trip():
    from b200rag.distributed import pack_candidates, unpack_gathered
    rng = np.random.default_rng(1)
    g, b, k = 3, 4, 5
    sc = torch.from_numpy(rng.standard_normal((g, b, k)))
    ids = torch.from_numpy(rng.integers(0, 1000, (g, b, k)))
    msg = torch.stack([pack_candidates(sc[r], ids[r]) for r in range(g)])
    assert msg.shape == (g, 2, b, k)                          # per rank: a plane of score bits, a plane of ids
    s2, i2 = unpack_gathered(msg, k)
    assert s2.shape == (b, g * k)
    for r in range(g):
        assert torch.equal(s2[:, r * k:(r + 1) * k], sc[r])
        assert torch.equal(i2[:, r * k:(r + 1) * k], ids[r])


@pytest.mark.parametrize("world,n", [(2, 4001), (3, 1000)])
def test_two_rank_gather_and_merge_equals_single_shard(tmp_path, oracle_lib, world, n):
    d, b, k = 64, 9, 20
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, d, b, k, str(tmp_path)), nprocs=world, join=True)
    o = oracle_lib
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[5] = x[n - 3]
    q = rng.standard_normal((b, d)).astype(np.float32)
    ref_s, ref_i = o.dense_topk(o.normalize_rows(x, o.F16), o.normalize_rows(q, o.F16), k, o.F16)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"i{r}.npy"), ref_i), f"rank {r} ids differ from the single-shard search"
        assert np.array_equal(np.load(tmp_path / f"s{r}.npy"), ref_s), f"rank {r} scores differ"
    # sparse: sharded postings with all-reduced df / avgdl == the single-shard index, weights and results bit for bit
    from b200rag import bm25, synth
    vocab, nq, ks = 300, 7, 15
    dp, ti, tf = synth.zipf_corpus(n, vocab, 3, mean_len=12)
    qp, qt, qv = synth.zipf_queries(nq, vocab, 4, n_terms=5, skip_top=5)
    w = bm25.bm25_weights(dp, ti, tf, vocab)
    assert np.array_equal(np.concatenate([np.load(tmp_path / f"w{r}.npy") for r in range(world)]).view(np.uint32), w.view(np.uint32))
    tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
    rs, ri, rc = o.sparse_topk(tp, pd, pw, n, qp, qt, qv, ks)
    ri = np.where(np.arange(ks)[None, :] < rc[:, None], ri, -1)
    for r in range(world):
        got_i, got_s = np.load(tmp_path / f"si{r}.npy"), np.load(tmp_path / f"ss{r}.npy")
        assert np.array_equal(got_i, ri), f"rank {r} sparse ids differ from the single-shard search"
        assert np.array_equal(got_s[ri >= 0].astype(np.float32).view(np.uint32), rs[ri >= 0].view(np.uint32))
