"""Multi-GPU parity check (run under torchrun on a GPU box): the row-sharded searches + NCCL all-gather + merge kernel must
return exactly what the CPU oracle returns for the whole corpus, on every rank -- for the dense index, for the row-sharded
manager's dense AND sparse collections (document-range postings shards, global idf / avgdl), with a metadata filter and
deleted rows, and for the whole hybrid chain (dense + sparse -> RRF -> MMR, fusion split by query across the ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/scripts/dist_check.py
"""
import asyncio
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]
from b200rag import bm25, distributed as bdist, synth  # noqa: E402
from b200rag.config import RetrievalConfig  # noqa: E402
from b200rag.retriever import B200HybridRetriever  # noqa: E402
from oracle import oracle, pipeline as opipe  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True


def report(name, good):
    global ok
    print(f"rank {rank}/{world}: {name}: {'bit-exact vs oracle' if good else 'MISMATCH'}", flush=True)
    ok &= bool(good)


for (n, d, b, k, dt) in ((200_003, 128, 64, 100, "f16"), (50_000, 256, 200, 10, "bf16")):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[7] = x[n - 2]                                          # cross-shard exact tie
    q = rng.standard_normal((b, d)).astype(np.float32)
    idx = bdist.ShardedDenseIndex(d, n, dt, "COSINE", dev)
    idx.local.add(torch.from_numpy(x[idx.start:idx.end]))
    s, i = idx.search(torch.from_numpy(q), k)
    code = oracle.F16 if dt == "f16" else oracle.BF16
    rs, ri = oracle.dense_topk(oracle.normalize_rows(x, code), oracle.normalize_rows(q, code), k, code)
    report(f"dense n={n} d={d} b={b} k={k} {dt}", np.array_equal(i.cpu().numpy(), ri) and np.array_equal(s.cpu().numpy(), rs))

# ---- the sharded manager: every rank passes the same global batch, keeps its own rows' vectors and postings
n, d, vocab, bd = 40_000, 64, 2000, 1024
x = synth.dense_rows(n, d, 5)
dp, ti, tf = synth.zipf_corpus(n, vocab, 5, mean_len=30)
w = bm25.bm25_weights(dp, ti, tf, vocab)
contents = [" ".join(f"w{t}" for t in ti[dp[r]: dp[r + 1]]) for r in range(n)]
meta = [{"doc_id": f"d{r // 4}", "chunk_index": r % 4, "entropy": float(r % 10) / 10.0} for r in range(n)]
ids = [f"c{r:06d}" for r in range(n)]
mgr = bdist.ShardedIndexManager(n, semantic_dim=d, sparse_dim=vocab, domain_dim=32, device=dev, sparse_block_docs=bd)
for a, e in ((0, 15_000), (15_000, n)):
    mgr.add(ids[a:e], contents[a:e], x[a:e], (dp[a: e + 1] - dp[a], ti[dp[a]: dp[e]], w[dp[a]: dp[e]]), None, meta[a:e])
assert mgr._sem.n == mgr.end - mgr.start and mgr._sparse.n_docs == mgr.end - mgr.start and mgr.num_rows == n
q = synth.dense_rows(48, d, 77)
qp, qt, qv = synth.zipf_queries(48, vocab, 6, n_terms=8, skip_top=20)
xb, qb = oracle.normalize_rows(x, oracle.F16), oracle.normalize_rows(q, oracle.F16)
tp, pd, pw = synth.doc_major_to_term_major(dp, ti, w, vocab)
arr = mgr.search_batch_arrays(q, "semantic_index", 40)
rs, ri = oracle.dense_topk(xb, qb, 40, oracle.F16)
report("manager dense", np.array_equal(arr.rows, ri) and np.array_equal(arr.scores, rs) and (arr.counts == 40).all())
sa = mgr.search_batch_arrays((qp, qt, qv), "sparse_index", 40)
ss, si, sc = oracle.sparse_topk(tp, pd, pw, n, qp, qt, qv, 40)
si = np.where(np.arange(40)[None, :] < sc[:, None], si, -1)
report("manager sparse", np.array_equal(sa.rows, si) and np.array_equal(sa.counts, sc) and
       np.array_equal(sa.scores[si >= 0].astype(np.float32).view(np.uint32), ss[si >= 0].view(np.uint32)))
# filter + deletes: replicated columns -> global mask -> every rank applies its own words
asyncio.new_event_loop().run_until_complete(mgr.delete_by_filter("semantic_index", "chunk_index == 3"))
keep = np.asarray([r % 4 != 3 and (r % 10) / 10.0 >= 0.5 for r in range(n)])
rows = np.flatnonzero(keep)
fa = mgr.search_batch_arrays(q, "semantic_index", 25, filters="entropy >= 0.5")
rs, ri = oracle.dense_topk(xb[rows], qb, 25, oracle.F16)
report("manager dense, filter + deleted rows", np.array_equal(fa.rows, rows[ri]) and np.array_equal(fa.scores, rs))
# the SPMD front door: every rank brings ITS OWN queries (48 -> world slices of equal size, the rest padded with query 0)
per = -(-48 // world)
own = np.stack([q[j] if j < 48 else q[0] for j in range(rank * per, (rank + 1) * per)])
oa = mgr.search_own_queries_arrays(own, "semantic_index", 25, filters="entropy >= 0.5")
want = [j if j < 48 else 0 for j in range(rank * per, (rank + 1) * per)]
report("manager dense, own queries (all-gather queries, all-to-all candidates)",
       np.array_equal(oa.rows, rows[ri[want]]) and np.array_equal(oa.scores, rs[want]) and (oa.counts == 25).all())
hits = mgr.search_batch(q[:2], "semantic_index", 3, filters="entropy >= 0.5")
report("manager payload lookup", [h["id"] for h in hits[1]] == [ids[r] for r in rows[ri[1, :3]]] and
       hits[1][0]["metadata"]["entropy"] == meta[rows[ri[1, 0]]]["entropy"])

# ---- hybrid chain on a fresh sharded manager (no deletes): dense + sparse -> RRF -> MMR, fusion split by query
mgr2 = bdist.ShardedIndexManager(n, semantic_dim=d, sparse_dim=vocab, domain_dim=32, device=dev, sparse_block_docs=bd)
mgr2.add(ids, contents, x, (dp, ti, w), None, meta)
retr = B200HybridRetriever(mgr2, RetrievalConfig(top_k=20))
cfgs = [RetrievalConfig(top_k=20, enable_mmr=(j % 3 == 0), mmr_lambda=0.7) for j in range(48)]
sparse_q = [{"indices": qt[qp[j]: qp[j + 1]].tolist(), "values": qv[qp[j]: qp[j + 1]].tolist()} for j in range(48)]
res = retr.retrieve_batch_embedded(q, sparse_q, cfgs)
corpus = opipe.ArrayCorpus(x, None, dp, ti, w, vocab, contents)
good = True
for j in range(0, 48, 5):
    want_ids, want_sc, _ = opipe.retrieve(corpus, q[j], sparse_q[j], None, 20, enable_mmr=cfgs[j].enable_mmr, mmr_lambda=0.7)
    m = int(res.n[j])
    good &= res.rows[j, :m].cpu().tolist() == want_ids and res.scores[j, :m].cpu().tolist() == want_sc
report("hybrid dense+sparse -> RRF -> MMR (query-split fusion)", good)

t = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(t)
dist.destroy_process_group()
sys.exit(int(t.item() != 0))
