"""Same-box A/B of the sparse scan (BASELINE config 4's BM25 side: 1M documents, 100K-term Zipf vocabulary, 256 queries of
8 terms, top-500) under the kernel's switches:

    option "sparse_flags"   bit 0 = dense collect, bit 1 = bulk append (default 3); + 8 = the experimental term-mask kernel
    option "sparse_slices"  slices per query (0 = the library's choice)
    --block-docs a,b,...    documents per postings block
    --variants flags:slices,...

Every variant must return the same ids / scores as the first one (bit exact); times are CUDA-event medians.
No oracle use: this is a profiling helper, parity lives in tests/.

    python tools/sparse_ab.py [--docs 1000000] [--variants dense:slices,...]
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]
from b200rag import _lib, bm25, engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=1_000_000)
ap.add_argument("--vocab", type=int, default=100_000)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--k", type=int, default=500)
ap.add_argument("--reps", type=int, default=15)
ap.add_argument("--block-docs", default="16384")
ap.add_argument("--variants", default="3:0,3:1,3:2,0:0,11:0,11:4")
ap.add_argument("--stats", action="store_true", help="per-phase cycle counters of one launch (b200rag_debug_set_stats_buffer)")
args = ap.parse_args()
dev = torch.device("cuda:0")


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
    return statistics.median(ts), min(ts), r


doc_ptr, term_ids, tf = synth.zipf_corpus_device(args.docs, args.vocab, 0, dev)
w = bm25.bm25_weights_device(doc_ptr, term_ids, tf, args.vocab)
qs = [synth.zipf_queries(args.batch, args.vocab, 100 + i, n_terms=8, skip_top=100) for i in range(4)]
ref = None
for bd in [int(x) for x in args.block_docs.split(",")]:
    idx = engine.SparseIndex(doc_ptr, term_ids, w, args.vocab, dev, block_docs=bd)
    nbytes = statistics.mean(idx.query_bytes(q[0], q[1]) for q in qs)
    for var in args.variants.split(","):
        dense, slices = var.split(":")
        _lib.set_option("sparse_flags", int(dense))
        _lib.set_option("sparse_slices", int(slices) if int(slices) > 0 else -1)
        it = [0]

        def fn():
            it[0] += 1
            qp, qt, qv = qs[it[0] % 4]
            return idx.search(qp, qt, qv, args.k)

        med, best, _ = timed(fn, args.reps)
        s, i, c = idx.search(*qs[0], args.k)
        got = (s.cpu(), i.cpu(), c.cpu())
        same = True
        if ref is None:
            ref = got
        else:
            same = all(torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a, b.view(torch.int32) if b.dtype == torch.float32 else b)
                       for a, b in zip(got, ref))
        print(json.dumps({"block_docs": bd, "flags": int(dense), "slices": int(slices), "ms": round(med, 4), "best_ms": round(best, 4),
                          "gbs": round(nbytes / med / 1e6, 1), "same_as_first": bool(same)}), flush=True)
        if args.stats:
            lib = _lib.load()
            buf = torch.zeros((1024, 12), dtype=torch.int64, device=dev)
            lib.b200rag_debug_set_stats_buffer(1, buf.data_ptr(), buf.numel())
            idx.search(*qs[0], args.k)
            torch.cuda.synchronize()
            lib.b200rag_debug_set_stats_buffer(1, None, 0)
            used = buf.cpu().numpy().astype(np.float64)
            used = used[used.sum(1) > 0]
            if int(dense) & 8:      # sparse_mask.cu: SMS_* slots
                names = ((0, "init + finalize"), (9, "block top (wait)"), (1, "mark"), (2, "score"), (3, "drain to top-k"))
            else:                   # sparse_bm25.cu: phase slots of the product kernel
                names = ((0, "init"), (2, "fetch + term 1"), (3, "terms 2.."), (4, "scan accumulators"), (5, "bulk append"),
                         (1, "compaction"), (8, "candidate rounds"), (6, "finalize"))
            tot = used[:, [i for i, _ in names]].sum(1)
            print(f"    {len(used)} CTAs; cycles per CTA: mean {tot.mean():.0f}  min {tot.min():.0f}  max {tot.max():.0f}")
            for i, n_ in names:
                v = used[:, i].mean()
                print(f"    {n_:18s} {v:10.0f} cycles  {100 * v / tot.mean():5.1f}%")
    del idx
    torch.cuda.empty_cache()
