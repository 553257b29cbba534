"""Debug helper: one dense shape through every mode, with and without the tier-0 re-scan, against the oracle."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import numpy as np, torch
from b200rag import engine as eng, _lib
from oracle import oracle as o
o.build()
for (n, d, b, k, dt) in ((40000, 200, 64, 1, "f16"), (20000, 1024, 64, 100, "bf16"), (100000, 768, 256, 100, "f16"), (60000, 384, 1024, 10, "f16")):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((b, d)).astype(np.float32)
    code = o.F16 if dt == "f16" else o.BF16
    xb, qb = o.normalize_rows(x, code), o.normalize_rows(q, code)
    ref_s, ref_i = o.dense_topk(xb, qb, k, code)
    for t0 in (1, 0):
        _lib.set_option("no_tier0", t0)
        idx = eng.DenseIndex(d, dt, "COSINE", "cuda:0")
        idx.add(torch.from_numpy(x))
        for rep in range(3):
            for mode in (eng.DENSE_TENSOR, eng.DENSE_AUTO):
                s, i, f = idx.search(torch.from_numpy(q), k, mode=mode)
                ok_i = np.array_equal(i.cpu().numpy(), ref_i)
                ok_s = np.array_equal(s.cpu().numpy(), ref_s)
                bad = int((i.cpu().numpy() != ref_i).any(1).sum())
                print(f"shape {(n, d, b, k, dt)} no_tier0={t0} rep={rep} mode={mode} ids_ok={ok_i} scores_ok={ok_s} bad_queries={bad} flags={int(f.sum())}", flush=True)
    _lib.set_option("no_tier0", -1)
