"""Blind-debug aid for the tensor-core scan: tiny shapes first, printing what differs from the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]
from b200rag import engine  # noqa: E402
from oracle import oracle  # noqa: E402


def case(n, d, b, k, dt="f16", seed=0, mode=engine.DENSE_TENSOR):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((b, d)).astype(np.float32)
    oc = oracle.F16 if dt == "f16" else oracle.BF16
    xb, qb = oracle.normalize_rows(x, oc), oracle.normalize_rows(q, oc)
    ref_s, ref_i = oracle.dense_topk(xb, qb, k, oc)
    idx = engine.DenseIndex(d, dt, "COSINE", "cuda:0")
    idx.add(torch.from_numpy(x))
    err = torch.zeros(b, dtype=torch.float32, device="cuda:0")
    s, i, f = idx.search(torch.from_numpy(q), k, mode=mode, out_err=err)
    torch.cuda.synchronize()
    i, s = i.cpu().numpy(), s.cpu().numpy()
    ok_i = np.array_equal(i, ref_i)
    ok_s = np.array_equal(s, ref_s)
    print(f"n={n} d={d} b={b} k={k} {dt}: ids {'OK' if ok_i else 'DIFF'} scores {'OK' if ok_s else 'DIFF'} "
          f"flags={int(f.sum())} max_err={float(err.max()):.3e}", flush=True)
    if not ok_i:
        bad = np.flatnonzero((i != ref_i).any(1))
        r = bad[0]
        print("  first bad query", r, "of", bad.size, "\n  got ", i[r, :12], "\n  ref ", ref_i[r, :12])
        print("  got s", s[r, :6], "\n  ref s", ref_s[r, :6])
    return ok_i and ok_s


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), engine.device_info(), "SCAN_VERSION", os.environ.get("B200RAG_SCAN_VERSION"),
          "CLUSTER", os.environ.get("B200RAG_CLUSTER"), flush=True)
    ok = True
    for args in [(256, 64, 128, 10), (256, 128, 128, 10), (300, 64, 5, 10), (1000, 768, 128, 100), (5000, 384, 300, 40),
                 (100000, 768, 256, 100), (60000, 384, 1024, 10), (50000, 768, 512, 100)]:
        ok &= case(*args)
    ok &= case(20000, 1024, 64, 100, dt="bf16")
    ok &= case(30000, 512, 256, 40, dt="bf16")
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)
