"""MMR kernel phase split: time at k = 1 / 10 / 100 picks (C4 shape) for the inverted-list and bitset kernels."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import numpy as np, torch
from b200rag import engine as eng, _lib, synth
DEV = "cuda:0"
vocab, n_docs, b, n_max = 100_000, 200_000, 1024, 1000
dp, ti, _ = synth.zipf_corpus(n_docs, vocab, 3)
rng = np.random.default_rng(17)
cand = rng.integers(0, n_docs, size=(b, n_max)).astype(np.int32)
cand = np.sort(cand, axis=1); cand += np.arange(n_max, dtype=np.int32)[None, :] * 0; 
for r in range(b):
    cand[r] = rng.choice(n_docs, size=n_max, replace=False)
n = np.full(b, n_max, np.int32)
rel = np.sort(rng.random((b, n_max)) * 0.016, axis=1)[:, ::-1].copy()
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
args = (t(cand), t(rel), t(n), t(dp), t(ti.astype(np.int32)), vocab, t(np.full(b, 0.7)))
for path in (3, 2):
    _lib.set_option("mmr_path", path)
    for k in (1, 2, 10, 100):
        ks = t(np.full(b, k, np.int32))
        for _ in range(3): eng.mmr_select(*args, ks, 100)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): eng.mmr_select(*args, ks, 100)
        e1.record(); torch.cuda.synchronize()
        print(f"path {path} k {k:3d}: {e0.elapsed_time(e1)/10:.3f} ms", flush=True)
