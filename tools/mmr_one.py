"""One MMR launch at the C4 shape (256 queries x 1000 candidates, k = 100) for ncu."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import numpy as np, torch
from b200rag import engine as eng, _lib, synth
DEV = "cuda:0"
vocab, n_docs, b, n_max = 100_000, 200_000, int(sys.argv[2]) if len(sys.argv) > 2 else 256, 1000
dp, ti, _ = synth.zipf_corpus(n_docs, vocab, 3)
rng = np.random.default_rng(17)
cand = np.stack([rng.choice(n_docs, size=n_max, replace=False) for _ in range(b)]).astype(np.int32)
n = np.full(b, n_max, np.int32)
rel = np.sort(rng.random((b, n_max)) * 0.016, axis=1)[:, ::-1].copy()
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
args = (t(cand), t(rel), t(n), t(dp), t(ti.astype(np.int32)), vocab, t(np.full(b, 0.7)))
_lib.set_option("mmr_path", int(sys.argv[1]) if len(sys.argv) > 1 else 3)
ks = t(np.full(b, 100, np.int32))
for _ in range(3): eng.mmr_select(*args, ks, 100)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): eng.mmr_select(*args, ks, 100)
e1.record(); torch.cuda.synchronize()
print(f"b {b}: {e0.elapsed_time(e1)/5:.3f} ms")
