"""One dense search at the config-4 shape (1M x 1024 bf16, batch 256, k 500) for an ncu launch list."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import torch
from b200rag import engine as eng
dev = "cuda:0"
rows, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 500
g = torch.Generator(device=dev); g.manual_seed(1)
idx = eng.DenseIndex(1024, "bf16", "COSINE", dev)
for s in range(0, rows, 250_000):
    idx.add(torch.randn(min(250_000, rows - s), 1024, generator=g, device=dev))
q = torch.randn(256, 1024, generator=g, device=dev)
from b200rag import _lib
for fv in (0, 2, 1):
    _lib.set_option("finish_version", fv)
    for kk in (k, 200, 100):
        for _ in range(4): idx.search(q, kk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): idx.search(q, kk)
        e1.record(); torch.cuda.synchronize()
        print(f"finish_version {fv} rows {rows} k {kk}: {e0.elapsed_time(e1)/10:.3f} ms")
