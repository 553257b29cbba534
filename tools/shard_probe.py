"""Dense top-k at small shard sizes / deep k: time, flags (C4 sharded shape: 1024-dim bf16, batch 256, k 500)."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import torch
from b200rag import engine as eng, _lib
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(1)
for rows in (62_500, 125_000, 187_500, 250_000, 1_000_000):
    for k in (100, 500):
        x = torch.randn(rows, 1024, generator=g, device=dev)
        idx = eng.DenseIndex(1024, "bf16", "COSINE", dev)
        idx.add(x)
        q = torch.randn(256, 1024, generator=g, device=dev)
        for _ in range(3): idx.search(q, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = idx.search(q, k)
        e1.record(); torch.cuda.synchronize()
        fl = idx.last_flags
        print(f"rows {rows} k {k}: {e0.elapsed_time(e1)/5:.3f} ms flags any {int((fl != 0).sum())} bit0 {int((fl & 1).ne(0).sum())} bit1 {int((fl & 2).ne(0).sum())}", flush=True)
        del idx, x
