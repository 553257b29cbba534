"""One search over a near-duplicate corpus (64 copies of every row) for an ncu launch list: what the tier-0 re-scan costs."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import torch
from b200rag import engine
dev = "cuda:0"
n, d, b, k, copies = 1_000_000, 768, 1024, 100, 64
g = torch.Generator(device=dev).manual_seed(3)
base = torch.randn(n // copies, d, generator=g, device=dev)
q = torch.randn(b, d, generator=g, device=dev)
idx = engine.DenseIndex(d, "f16", "COSINE", dev, capacity=n)
idx.add(base.repeat(copies, 1)[:n])
for _ in range(3): idx.search(q, k)
torch.cuda.synchronize()
