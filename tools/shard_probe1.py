import sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import torch
from b200rag import engine as eng
dev = "cuda:0"
rows, k = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.randn(rows, 1024, generator=g, device=dev)
idx = eng.DenseIndex(1024, "bf16", "COSINE", dev); idx.add(x)
q = torch.randn(256, 1024, generator=g, device=dev)
for _ in range(3): idx.search(q, k)
torch.cuda.synchronize()
