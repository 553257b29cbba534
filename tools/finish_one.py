import sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import torch
from b200rag import engine as eng, _lib
dev = "cuda:0"
rows, dim, dt, b, k, fv = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
g = torch.Generator(device=dev); g.manual_seed(1)
idx = eng.DenseIndex(dim, dt, "COSINE", dev)
for s in range(0, rows, 250_000):
    idx.add(torch.randn(min(250_000, rows - s), dim, generator=g, device=dev))
q = torch.randn(b, dim, generator=g, device=dev)
_lib.set_option("finish_version", fv)
for _ in range(3): idx.search(q, k)
torch.cuda.synchronize()
