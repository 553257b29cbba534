"""Does the sharded search chain (local scan -> all-gather -> merge) replay from a CUDA graph?  torchrun, N >= 2."""
import os, sys, time, statistics
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import numpy as np, torch, torch.distributed as dist
from b200rag import distributed as bdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = f"cuda:{int(os.environ['LOCAL_RANK'])}"
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=torch.device(dev))
rows_total, dim, k = 1_250_000 * world, 768, 100
res = {}
for graphs in (False, True):
    mgr = bdist.ShardedIndexManager(rows_total, semantic_dim=dim, sparse_dim=8, domain_dim=8, device=dev, dtype="f16",
                                    enable_sparse=False, use_graphs=graphs)
    g = torch.Generator(device=dev); g.manual_seed(rank)
    for s in range(mgr.start, mgr.end, 250_000):
        mgr.add_vectors(torch.randn(min(250_000, mgr.end - s), dim, generator=g, device=dev))
    gq = torch.Generator(); gq.manual_seed(5)
    for b in (1, 1024):
        qs = [torch.randn(b, dim, generator=gq).pin_memory() for _ in range(8)]
        outs, lat = [], []
        for j in range(40):
            dist.barrier()
            t0 = time.perf_counter()
            arr = mgr.search_batch_arrays(qs[j % 8], "semantic_index", k)
            lat.append((time.perf_counter() - t0) * 1e3)
            if j >= 32: outs.append((arr.rows.copy(), arr.scores.copy()))
        res[(graphs, b)] = outs
        print(f"rank {rank} graphs {graphs} batch {b}: p50 {statistics.median(lat[10:]):.3f} ms", flush=True)
    del mgr
for b in (1, 1024):
    same = all(np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]) for a, c in zip(res[(False, b)], res[(True, b)]))
    print(f"rank {rank} batch {b}: graph replay == eager: {same}", flush=True)
dist.destroy_process_group()
