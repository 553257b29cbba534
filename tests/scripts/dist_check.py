"""Multi-GPU parity check (run under torchrun on a GPU box): the row-sharded search + NCCL all-gather + merge kernel must
return exactly what the CPU oracle returns for the whole corpus, on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/scripts/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]
from b200rag import distributed as bdist  # noqa: E402
from oracle import oracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (n, d, b, k, dt) in ((200_003, 128, 64, 100, "f16"), (50_000, 256, 200, 10, "bf16")):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[7] = x[n - 2]                                          # cross-shard exact tie
    q = rng.standard_normal((b, d)).astype(np.float32)
    idx = bdist.ShardedDenseIndex(d, n, dt, "COSINE", dev)
    idx.local.add(torch.from_numpy(x[idx.start:idx.end]))
    s, i = idx.search(torch.from_numpy(q), k)
    code = oracle.F16 if dt == "f16" else oracle.BF16
    rs, ri = oracle.dense_topk(oracle.normalize_rows(x, code), oracle.normalize_rows(q, code), k, code)
    good = np.array_equal(i.cpu().numpy(), ri) and np.array_equal(s.cpu().numpy(), rs)
    print(f"rank {rank}/{world}: n={n} d={d} b={b} k={k} {dt}: {'bit-exact vs oracle' if good else 'MISMATCH'}", flush=True)
    ok &= good
t = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(t)
dist.destroy_process_group()
sys.exit(int(t.item() != 0))
