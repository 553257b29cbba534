"""Where a sharded dense step goes (torchrun, any N): local search into the send buffer / all-gather / merge, CUDA events."""
import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import torch, torch.distributed as dist
from b200rag import engine as eng
from b200rag.distributed import SendBuffer
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
b, k, dim = 1024, 100, 768
g = torch.Generator(device=dev); g.manual_seed(rank)
idx = eng.DenseIndex(dim, "f16", "COSINE", dev, id_offset=rank * rows)
for s in range(0, rows, 250_000):
    idx.add(torch.randn(min(250_000, rows - s), dim, generator=g, device=dev))
q = torch.randn(b, dim, generator=g, device=dev)
send = SendBuffer(dev)
def step(ev=None):
    msg, ps, pi = send.planes(b, k)
    if ev: ev[0].record()
    idx.search(q, k, out=(ps, pi))
    if ev: ev[1].record()
    if world > 1:
        out = torch.empty((world * 2 * b, k), dtype=msg.dtype, device=dev)
        dist.all_gather_into_tensor(out, msg.view(2 * b, k))
        if ev: ev[2].record()
        r = eng.merge_gathered(out.view(world, 2, b, k), k)
    elif ev: ev[2].record()
    if ev: ev[3].record()
for _ in range(5): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
acc = [0.0, 0.0, 0.0]
n = 20
for _ in range(n):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    step(ev); torch.cuda.synchronize()
    for j in range(3): acc[j] += ev[j].elapsed_time(ev[j + 1])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): step()
e1.record(); torch.cuda.synchronize()
print(f"rank {rank}/{world} rows {rows}: local {acc[0]/n:.3f} ms, all_gather {acc[1]/n:.3f} ms, merge {acc[2]/n:.3f} ms; back-to-back step {e0.elapsed_time(e1)/n:.3f} ms", flush=True)
# local search pieces
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
q16 = idx.prepare_queries(q)
e[0].record()
for _ in range(n): idx.search_prepared(q16, k)
e[1].record(); torch.cuda.synchronize()
print(f"rank {rank}: search_prepared {e[0].elapsed_time(e[1])/n:.3f} ms", flush=True)
if world > 1: dist.destroy_process_group()
