"""A/B of the finish stage generations (option finish_version: 0 auto, 1 block-wide, 2 warp-select monolithic, 3 split):
results must be bit-identical; times are whole-search CUDA-event means on one resident index per shape."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import torch
from b200rag import engine as eng, _lib
dev = "cuda:0"
for (rows, dim, dt, b, k) in ((1_250_000, 768, "f16", 1024, 100), (1_000_000, 1024, "bf16", 256, 500), (1_000_000, 1024, "bf16", 256, 200),
                              (300_000, 384, "f16", 64, 10), (200_000, 128, "f16", 1024, 1000)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    idx = eng.DenseIndex(dim, dt, "COSINE", dev)
    for s in range(0, rows, 250_000):
        idx.add(torch.randn(min(250_000, rows - s), dim, generator=g, device=dev))
    q = torch.randn(b, dim, generator=g, device=dev)
    ref = None
    for fv in (0, 1, 2, 3):
        _lib.set_option("finish_version", fv)
        err = torch.zeros(b, dtype=torch.float32, device=dev)
        for _ in range(3): out = idx.search(q, k, out_err=err)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): idx.search(q, k)
        e1.record(); torch.cuda.synchronize()
        res = (out[0].clone(), out[1].clone(), out[2].clone(), err.clone())
        if ref is None: ref = res
        same = all(torch.equal(a, c) for a, c in zip(ref[:3], res[:3]))
        print(f"rows {rows} dim {dim} B {b} k {k} finish_version {fv}: {e0.elapsed_time(e1)/10:.3f} ms  same_as_auto {same}  err_max {float(res[3].max()):.3e} flags {int(res[2].sum())}", flush=True)
    _lib.set_option("finish_version", -1)
    del idx
