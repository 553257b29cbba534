"""Where does a batch-1 (and small-batch) search spend its time?  Wall-clock p50 of the public call, the full-scan kernel
alone (CUDA events through the profile hook), and the HBM floor."""
import argparse
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]
from b200rag import _lib, engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--batches", default="1,2,4,8,16,32,64,128")
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--warm-s", type=float, default=1.5)
args = ap.parse_args()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
idx = engine.DenseIndex(args.dim, "f16", "COSINE", dev, capacity=args.rows)
for s in range(0, args.rows, 250_000):
    idx.add(torch.randn(min(250_000, args.rows - s), args.dim, generator=g, device=dev))
lib = _lib.load()
floor = args.rows * args.dim * 2 / 6545e9 * 1e3
print(f"rows={args.rows} dim={args.dim} k={args.k}; HBM floor {floor:.3f} ms (6545 GB/s)")
qw = torch.randn(128, args.dim, generator=g, device=dev)
for b in [int(x) for x in args.batches.split(",")]:
    qs = [torch.randn(b, args.dim, generator=g, device=dev).cpu().pin_memory() for _ in range(8)]
    t_end = time.time() + args.warm_s          # bring the clocks out of idle before measuring
    while time.time() < t_end:
        idx.search(qw, args.k)
        torch.cuda.synchronize()
    wall, scan, dev_ms = [], [], []
    for it in range(args.reps + 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e1.record()
        torch.cuda.synchronize()
        lib.b200rag_profile_next_scan(e0.cuda_event, e1.cuda_event)
        t0 = time.perf_counter()
        d0.record()
        s_, i_, f_ = idx.search(qs[it % 8].to(dev, non_blocking=True), args.k)
        d1.record()
        ih = i_.cpu()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        if it >= 5:
            wall.append((t1 - t0) * 1e3)
            scan.append(e0.elapsed_time(e1))
            dev_ms.append(d0.elapsed_time(d1))
    w, sc, dm = statistics.median(wall), statistics.median(scan), statistics.median(dev_ms)
    print(f"B={b:4d}: wall p50 {w:6.3f} ms | device span {dm:6.3f} ms | full-scan kernel {sc:6.3f} ms = "
          f"{args.rows * args.dim * 2 / sc / 1e6:6.0f} GB/s = {100 * floor / sc:4.1f}% of HBM peak | e2e {100 * floor / w:4.1f}% of floor")
