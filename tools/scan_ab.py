"""A/B the tensor-core scan generations in ONE process on one resident index, in steady state (power-capped clocks).

For each (scan_version, epi, qg_span, sample_mult) option setting (b200rag_set_option): `reps` back-to-back searches; the b200rag_profile_next_scan hook
times the FULL scan kernel alone, CUDA events time the whole search (prepare + sample pass + scan + finish), and the
per-role cycle counters of the last scan give the effective SM clock (mma_total cycles / scan time).
"""
import argparse
import ctypes
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]
from b200rag import _lib, engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--dtype", default="f16")
ap.add_argument("--reps", type=int, default=60)
ap.add_argument("--configs", default="3:0:0,3:0:1,3:0:2,1:0:0,3:0:0", help="version:unused:span, comma separated")
ap.add_argument("--mode", default="auto", choices=["auto", "tensor"])
args = ap.parse_args()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
idx = engine.DenseIndex(args.dim, args.dtype, "COSINE", dev, capacity=args.rows)
for s in range(0, args.rows, 250_000):
    idx.add(torch.randn(min(250_000, args.rows - s), args.dim, generator=g, device=dev))
qs = [torch.randn(args.batch, args.dim, generator=g, device=dev) for _ in range(8)]
lib = _lib.load()
mode = engine.DENSE_AUTO if args.mode == "auto" else engine.DENSE_TENSOR
flops = 2.0 * args.batch * args.rows * args.dim
print(f"{args.dtype} rows={args.rows} dim={args.dim} B={args.batch} k={args.k} reps={args.reps} mode={args.mode}")
for cfg in args.configs.split(","):
    ver, cs, span, mult = (cfg.split(":") + ["0", "8"])[:4] if cfg.count(":") >= 3 else (cfg.split(":") + ["0"])[:3] + ["8"]
    _lib.set_option("sample_mult", int(mult))           # fourth field: expected rows above the sampled threshold, in k'
    _lib.set_option("epi", int(cs))                     # second field: epilogue variant (0 = compare chain, 1 = sub-group maxima)
    _lib.set_option("scan_version", int(ver))
    _lib.set_option("qg_span", int(span) if int(span) > 0 else -1)
    stats = torch.zeros((256, 16), dtype=torch.int64, device=dev)
    for i in range(3):
        idx.search(qs[i], args.k, mode)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
    tot = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
    for a, b in evs:
        a.record(); b.record()
    torch.cuda.synchronize()
    flagged = 0
    for it in range(args.reps):
        if it == args.reps - 1:
            lib.b200rag_debug_set_stats_buffer(0, stats.data_ptr(), stats.numel())
        lib.b200rag_profile_next_scan(evs[it][0].cuda_event, evs[it][1].cuda_event)
        tot[it][0].record()
        s_, i_, f_ = idx.search(qs[it % 8], args.k, mode)
        tot[it][1].record()
    torch.cuda.synchronize()
    flagged = int(f_.sum())
    lib.b200rag_debug_set_stats_buffer(0, None, 0)
    buf = stats.cpu().numpy()
    used = buf[buf[:, 0] > 0].astype(np.float64)
    scan = [a.elapsed_time(b) for a, b in evs][5:]
    step = [a.elapsed_time(b) for a, b in tot][5:]
    sm, st = statistics.median(scan), statistics.median(step)
    last_scan = evs[-1][0].elapsed_time(evs[-1][1])
    mma_total = used[:, 0].mean() if len(used) else float("nan")
    mma_busy = 1.0 - (used[:, 1].mean() + used[:, 2].mean() + used[:, 3].mean()) / mma_total if len(used) else float("nan")
    print(f"v{ver} epi={cs} span={span} mult={mult}: scan median {sm:6.2f} ms = {flops / sm / 1e9:5.0f} TFLOP/s (min {min(scan):.2f} max {max(scan):.2f}); "
          f"search {st:6.2f} ms = {args.batch / st * 1e3:7.0f} QPS; clock ~{mma_total / last_scan / 1e6:.3f} GHz, "
          f"MMA issue busy {100 * mma_busy:.0f}%, CTAs with MMA {len(used)}, flagged(last) {flagged}")
for name in ("sample_mult", "epi", "scan_version", "qg_span"):
    _lib.set_option(name, -1)
