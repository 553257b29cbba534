"""Per-role cycle breakdown of the tensor-core scan (b200rag_debug_scan_stats): where does each warp role wait?"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]
from b200rag import _lib, engine  # noqa: E402

NAMES = ["mma_total", "mma_wait_full", "mma_wait_tempty", "mma_wait_q", "prod_total", "prod_wait_empty", "epi_total",
         "epi_wait_tfull", "epi_compact", "epi_qload", "epi_ncompact", "epi_nslow"]

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--dtype", default="f16")
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
idx = engine.DenseIndex(args.dim, args.dtype, "COSINE", dev, capacity=args.rows)
for s in range(0, args.rows, 250_000):
    idx.add(torch.randn(min(250_000, args.rows - s), args.dim, generator=g, device=dev))
q = torch.randn(args.batch, args.dim, generator=g, device=dev)
lib = _lib.load()
for _ in range(3 + args.reps):
    idx.search(q, args.k, engine.DENSE_TENSOR)
torch.cuda.synchronize()
lib.b200rag_debug_scan_stats(1, None, 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); e1.record()
lib.b200rag_profile_next_scan(e0.cuda_event, e1.cuda_event)
idx.search(q, args.k, engine.DENSE_TENSOR)
torch.cuda.synchronize()
buf = np.zeros((256, 16), dtype=np.uint64)
lib.b200rag_debug_scan_stats(0, buf.ctypes.data_as(ctypes.c_void_p), 256)
used = buf[buf[:, 0] > 0]
ms = e0.elapsed_time(e1)
print(f"v={os.environ.get('B200RAG_SCAN_VERSION')} cs={os.environ.get('B200RAG_CLUSTER')} {args.dtype} rows={args.rows} dim={args.dim} "
      f"B={args.batch}: scan {ms:.2f} ms, {2.0 * args.batch * args.rows * args.dim / ms / 1e9:.0f} TFLOP/s, CTAs={len(used)}")
mean = used.astype(np.float64).mean(0)
for i, nm in enumerate(NAMES):
    extra = ""
    if nm.startswith("mma_") and i:
        extra = f"  ({100 * mean[i] / mean[0]:.1f}% of mma_total)"
    if nm.startswith("prod_w"):
        extra = f"  ({100 * mean[i] / mean[4]:.1f}% of prod_total)"
    if nm.startswith("epi_") and i > 6 and i < 10:
        extra = f"  ({100 * mean[i] / mean[6]:.1f}% of epi_total)"
    print(f"  {nm:16s} {mean[i]:14.0f}{extra}")
