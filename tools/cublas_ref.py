"""What does cuBLAS sustain on the SAME operands and shape as the dense scan?  (context for the roofline fraction)

The scan is Q[B,D] x X[N,D]^T with a fused top-k; this times the plain GEMM part through torch.matmul (cuBLAS) on the
same normalised fp16 / bf16 rows, chunked over N so the [B, chunk] score block stays small, back to back for a few
seconds so the board reaches its power-capped steady state.  MEASURED_PEAKS.json's 1376 TFLOP/s is bf16 8192^3 randn.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--chunk", type=int, default=65536)
ap.add_argument("--seconds", type=float, default=3.0)
args = ap.parse_args()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
for dt in (torch.float16, torch.bfloat16):
    x = torch.empty(args.rows, args.dim, dtype=dt, device=dev)
    for s in range(0, args.rows, 250_000):
        r = torch.randn(min(250_000, args.rows - s), args.dim, generator=g, device=dev)
        x[s: s + r.shape[0]] = (r / r.norm(dim=1, keepdim=True)).to(dt)
    q = torch.randn(args.batch, args.dim, generator=g, device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).to(dt)
    out = torch.empty(args.batch, args.chunk, dtype=dt, device=dev)
    nchunks = args.rows // args.chunk

    def one_pass():
        for c in range(nchunks):
            torch.matmul(q, x[c * args.chunk: (c + 1) * args.chunk].T, out=out)

    one_pass()
    torch.cuda.synchronize()
    best = 1e9
    t_end = time.time() + args.seconds
    times = []
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_pass()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    flops = 2.0 * args.batch * nchunks * args.chunk * args.dim
    times.sort()
    print(f"cuBLAS {str(dt).split('.')[-1]:9s} [{args.batch}x{args.dim}] x [{nchunks * args.chunk}x{args.dim}]^T: "
          f"best {flops / times[0] / 1e9:.0f} TFLOP/s, median {flops / times[len(times) // 2] / 1e9:.0f} TFLOP/s, "
          f"last {flops / times[-1] / 1e9:.0f} ({len(times)} passes)")
    del x
