"""Why is bench.py's device-resident region slower per step than its end-to-end region?  A/B the suspects."""
import os, sys, time, threading
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")]
import bench
from b200rag import _lib, engine

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.load()
N, D, B, K = int(os.environ.get("ROWS", 10_000_000)), 768, 1024, 100
idx = engine.DenseIndex(D, "f16", "COSINE", dev, capacity=N)
g = torch.Generator(device=dev).manual_seed(0)
for s in range(0, N, 250_000):
    idx.add(torch.randn(min(250_000, N - s), D, generator=g, device=dev))
qs = [torch.randn(B, D, generator=g, device=dev) for _ in range(8)]
for i in range(3):
    idx.search(qs[i], K)
torch.cuda.synchronize()

def run(label, steps=10, sampler=False, hook=False, flagsum=False, period=0.02):
    smp = bench.ClockSampler(0, period) if sampler else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record(); b.record()
    tot = torch.zeros((), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    if smp: smp.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    walls = []
    for it in range(steps):
        w0 = time.perf_counter()
        if hook:
            lib.b200rag_profile_next_scan(evs[it][0].cuda_event, evs[it][1].cuda_event)
        s, i, f = idx.search(qs[it % 8], K)
        if flagsum:
            tot += f.sum()
        walls.append((time.perf_counter() - w0) * 1e3)
    e1.record()
    torch.cuda.synchronize()
    c = smp.stop() if smp else None
    print(f"{label:34s} {e0.elapsed_time(e1) / steps:7.3f} ms/step  walls: " + " ".join(f"{w:5.1f}" for w in walls), c or "")

run("plain")
run("plain again")
run("hook", hook=True)
run("flagsum", flagsum=True)
run("sampler 20ms", sampler=True)
run("sampler 100ms", sampler=True, period=0.1)
run("all", sampler=True, hook=True, flagsum=True)
run("plain 30 steps", steps=30)
