#!/usr/bin/env python
"""bench.py -- headline benchmark of the dense hot path (BASELINE.json: "QPS @10Mx768 top-100 batch 1024").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one batch of 1024 queries through the exact cosine top-100 search over a 10M x 768 fp16 corpus
(BASELINE config 3; it fits one B200).  With N GPUs the SAME 10M rows are sharded row-wise (strong scaling): each
rank scans its shard, one NCCL all-gather carries the k candidates per query, the merge kernel reduces them.

  value     whole-job queries/s with the query batch already resident in HBM (fp32), search = prepare + scan + finish
            (+ all-gather + merge for N>1), timed with CUDA events, max over ranks.
  e2e       same metric through the public Python API with HOST buffers: pinned fp32 queries -> H2D -> search ->
            D2H of ids + scores, every step, consumed by a double-buffered host loop (results of step i are on the host
            before step i+2 is issued).
  roofline  the dominant kernel (dense_scan3_kernel, the full scan): algorithmic FLOPs per launch / its CUDA-event duration, measured
            live inside the timed region through the b200rag_profile_next_scan hook.
  cpu_baseline / --impl reference: the same exact search on the box's host cores (numpy BLAS sgemm + partial sort,
            "Milvus mocked" -- see oracle/oracle.py) on a bounded row sample, extrapolated linearly in rows.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "advanced-rag-milvus_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "QPS @10Mx768 top-100 batch 1024"
BLOCK_ROWS = 65536     # corpus generation granularity: row block b depends only on (seed, b) -> sharding independent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-sample-rows", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"],
                "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region, in-process through NVML (pynvml).

    (Spawning `nvidia-smi -lms` next to the timed loop costs several ms per step while it initialises; an NVML handle
    opened before the warm-up does not disturb the GPU.)"""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("hw_power_brake_slowdown", 0x80), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int, period_s: float = 0.02):
        self.period = period_s
        self.samples, self.reasons, self.power = [], set(), []
        self._stop = threading.Event()
        self._thread = None
        self.h = None
        self.smax = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                      # noqa: BLE001 - NVML missing is reported, not fatal
            self.err = repr(e)
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for name, bit in self.REASONS:
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:                       # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.h is None:
            return
        self.samples, self.reasons, self.power = [], set(), []
        self._stop.clear()
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self._stop.set()
        self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w": statistics.median(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------------- reference / CPU arm
def cpu_dense_qps(rows_total: int, dim: int, batch: int, k: int, sample_rows: int, steps: int, warmup: int, seed: int):
    """The exact cosine top-k on the host cores: numpy BLAS sgemm + argpartition over a row sample (fp32, what a CPU
    deployment of the mocked-Milvus path runs), extrapolated linearly to rows_total."""
    import numpy as np
    from oracle import oracle
    rng = np.random.default_rng(seed)
    sample_rows = min(sample_rows, rows_total)
    x = rng.standard_normal((sample_rows, dim), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((batch, dim), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.dense_topk_blas(x, q, k)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    qps = batch / (t * rows_total / sample_rows)
    return qps, t, sample_rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 5))
    qps, t, srows = cpu_dense_qps(args.rows, args.dim, args.batch, args.k, args.cpu_sample_rows, steps, min(args.warmup, 1), args.seed)
    sample = (f"numpy BLAS sgemm + argpartition, {args.batch} queries x {srows} of {args.rows} rows per step "
              f"({t:.2f} s/step), QPS extrapolated linearly in rows")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3 * args.rows / srows, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.rows}x{args.dim} exact cosine top-{args.k}, query batch {args.batch}, host CPU"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- B200 arm
def build_shard(engine, dist_mod, args, device, rank, world):
    import torch
    start, end = dist_mod.shard_range(args.rows, rank, world)
    idx = engine.DenseIndex(args.dim, "f16", "COSINE", device, id_offset=start, capacity=end - start)
    g = torch.Generator(device=device)
    row = start
    while row < end:
        blk = row // BLOCK_ROWS
        b0 = blk * BLOCK_ROWS
        b1 = min(b0 + BLOCK_ROWS, args.rows)
        g.manual_seed(args.seed * 1_000_003 + blk)
        x = torch.randn((b1 - b0, args.dim), generator=g, device=device, dtype=torch.float32)
        lo, hi = max(row, b0) - b0, min(end, b1) - b0
        idx.add(x[lo:hi])
        row = b0 + hi
    return idx


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from b200rag import _lib, distributed as bdist, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    B, K, D = args.batch, args.k, args.dim

    idx = build_shard(engine, bdist, args, device, rank, world)
    n_local = idx.n
    g = torch.Generator(device=device).manual_seed(args.seed + 1000)
    POOL = 8
    q_dev = [torch.randn((B, D), generator=g, device=device, dtype=torch.float32) for _ in range(POOL)]
    q_host = [q.cpu().pin_memory() for q in q_dev]

    # N > 1: pack -> ONE NCCL all-gather -> merge kernel, on the compute stream.  (Running the exchange of batch b on a side
    # stream next to the scan of batch b+1 was measured and is slower: the persistent scan wants every SM pair at launch and
    # the NCCL / merge CTAs delay some of its clusters -- scan 1.45 -> 1.98 ms at 8 GPUs.)
    def search(q):
        s, i, f = idx.search(q, K, engine.DENSE_AUTO)
        if world > 1:
            s, i = bdist.gather_and_merge(s, i, K, engine.merge_topk, None, engine.merge_gathered)
        return s, i, f

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    flags_total = torch.zeros((), dtype=torch.int64, device=device)
    for it in range(args.warmup):
        s, i, f = search(q_dev[it % POOL])
        flags_total += f.sum()        # also warms torch's own reduce/add kernels (first use costs ~100 ms of module load)
    barrier()

    # ---------------- timed region 1: device-resident inputs -------------------------------------
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in scan_ev:          # torch creates the cudaEvent lazily on the first record(); the hook needs live handles
        a.record()
        b.record()
    flags_total.zero_()
    barrier()
    launches0 = int(lib.b200rag_kernel_launch_count())
    ev0.record()
    for it in range(args.steps):
        lib.b200rag_profile_next_scan(scan_ev[it][0].cuda_event, scan_ev[it][1].cuda_event)
        s, i, f = search(q_dev[(args.warmup + it) % POOL])
        flags_total += f.sum()
    ev1.record()
    barrier()
    launches = int(lib.b200rag_kernel_launch_count()) - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop()
    scan_ms = [a.elapsed_time(b) for a, b in scan_ev]
    scan_ms_avg = max_over_ranks(sum(scan_ms) / len(scan_ms))
    ms_per_step = ms_total / args.steps
    value = B * args.steps / (ms_total * 1e-3)

    # ---------------- timed region 2: end to end from host buffers -------------------------------
    # Every step: pinned fp32 queries -> H2D -> search -> D2H of ids + scores into pinned host buffers.  The consumer is
    # double buffered, as a serving loop would be: before issuing step i+1 the host WAITS until the results of step i-1
    # are in host memory (so it holds every step's results at most one step late), and the region ends when the last
    # step's results have landed.  Nothing is skipped; copies and compute of neighbouring steps may overlap.
    out_ids_h = [torch.empty((B, K), dtype=torch.int64).pin_memory() for _ in range(2)]
    out_sc_h = [torch.empty((B, K), dtype=torch.float64).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    for it in range(min(args.warmup, 2)):
        search(q_host[it % POOL].to(device, non_blocking=True))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    checksum = 0
    for it in range(args.steps):
        qd = q_host[(args.warmup + it) % POOL].to(device, non_blocking=True)
        s, i, _ = search(qd)
        if it >= 2:
            done[it % 2].synchronize()                    # results of step it-2 left this buffer pair long ago; it-1 may still fly
        out_ids_h[it % 2].copy_(i, non_blocking=True)
        out_sc_h[it % 2].copy_(s, non_blocking=True)
        done[it % 2].record()
        if it >= 1:
            done[(it - 1) % 2].synchronize()              # the caller now holds step it-1's results on the host
            checksum += int(out_ids_h[(it - 1) % 2][0, 0])
    done[(args.steps - 1) % 2].synchronize()
    checksum += int(out_ids_h[(args.steps - 1) % 2][0, 0])
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = B * args.steps / (e2e_ms * 1e-3)

    # ---------------- extras: batch-1 latency ----------------------------------------------------
    extras = {}
    if not args.no_extras:
        lat = []
        q1 = [q_host[j % POOL][j: j + 1].clone().pin_memory() for j in range(60)]
        for j in range(60):
            barrier()
            t0 = time.perf_counter()
            s, i, _ = search(q1[j].to(device, non_blocking=True))
            i_h = i.cpu()
            t1 = time.perf_counter()
            if j >= 10:
                lat.append((t1 - t0) * 1e3)
        lat_t = torch.tensor([statistics.median(lat)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)
        extras["batch1_p50_ms"] = float(lat_t.item())
        extras["batch1_hbm_floor_ms"] = n_local * D * 2 / (load_peaks()["hbm_gbs"] * 1e9) * 1e3
        # in-run exactness check (outside every timed region): the tensor-core path against the CUDA-core exact scan
        # (canonical fp64 arithmetic, itself bit-exact against the CPU oracle in tests/) on a query subset of this very index
        sub = torch.tensor([0, 1, 127, 128, 500, B - 1], device=device).clamp(max=B - 1).unique()
        qs = q_dev[0][sub]
        sa, ia, fa = idx.search(qs, K, engine.DENSE_AUTO)
        sx, ix, _ = idx.search(qs, K, engine.DENSE_EXACT)
        ok = bool(torch.equal(ia, ix) and torch.equal(sa, sx))
        ok_t = torch.tensor([1 if ok else 0], device=device)
        if world > 1:
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        extras["exact_check"] = {"queries": int(sub.numel()), "ids_and_fp64_scores_equal_to_exact_scan": bool(ok_t.item()),
                                 "recall_at_k": float((ia == ix).float().mean().item()), "flagged": int(fa.sum().item())}

    if rank == 0:
        peaks = load_peaks()
        flops = 2.0 * B * n_local * D                      # per scan launch on one rank
        bytes_alg = n_local * D * 2 + B * D * 2 + B * K * 12
        achieved_tf = flops / (scan_ms_avg * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "scan_traffic.json")
        if os.path.exists(tpath):
            try:
                with open(tpath) as fh:
                    tj = json.load(fh)
                # the ncu capture is of ONE workload; report it only for the same per-GPU shard and batch
                if (tj.get("rows"), tj.get("dim"), tj.get("batch")) == (n_local, D, B):
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        scan_kernel = "dense_scan_kernel" if os.environ.get("B200RAG_SCAN_VERSION", "") == "1" else "dense_scan3_kernel"
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"{args.rows}x{D} fp16 exact cosine top-{K}, query batch {B}, row-sharded over {world} GPU(s)",
                       "l2": f"inputs larger than L2 ({n_local * D * 2 / 1e9:.1f} GB per GPU streamed every step)",
                       "mode": "AUTO (tcgen05 scan + exact fp64 re-score, exact fallback for unproven queries)",
                       "flagged_queries_in_timed_region": int(flags_total.item())},
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * K * 16,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": scan_kernel, "bound": "tensor", "achieved": achieved_tf,
                         "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved_tf / peaks["tflops_sustained"],
                         "peak_kind": f"bf16 cuBLAS sustained, {peaks['source']} (burst {peaks['tflops_burst']})",
                         "frac_of_burst": achieved_tf / peaks["tflops_burst"],
                         "ms_per_launch": scan_ms_avg, "flops_per_launch": flops, "algorithmic_bytes_per_launch": bytes_alg,
                         "hbm_frac": bytes_alg / (scan_ms_avg * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "scan_share_of_step": scan_ms_avg / ms_per_step, "traffic": traffic},
        }
        if extras:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            qps, t, srows = cpu_dense_qps(args.rows, D, B, K, args.cpu_sample_rows, 5, 1, args.seed)     # ~12 s of host work
            line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                    "sample": f"numpy BLAS sgemm + argpartition, {B} queries x {srows} of {args.rows} rows "
                                              f"({t:.2f} s per pass, 5 timed passes), extrapolated linearly in rows"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
