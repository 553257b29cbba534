#!/usr/bin/env python
"""bench.py -- headline benchmark of the dense hot path (BASELINE.json: "QPS @10Mx768 top-100 batch 1024").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one batch of 1024 queries through the exact cosine top-100 search over a 10M x 768 fp16 corpus
(BASELINE config 3; it fits one B200).  With N GPUs the SAME 10M rows are sharded row-wise (strong scaling): each
rank scans its shard, one NCCL all-gather carries the k candidates per query, the merge kernel reduces them.
Every number goes through the reference-facing plugin object -- B200IndexManager (N = 1) / its row-sharded subclass
(N > 1), the drop-in for MilvusIndexManager (reference src/advanced_rag/indexing.py:445-551):

  value     whole-job queries/s with the query batch already resident in HBM (fp32): manager.search_batch_ids = prepare +
            scan + finish (+ all-gather + merge for N>1), results stay on the device; CUDA events, max over ranks.
  e2e       the same metric through manager.search_batch_arrays with HOST buffers: pinned fp32 queries -> H2D -> search ->
            D2H of ids + scores + counts into numpy arrays the caller holds when the call returns, every step, one
            synchronous call per step (no overlap between steps).
  roofline  the dominant kernel (dense_scan3_kernel, the full scan): algorithmic FLOPs per launch / its CUDA-event duration,
            measured live inside the timed region through the b200rag_profile_next_scan hook.
  cpu_baseline / --impl reference: the same exact search on the box's host cores (numpy BLAS sgemm + partial sort,
            "Milvus mocked" -- see oracle/oracle.py) with every host thread, on a bounded row sample per step; `steps` and
            `ms_per_step` are what was really run and measured, `value` extrapolates linearly in rows to the 10M-row workload.
  extras    batch-1 latency, result-dict materialisation cost, in-run exactness checks (merged result for N>1), and the
            other BASELINE configurations (bench_extras.py): c1 with the reference chain on the host cores, c2, c4 stage
            table, c5 shard, near-duplicate corpus.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
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
    ap.add_argument("--extras", default="c1,c2,c4,c5,dups,ingest", help="comma list of secondary configurations to run (N = 1)")
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
def use_all_host_threads() -> int:
    """BLAS / OpenMP on every host core, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers).  Must run
    before numpy is imported for the environment part; threadpoolctl raises the limit of pools that are already loaded."""
    cores = os.cpu_count() or 1
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(cores)
    return cores


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
        info = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") in ("blas", "openmp")]
        return max(info) if info else 1
    except Exception:                               # noqa: BLE001
        return int(os.environ.get("OMP_NUM_THREADS", "1"))


def cpu_dense_sample(rows_total: int, dim: int, batch: int, k: int, sample_rows: int, steps: int, warmup: int, seed: int):
    """The exact cosine top-k on the host cores: numpy BLAS sgemm + argpartition over a row sample (fp32, what a CPU
    deployment of the mocked-Milvus path runs).  Returns (seconds per sample step, sample rows, threads used)."""
    import numpy as np
    from oracle import oracle
    threads = blas_threads()
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
    return sum(times) / len(times), sample_rows, threads


def cpu_baseline_block(args, steps: int, warmup: int):
    t, srows, threads = cpu_dense_sample(args.rows, args.dim, args.batch, args.k, args.cpu_sample_rows, steps, warmup, args.seed)
    qps = args.batch / (t * args.rows / srows)
    return {"value": qps, "unit": "queries/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
            "sample": f"numpy BLAS sgemm + argpartition on {threads} threads: {steps} timed steps of {args.batch} queries x "
                      f"{srows} of the {args.rows} rows ({t * 1e3:.0f} ms measured per step), QPS extrapolated linearly in rows",
            "sample_ms_per_step": t * 1e3, "sample_rows": srows}, t, srows


def run_reference(args):
    """--impl reference: the CPU implementation of the path on this box's host cores.  Rank 0 only; the line reports what was
    actually run (`steps` timed steps of `ms_per_step` each, over a row sample) and extrapolates only `value`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 2))
    cb, t, srows = cpu_baseline_block(args, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "queries/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.rows}x{args.dim} exact cosine top-{args.k}, query batch {args.batch}, host CPU; each timed step "
                               f"scans a {srows}-row sample ({srows / args.rows:.3f} of the workload) and `value` = batch / "
                               f"(ms_per_step * rows / sample_rows)",
                   "sample_rows": srows, "rows": args.rows},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- B200 arm
def build_manager(args, device, rank, world):
    """The 10M x 768 corpus behind the plugin object: payload-less bulk rows (ids read f"c{row:09d}")."""
    import torch
    from b200rag import distributed as bdist
    from b200rag.index_manager import B200IndexManager
    kw = dict(semantic_dim=args.dim, sparse_dim=8, domain_dim=8, device=device, dtype="f16", enable_sparse=False)
    if world > 1:
        mgr = bdist.ShardedIndexManager(args.rows, **kw)
        start, end = mgr.start, mgr.end
    else:
        mgr = B200IndexManager(**kw)
        start, end = 0, args.rows
    g = torch.Generator(device=device)
    row = start
    while row < end:
        blk = row // BLOCK_ROWS
        b0 = blk * BLOCK_ROWS
        b1 = min(b0 + BLOCK_ROWS, args.rows)
        g.manual_seed(args.seed * 1_000_003 + blk)
        x = torch.randn((b1 - b0, args.dim), generator=g, device=device, dtype=torch.float32)
        lo, hi = max(row, b0) - b0, min(end, b1) - b0
        mgr.add_vectors(x[lo:hi])
        row = b0 + hi
    return mgr


def main():
    args = parse_args()
    if args.impl == "reference":
        use_all_host_threads()
        run_reference(args)
        return

    import numpy as np
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
    COL = "semantic_index"

    mgr = build_manager(args, device, rank, world)
    n_local = mgr._sem.n
    g = torch.Generator(device=device).manual_seed(args.seed + 1000)
    POOL = 8
    q_dev = [torch.randn((B, D), generator=g, device=device, dtype=torch.float32) for _ in range(POOL)]
    q_host = [q.cpu().pin_memory() for q in q_dev]

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
        s, i, c = mgr.search_batch_ids(q_dev[it % POOL], COL, K)
        flags_total += mgr._sem.last_flags.sum()   # also warms torch's own reduce/add kernels (first use costs ~100 ms of module load)
    barrier()

    # ---------------- timed region 1: device-resident inputs, results stay on the device ---------
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
        s, i, c = mgr.search_batch_ids(q_dev[(args.warmup + it) % POOL], COL, K)
        flags_total += mgr._sem.last_flags.sum()
    ev1.record()
    barrier()
    launches = int(lib.b200rag_kernel_launch_count()) - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop()
    scan_ms = [a.elapsed_time(b) for a, b in scan_ev]
    scan_ms_avg = max_over_ranks(sum(scan_ms) / len(scan_ms))
    ms_per_step = ms_total / args.steps
    value = B * args.steps / (ms_total * 1e-3)

    # ---------------- timed region 2: end to end through the columnar plugin call ----------------
    # Every step: pinned fp32 host queries -> manager.search_batch_arrays -> numpy rows / scores / counts on the host.  The
    # call is synchronous (the caller holds the step's results when it returns), so H2D, search and D2H of one step do not
    # overlap with the next; nothing is skipped or cached.
    # At N > 1 the batch enters SPMD-style: rank r brings queries [r*B/N, (r+1)*B/N) from ITS host (every rank serves its own
    # clients), the manager all-gathers them over NVLink, and each rank receives, merges and copies back the results of its own
    # queries (ShardedIndexManager.search_own_queries_arrays).  Summed over ranks the step still moves the whole batch's
    # queries H2D and the whole result D2H.  extras.e2e_replicated is the other calling convention (every rank passes the
    # whole batch and gets the whole result).
    own = world > 1 and B % world == 0
    per = B // world if own else B

    def e2e_call(j):
        if own:
            return mgr.search_own_queries_arrays(q_host[j % POOL][rank * per: (rank + 1) * per], COL, K)
        return mgr.search_batch_arrays(q_host[j % POOL], COL, K)

    def timed_e2e(call):
        for it in range(max(args.warmup, 4)):          # (the third call of a shape captures its CUDA graph)
            call(it)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        chk = 0
        t_wall0 = time.perf_counter()
        e0.record()
        for it in range(args.steps):
            arr = call(args.warmup + it)
            chk += int(arr.rows[0, 0]) + int(arr.counts[-1])
        e1.record()
        barrier()
        wall = (time.perf_counter() - t_wall0) * 1e3
        return max_over_ranks(max(e0.elapsed_time(e1), 0.0)), wall, chk

    e2e_ms, e2e_wall_ms, checksum = timed_e2e(e2e_call)
    e2e_value = B * args.steps / (e2e_ms * 1e-3)
    e2e_replicated = None
    if own and not args.no_extras:
        rep_ms, _, _ = timed_e2e(lambda j: mgr.search_batch_arrays(q_host[j % POOL], COL, K))
        e2e_replicated = {"what": "every rank passes the whole batch from its host and receives the whole result",
                          "qps": B * args.steps / (rep_ms * 1e-3), "ms_per_step": rep_ms / args.steps}

    # ---------------- extras ------------------------------------------------------------------
    extras = {}
    if e2e_replicated is not None:
        extras["e2e_replicated"] = e2e_replicated
    if not args.no_extras:
        lat = []
        q1 = [q_host[j % POOL][j: j + 1].clone().pin_memory() for j in range(60)]
        for j in range(210):                                  # 10 warm-up + 200 timed single-query calls
            barrier()
            t0 = time.perf_counter()
            arr1 = mgr.search_batch_arrays(q1[j % 60], COL, K)
            t1 = time.perf_counter()
            if j >= 10:
                lat.append((t1 - t0) * 1e3)
        lat_t = torch.tensor([statistics.median(lat)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)
        extras["batch1_p50_ms"] = float(lat_t.item())
        extras["batch1_hbm_floor_ms"] = n_local * D * 2 / (load_peaks()["hbm_gbs"] * 1e9) * 1e3
        # result-dict materialisation (reference indexing.py:534-551 shape) on top of the columnar call, top_k = 20 and K
        for kk in (20, K):
            ts = []
            for j in range(4):
                barrier()
                t0 = time.perf_counter()
                hits = mgr.search_batch(q_host[j % POOL], COL, kk)
                ts.append((time.perf_counter() - t0) * 1e3)
            extras[f"dict_results_top{kk}"] = {"ms_per_batch": statistics.median(ts[1:]), "qps": B / statistics.median(ts[1:]) * 1e3,
                                               "dicts_per_batch": sum(len(h) for h in hits)}
        # in-run exactness check (outside every timed region): what the plugin call returns (tensor-core path, merged over
        # the ranks for N > 1) against the CUDA-core exact scan (canonical fp64 arithmetic, itself bit-exact against the CPU
        # oracle in tests/) pushed through the same gather + merge, on a query subset of this very index
        sub = torch.tensor([0, 1, 127, 128, 500, B - 1], device=device).clamp(max=B - 1).unique()
        qs = q_dev[0][sub].contiguous()
        sa, ia, _ = mgr.search_batch_ids(qs, COL, K)
        fa = mgr._sem.last_flags
        sx, ix, _ = mgr._sem.search(qs, K, engine.DENSE_EXACT)
        if world > 1:
            sx, ix = bdist.gather_and_merge(sx, ix, K, engine.merge_topk, None, engine.merge_gathered)
        ok = bool(torch.equal(ia, ix) and torch.equal(sa, sx))
        ok_t = torch.tensor([1 if ok else 0], device=device)
        if world > 1:
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        extras["exact_check"] = {"queries": int(sub.numel()), "what": "plugin result (merged over ranks) == exact fp64 scan -> same gather + merge",
                                 "ids_and_fp64_scores_equal_to_exact_scan": bool(ok_t.item()),
                                 "recall_at_k": float((ia == ix).float().mean().item()), "flagged": int(fa.sum().item())}

        # approximate mode (north star: "recall@k of any approximate mode is reported against the exact scan"): rank by the
        # tensor-core fp32 scores, no fp64 re-score / proof / fallback -- on this rank's shard, against its exact result
        sa, ia, _ = mgr._sem.search(q_dev[1], K, engine.DENSE_AUTO)
        sp_, ip_, _ = mgr._sem.search(q_dev[1], K, engine.DENSE_APPROX)
        hit = (ip_.unsqueeze(2) == ia.unsqueeze(1)).any(2).float().mean()
        ta, tb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for j in range(6):                                  # (the GPU idled during the host-side sections above: clocks ramp up again)
            mgr._sem.search(q_dev[j % POOL], K, engine.DENSE_AUTO)
        torch.cuda.synchronize()
        ta.record()
        for j in range(10):
            mgr._sem.search(q_dev[j % POOL], K, engine.DENSE_APPROX)
        tb.record()
        torch.cuda.synchronize()
        extras["approx_mode"] = {"what": "B200RAG_DENSE_APPROX on this rank's shard vs its exact top-k", "recall_at_k": float(hit.item()),
                                 "same_order_fraction": float((ip_ == ia).float().mean().item()),
                                 "max_rel_score_diff": float(((sp_ - sa).abs() / sa.abs().clamp(min=1e-30)).max().item()),
                                 "ms_per_step": ta.elapsed_time(tb) / 10}

    # ---------------- secondary configurations (N = 1; they need the memory the headline index holds) ---
    if not args.no_extras and world == 1:
        import bench_extras as bx
        del mgr
        torch.cuda.empty_cache()
        want = set(args.extras.split(","))
        for key, name, fn in (("c1", "c1", lambda: bx.c1(device)),
                              ("c2", "c2", lambda: bx.dense_config(device, 1_000_000, 768, 1024, 100, "IP")),
                              ("c4", "c4", lambda: bx.c4(device)),
                              ("c5", "c5_shard", lambda: bx.dense_config(device, 12_500_000, 384, 4096, 10, "COSINE", reps=5)),
                              ("dups", "near_duplicates", lambda: bx.near_duplicates(device)),
                              ("ingest", "filters_and_ingest", lambda: bx.filters_and_ingest(device))):
            if key not in want:
                continue
            try:
                t0 = time.time()
                extras[name] = fn()
                extras[name]["wall_s"] = round(time.time() - t0, 1)
            except Exception as e:                  # noqa: BLE001 - a secondary configuration must not take the headline down
                extras[name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()

    # ---------------- secondary configurations on N > 1 GPUs: config 5 and the sharded hybrid chain ----------
    if not args.no_extras and world > 1:
        import bench_extras as bx
        del mgr
        torch.cuda.empty_cache()
        want = set(args.extras.split(","))
        for key, name, fn in (("c5", "c5_sharded", lambda: bx.c5_sharded(device, world)),
                              ("c4", "c4_sharded", lambda: bx.c4_sharded(device, world))):
            if key not in want:
                continue
            try:
                t0 = time.time()
                extras[name] = fn()
                extras[name]["wall_s"] = round(time.time() - t0, 1)
            except Exception as e:                  # noqa: BLE001
                extras[name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
            barrier()

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
        scan_kernel = "dense_scan_kernel" if lib.b200rag_get_option(b"scan_version") == 1 else "dense_scan3_kernel"
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"{args.rows}x{D} fp16 exact cosine top-{K}, query batch {B}, row-sharded over {world} GPU(s)",
                       "api": "B200IndexManager.search_batch_ids (value) / .search_batch_arrays (e2e)" if world == 1 else
                              "ShardedIndexManager.search_batch_ids (value) / .search_own_queries_arrays (e2e: every rank brings "
                              "B/N of the batch from its host and reads back those queries' results)",
                       "l2": f"inputs larger than L2 ({n_local * D * 2 / 1e9:.1f} GB per GPU streamed every step)",
                       "mode": "AUTO (tcgen05 scan + exact fp64 re-score, exact fallback for unproven queries)",
                       "flagged_queries_in_timed_region": int(flags_total.item())},
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * K * 16 + B * 4,
                    "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall_ms / args.steps, "checksum": checksum},
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
        if not args.no_cpu_baseline:
            # ~10-15 s of host work at N = 1; a shorter sample under torchrun (the other ranks wait at the barrier below)
            cb, _, _ = cpu_baseline_block(args, 5 if world == 1 else 2, 1)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
