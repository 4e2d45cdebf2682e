#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its configs[1] workload.

Workload "C2": FMT* precompute in the 2-D unit square, N = 1M synthetic uniform samples per
GPU, obstacle set ISRR_2H: one step = uniform-grid build + r-ball neighbour table (K1+K2) +
point validity (K6) + validity of every stored edge (K7: column classify + per-edge kernel).  metric = collision-checked edges/s
(the NN-queries/s figure of the same step is reported beside it).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1: launched by torch.distributed.run, one rank per GPU; weak scaling -- the sample set grows
to N x 1M (replicated on every GPU), rank g owns the query columns [g*1M, (g+1)*1M); the per-rank
CSC shards concatenate, and every step ends with the one exchange the host planner needs: an
NCCL all-gather of the shard colptrs and edge-validity words (no other data-path collective).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SAMPLES_PER_GPU = 1_000_000
SEED = 20240602
METRIC = "collision_checked_edges_per_sec"
UNIT = "edges/s"


def fill_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of rball_fill<2, U> on this workload, from the
    committed `ncu --set full` capture of the newest round (profiles/rN/traffic.json, scripts/ncu_traffic.py);
    None when no capture is committed.  Measured under the profiler, so it is traffic only, never a time."""
    import glob
    for path in sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*", "traffic.json")),
                       reverse=True):
        try:
            with open(path) as f:
                t = json.load(f)
        except (OSError, ValueError):
            continue
        for k, v in t.items():
            if k.startswith("rball_fill<2"):
                return v["dram_bytes_per_launch"]
    return None


def morton_order(V, bits=16):
    """argsort of 2-D points along a Z-order curve (host, numpy) -- for the auxiliary 'renumbered samples' figure"""
    q = np.minimum((V * (1 << bits)).astype(np.uint64), (1 << bits) - 1)

    def spread(x):
        x = (x | (x << 16)) & 0x0000FFFF0000FFFF
        x = (x | (x << 8)) & 0x00FF00FF00FF00FF
        x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0F
        x = (x | (x << 2)) & 0x3333333333333333
        x = (x | (x << 1)) & 0x5555555555555555
        return x
    return np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << 1), kind="stable")


def fmt_radius(N, d, rm=1.0, vol=1.0):
    import math
    return rm * 2 * (1 / d * vol / (math.pi ** (d / 2) / math.gamma(d / 2 + 1)) * math.log(N) / N) ** (1 / d)


def make_samples(n_total):
    return np.random.Generator(np.random.PCG64(SEED)).random((n_total, 2))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (pynvml, 2 ms period: the region is tens of ms)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for k, bit in names.items():
                if mask & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self._sample()
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
def oracle_step(orc, V, r, q0, q1, tree, O, So):
    """The reference's CPU path for columns [q0, q1): per-query kd-tree inball (sorted sparse
    column, nearneighbors.jl:179-183) + is_free_state for those samples + is_free_motion for
    every stored edge (robots2D.jl:13-14)."""
    colptr, rowval, nzval = tree.rball(r, q0, q1)
    orc.states_free(O, So, V[q0:q1])
    orc.edges_free_csc(O, So, V, colptr, rowval, q0)
    return len(rowval)


def run_oracle(V, r, q0, q1, threads, tree=None):
    """Time the oracle on columns [q0,q1) with `threads` host threads (ctypes drops the GIL).
    The kd-tree build (geometric.jl:14) is part of the timed path, done once, single-threaded
    as in the reference; pass `tree=(tree, t_build)` to reuse one."""
    from oracle import oracle as orc
    import fixtures as fx
    O = orc.Obstacles2D(fx.ISRR_2H)
    So = orc.StateSpace([0, 0], [1, 1])
    if tree is None:
        t0 = time.perf_counter()
        tree = orc.KDTree(V)
        t_build = time.perf_counter() - t0
    else:
        tree, t_build = tree
    bounds = np.linspace(q0, q1, threads + 1).astype(np.int64)
    edges = [0] * threads

    def work(i):
        edges[i] = oracle_step(orc, V, r, int(bounds[i]), int(bounds[i + 1]), tree, O, So)

    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    t_q = time.perf_counter() - t0
    return sum(edges), t_build, t_q


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def reference_arm(args, rank, world):
    """--impl reference: the oracle port of the reference's CPU path, all host threads, rank 0 only."""
    if rank != 0:
        return
    n_total = SAMPLES_PER_GPU * world
    V = make_samples(n_total)
    r = fmt_radius(n_total, 2)
    threads = host_threads()
    # bounded sample: a contiguous block of query columns, sized from one probe so that
    # (warmup + steps) stays within ~2 minutes
    probe = 20_000
    from oracle import oracle as orc
    t0 = time.perf_counter()
    tree = (orc.KDTree(V), 0.0)
    tree = (tree[0], time.perf_counter() - t0)
    _, tb, tq = run_oracle(V, r, 0, probe, threads, tree)
    per_col = tq / probe
    budget = 100.0 / max(1, args.steps + args.warmup)
    ncols = int(min(SAMPLES_PER_GPU * world, max(probe, budget / per_col)))
    times, edges = [], 0
    for it in range(args.warmup + args.steps):
        e, tb, tq = run_oracle(V, r, 0, ncols, threads, tree)
        # the tree build is amortised over the full sample set in the real workload: charge the
        # sample its proportional share
        t = tq + tb * (ncols / n_total)
        if it >= args.warmup:
            times.append(t)
            edges = e
    ms = 1e3 * float(np.mean(times))
    value = edges / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: FMT* 2-D unit square, ISRR_2H, N=%d uniform samples, r=%.7f" % (n_total, r),
                   "sample_columns": ncols},
        "nn_queries_per_sec": ncols / (ms / 1e3),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of %d query columns (kd-tree inball + point + edge checks), tree build charged pro rata"
                                   % (ncols, n_total)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import mpb200
    from mpb200 import _lib

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = mpb200.init(local)
    # ONE explicit stream for everything in the timed region: the L2 flush, the timing events, the library's
    # kernels and (N > 1) the NCCL exchange.  (torch's default stream has handle 0, which mpb200_set_stream
    # reads as "use the library's own non-blocking stream": the flush would then overlap the step.)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    _lib.check(lib.mpb200_set_stream(_lib.c_vp(stream.cuda_stream)))

    n_total = SAMPLES_PER_GPU * world
    r = fmt_radius(n_total, 2)
    samples = make_samples(n_total)
    if world > 1:
        # stripe order: each rank's query range is then a spatial stripe and its grid covers only the
        # stripe + r (FMT* is invariant to the order of i.i.d. samples; N = 1 keeps the raw order)
        samples = samples[np.argsort(samples[:, 0], kind="stable")]
    V_host = torch.from_numpy(np.ascontiguousarray(samples)).pin_memory()
    V = V_host.numpy()
    q0, q1 = rank * SAMPLES_PER_GPU, (rank + 1) * SAMPLES_PER_GPU
    CC = mpb200.PointRobot2D(mpb200.obstaclesets.ISRR_2H())
    SS = mpb200.UnitHypercube(2)
    CC.handle()

    NN = mpb200.MetricNN(V)
    NN.set_query_range(q0, q1)
    NN.handle()  # inputs resident in HBM before the timed region
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    exchange = None
    if world > 1:
        from mpb200 import sharding
        nnz0 = NN.build_table(r)
        cap = torch.tensor([nnz0], dtype=torch.int64, device="cuda")
        dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        exchange = sharding.ValidityExchange(SAMPLES_PER_GPU, (int(cap) * 21 // 20 + 63) // 64)

    phases = np.zeros(6)
    edge_ms, point_ms = [], []
    launches0 = launches1 = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    nnz = 0
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            sampler.start()
            launches0 = lib.mpb200_launch_count()
            t_wall0 = time.perf_counter()
        flush.zero_()  # L2 flush between iterations (outside the event pair)
        if it >= args.warmup:
            ev[it - args.warmup][0].record(stream)
        # the three calls of a planning step, issued back to back: only build_table waits (for nnz, once,
        # between count and fill); the validity calls enqueue their kernels and return
        NN.points_free(CC, SS, fetch=False)           # K6
        nnz = NN.build_table(r)                       # K1 + K2
        NN.edges_free(NN.table, CC, SS, fetch=False, count=False)  # K7
        if exchange is not None:
            exchange.run(NN.table)          # NCCL all-gather of shard colptrs + validity words
        if it >= args.warmup:
            ev[it - args.warmup][1].record(stream)
            # per-kernel times of THIS step (events recorded inside the calls, read after the step's end event)
            phases[:5] += [lib.mpb200_last_ms_of(_lib.OP_TABLE, k) for k in range(5)]
            point_ms.append(lib.mpb200_last_ms_of(_lib.OP_POINTS, 1))
            edge_ms.append(lib.mpb200_last_ms_of(_lib.OP_EDGES, 1))
            if sampler.nv and (it - args.warmup) % 8 == 0:
                sampler._sample()   # also from this thread, between steps (outside the event pairs): the region is short
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches1 = lib.mpb200_launch_count()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(np.sum(step_ms))
    tt = torch.tensor([total_ms, float(nnz)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, edges_all = float(tmax[0]), float(tsum[1])
    else:
        edges_all = float(nnz)
    ms_per_step = total_ms / args.steps
    value = edges_all / (ms_per_step / 1e3)
    queries_all = SAMPLES_PER_GPU * world

    # ---- auxiliary figure: the SAME samples numbered along a Z-order curve (what mpb200_sample_free's Morton
    # option produces).  FMT* does not care how i.i.d. samples are numbered; the kernels do (column writes and
    # gathers become local).  Reported beside the headline, never instead of it.
    renumbered = None
    if world == 1:
        NNm = mpb200.MetricNN(np.ascontiguousarray(V[morton_order(V)]))
        NNm.handle()
        evm = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for it in range(3 + len(evm)):
            flush.zero_()
            if it >= 3:
                evm[it - 3][0].record(stream)
            NNm.points_free(CC, SS, fetch=False)
            nnz_m = NNm.build_table(r)
            NNm.edges_free(NNm.table, CC, SS, fetch=False, count=False)
            if it >= 3:
                evm[it - 3][1].record(stream)
        torch.cuda.synchronize()
        ms_m = float(np.mean([a.elapsed_time(b) for a, b in evm]))
        renumbered = {"order": "Z-order (Morton) numbering of the same samples", "ms_per_step": ms_m,
                      "value": nnz_m / (ms_m / 1e3), "unit": UNIT, "steps": len(evm)}
        NNm.close()

    # ---- end to end through the public API with HOST buffers (H2D + D2H inside the timed region)
    e2e_times = []
    h2d = V.nbytes
    d2h = 0
    for it in range(2 + args.e2e_steps):
        NN2 = mpb200.MetricNN(V)          # fresh sample set: H2D of the inputs is inside
        NN2.pool = NN.pool                # result buffers (pinned) are reused across steps
        NN2.set_query_range(q0, q1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        cache, Eb, _ = NN2.precompute_checked(r, CC, SS)   # H2D samples, fused build, D2H colptr/rowval/nzval/edge bits
        Fb = NN2.points_free(CC, SS)      # D2H point bits
        t1 = time.perf_counter()
        if it >= 2:
            e2e_times.append(t1 - t0)
        d2h = cache.D.colptr.nbytes + cache.D.rowval.nbytes + cache.D.nzval.nbytes + Fb.nbytes + Eb.nbytes
        NN2.pool = _lib.PinnedPool()
        NN2.close()
    e2e_t = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = edges_all / float(e2e_t[0])

    if rank == 0:
        peak, peak_kind = peaks()
        deg = nnz / SAMPLES_PER_GPU
        fill_ms = phases[4] / args.steps
        alg_bytes = (8 * 2 + 8 + 16 * deg) * SAMPLES_PER_GPU   # 8d + 8 + 16*deg per query
        achieved = alg_bytes / (fill_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: FMT* 2-D unit square, ISRR_2H, N=%d uniform samples (%d query columns per GPU), r=%.7f"
                                   % (n_total, SAMPLES_PER_GPU, r),
                       "l2": "512 MiB flush between timed iterations", "index_type": "int64 (reference ABI)",
                       "parallelism": "query-range shards x%d (samples in stripe order when x > 1), samples+obstacles replicated%s"
                                      % (world, ", NCCL all-gather of colptr + validity words per step" if world > 1 else "")},
            "nn_queries_per_sec": queries_all / (ms_per_step / 1e3),
            "edges_per_step": edges_all, "mean_degree": deg,
            "phase_ms": {"grid_build": phases[1] / args.steps, "count_scan": phases[2] / args.steps,
                         "host_gap": phases[3] / args.steps, "fill": fill_ms,
                         "points_kernel": float(np.mean(point_ms)), "edges_kernels": float(np.mean(edge_ms)),
                         "inball_total": phases[0] / args.steps},
            "gpu_launches": int(launches1 - launches0),
            "renumbered_samples": renumbered,
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * float(e2e_t[0]), "steps": args.e2e_steps},
            "roofline": {"kernel": "rball_fill<2>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": fill_traffic(), "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": alg_bytes},
        }
        if world == 1 and not args.no_cpu_baseline:
            # the oracle port on this box's host cores: full workload, single thread (the reference is
            # single-threaded) -- reported baseline only
            e1, tb1, tq1 = run_oracle(V, r, 0, SAMPLES_PER_GPU, 1)
            line["cpu_baseline"] = {"value": e1 / (tb1 + tq1), "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "full workload: kd-tree build %.2fs + %d query columns, %d edges in %.2fs"
                                              % (tb1, SAMPLES_PER_GPU, e1, tq1),
                                    "host_cores_available": host_threads()}
        print(json.dumps(line), flush=True)
    NN.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
