#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs C1, C3, C4, C5 (bench.py is the headline C2 line).
One JSON line per config: device-resident GPU throughput + the oracle port on a bounded sample of
the same workload (single host thread, extrapolated and labelled as such).

  python bench_configs.py [--configs C1,C3,C4,C5] [--scale 1.0]
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


SYNC = None  # set to lib.mpb200_synchronize once the library is loaded


def timed(f, reps=1):
    best = math.inf
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = f()
        if SYNC is not None:
            SYNC()   # some entry points only enqueue their last kernels
        best = min(best, time.perf_counter() - t0)
    return best, out


def peaks_json():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def pipe_peak(lib, kind):
    """mpb200_pipe_peak: measured operations / s of an arithmetic pipe (0 = DADD+DMUL no-FMA, 1 = DFMA, 2 = FFMA)"""
    import ctypes
    v = ctypes.c_double(0.0)
    rc = lib.mpb200_pipe_peak(int(kind), ctypes.byref(v))
    if rc != 0:
        raise RuntimeError("mpb200_pipe_peak failed")
    return v.value


def table_parity(mp, NN, table, ref, c0, c1):
    """columns [c0, c1) of a device table against oracle columns ref = (colptr, rowval, nzval): byte-equal.
    Returns the number of columns checked; raises on a mismatch (a fast table that differs is not a result)."""
    from mpb200 import sharding
    colptr, rowval, nzval, _ = sharding.table_device_tensors(table)
    cp = colptr[c0:c1 + 1].cpu().numpy()
    lo, hi = int(cp[0]) - 1, int(cp[-1]) - 1
    rv = rowval[lo:hi].cpu().numpy() if hi > lo else np.zeros(0, dtype=np.int64)
    nz = nzval[lo:hi].cpu().numpy() if hi > lo else np.zeros(0)
    ok = (np.array_equal(cp - cp[0], ref[0] - ref[0][0]) and np.array_equal(rv, ref[1])
          and nz.tobytes() == np.ascontiguousarray(ref[2]).tobytes())
    if not ok:
        raise RuntimeError("PARITY FAILURE: device table differs from the oracle on columns [%d, %d)" % (c0, c1))
    return int(c1 - c0)


def c1(mp, orc, fx, args):
    """FMT* 2-D, ISRR_2H, N=1000, end to end (planner included)"""
    N = 1000
    SS = mp.UnitHypercube(2)
    CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H())
    O = orc.Obstacles2D(fx.ISRR_2H)
    So = orc.StateSpace([0, 0], [1, 1])
    cand = fx.uniform_samples(3 * N, 2, 20240601)
    V = np.vstack([[0.1, 0.1], cand[orc.states_free(O, So, cand)][:N - 2], [0.9, 0.9]])
    def run():
        P = mp.MPProblem(SS, [0.1, 0.1], mp.PointGoal([0.9, 0.9]), CC, V=mp.MetricNN(V, SS.dist, V[0]))
        out = mp.fmtstar(P, rm=1.0)
        P.V.close()
        return out
    run()
    t, (status, cost, _) = timed(run, 3)
    return dict(config="C1", N=N, status=status, cost=cost, gpu_s=t, note="host FMT* + GPU tables, N=1000 (latency-bound)")


def c3(mp, orc, fx, args):
    """10-D unit hypercube, random hyperboxes, N=2M: all-pairs r-ball + box edge checks"""
    d = 10
    N = int(2_000_000 * args.scale)
    V = fx.uniform_samples(N, d, 20240603)
    r = fx.fmt_radius(N, d)
    boxes = fx.random_hyperboxes(64, d, 20240613)
    CC = mp.PointRobotNDBoxes([mp.BoxBounds(*b) for b in boxes])
    SS = mp.UnitHypercube(d)
    NN = mp.MetricNN(V)
    NN.handle()
    lib = mp.load()
    global SYNC
    SYNC = lib.mpb200_synchronize
    from mpb200 import _lib
    t_nn, nnz = timed(lambda: NN.build_table(r), reps=2)   # best of 2: the first call also allocates
    ph = [lib.mpb200_last_ms_of(_lib.OP_TABLE, k) for k in range(5)]
    t_e, (_, checks) = timed(lambda: NN.edges_free(NN.table, CC, SS, fetch=False), reps=2)
    # oracle port on a bounded sample of query columns (brute-force truth; the kd-tree degenerates in 10-D)
    q = 256
    t_cpu, ref = timed(lambda: orc.rball_brute(V, r, 0, 0, q))
    B = orc.Boxes(boxes)
    So = orc.StateSpace(np.zeros(d), np.ones(d))
    t_cpu_e, _ = timed(lambda: orc.edges_free_csc(B, So, V, ref[0], ref[1], 0))
    # parity at full size (bench "parity_checked"): the oracle's brute-force columns against the device table
    parity = table_parity(mp, NN, NN.table, ref, 0, q)
    bf16 = peaks_json().get("bf16_tflops_sustained")
    flops = 2.0 * float(N) * float(N) * d        # the dense contraction 2 N^2 d
    out = dict(config="C3", workload="10-D unit hypercube, 64 random hyperboxes (seed 20240613), N=%d, r=%.5f: all-pairs r-ball table + box edge checks" % (N, r),
               N=N, d=d, r=r, nnz=int(nnz), mean_degree=nnz / N,
               nn_gpu_s=t_nn, nn_queries_per_s=N / t_nn, pair_tests_per_s=float(N) * N / t_nn, phase_ms=ph,
               edges_gpu_s=t_e, edges_per_s=nnz / t_e, checks=int(checks), parity_checked=parity,
               metric="nn_queries_per_sec", value=N / t_nn, unit="queries/s",
               roofline=dict(kernel="tc_rball_kernel<10>", bound="tensor", achieved=flops / t_nn / 1e12, peak=bf16,
                             unit="TFLOP/s", frac=(flops / t_nn / 1e12 / bf16) if bf16 else None, traffic=None,
                             peak_kind="measured bf16 sustained (MEASURED_PEAKS.json); the kernel issues TF32 MMAs, whose dense rate is half of bf16",
                             algorithmic_flops_per_launch=flops,
                             note="the contraction is K = 16 per pair, so the tensor pipe is never the limiter: the kernel is bound by its "
                                  "epilogue (accumulator read-back + running maxima) and, until candidates were queued per warp, by the rare "
                                  "exact-recheck path (ncu: DESIGN.md 10); the sweep multiplies each unordered pair once (symmetric form), so "
                                  "half of the algorithmic 2 N^2 d is actually issued"),
               cpu_baseline=dict(value=q / t_cpu, unit="queries/s", cores=1, kind="port",
                                 sample="brute-force oracle on %d of %d query columns; edges %.3g/s on their %d stored edges"
                                        % (q, N, len(ref[1]) / max(t_cpu_e, 1e-9), len(ref[1]))))
    NN.close()
    return out


def c4(mp, orc, fx, args):
    """double integrator, N=200k: ControlNN tables (both directions) + swept LQ edge checks"""
    N = int(200_000 * args.scale)
    rng = np.random.Generator(np.random.PCG64(20240604))
    SS = mp.DoubleIntegrator(2)
    V = SS.lo + rng.random((N, 4)) * (SS.hi - SS.lo)
    # calibrate r to a mean out-degree of ~64 on a query subsample
    L = orc.DoubleIntegratorLQ(2)
    sub = V[:20000]
    lo_r, hi_r = 0.3, 1.5
    for _ in range(12):
        r = 0.5 * (lo_r + hi_r)
        cp, _, _ = L.inball(sub, r, True, 0, 48)
        deg = (cp[-1] - 1) / 48 * (N / len(sub))
        lo_r, hi_r = (r, hi_r) if deg < 64 else (lo_r, r)
    CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H())
    NN = mp.QuasiMetricNN(V, SS.dist)
    NN.handle()
    lib = mp.load()
    global SYNC
    SYNC = lib.mpb200_synchronize
    from mpb200 import _lib
    t_nn, (nF, nB) = timed(lambda: NN.build_tables(r), reps=2)   # best of 2: the first call also allocates
    ph = [lib.mpb200_last_ms_of(_lib.OP_TABLE, k) for k in range(4)]
    t_e, (_, checks) = timed(lambda: NN.lq_edges_free(CC, SS, fetch=False), reps=2)
    q = 16
    t_cpu, ref = timed(lambda: L.inball(V, r, False, 0, q))
    C = np.hstack([np.eye(2), np.zeros((2, 2))])
    So = orc.StateSpace(SS.lo, SS.hi, ("matrix", C))
    t_cpu_e, _ = timed(lambda: L.edges_free_csc(orc.Obstacles2D(fx.ISRR_2H), So, r, V, ref[0], ref[1], 0))
    parity = table_parity(mp, NN, NN.tableB, ref, 0, q)
    # K5's all-pairs stage is an FP32 FMA prefilter since round 2 (both directions of a (query, sample) visit share
    # a = dp'R dp, b = (vx+vy)'R dp, g; 17 FMA + 6 add/mul = 40 flop per visit); the exact FP64 work
    # (candidate test, Newton, cost) only runs on the ~0.07% of pairs that survive both necessary conditions
    peak = pipe_peak(lib, 2)                       # measured FFMA rate, flop / s (an FMA counts as 2)
    ops = 40.0 * float(N) * float(N)
    out = dict(config="C4", workload="double integrator (4-D state), ISRR_2H, N=%d, r=%.5f (mean out-degree ~64): ControlNN tables both directions + swept LQ edge checks" % (N, r),
               N=N, r=r, nnzF=int(nF), nnzB=int(nB), mean_degree=nB / N,
               nn_gpu_s=t_nn, nn_queries_per_s=2 * N / t_nn, ordered_pairs_per_s=float(N) * N / t_nn, phase_ms=ph,
               edges_gpu_s=t_e, edges_per_s=nB / t_e, segment_checks=int(checks), parity_checked=parity,
               metric="nn_queries_per_sec", value=2 * N / t_nn, unit="queries/s",
               roofline=dict(kernel="lq_inball_kernel", bound="fp32", achieved=ops / t_nn / 1e9, peak=peak / 1e9, unit="GFLOP/s",
                             frac=ops / t_nn / peak, traffic=None,
                             peak_kind="measured live: mpb200_pipe_peak(FFMA); the all-pairs stage is FP32, the exact FP64 stage sees 0.07% of the pairs",
                             algorithmic_flops_per_launch=ops),
               cpu_baseline=dict(value=q / t_cpu, unit="queries/s", cores=1, kind="port",
                                 sample="oracle steer_pairwise on %d of %d backward columns (one direction); edges %.3g/s"
                                        % (q, N, len(ref[1]) / max(t_cpu_e, 1e-9))))
    NN.close()
    return out


def c5_problem(mp):
    """C5's problem (LQG-tracked double integrator past ISRR_2H, T = 100, defensive mixture proposal) -> (P, CC, naive)"""
    T, dt = 100, 0.05
    A = np.block([[np.eye(2), dt * np.eye(2)], [np.zeros((2, 2)), np.eye(2)]])
    B = np.vstack([0.5 * dt * dt * np.eye(2), dt * np.eye(2)])
    C = np.hstack([np.eye(2), np.zeros((2, 2))])
    F, G = mp.montecarlo.lqg_closed_loop(A, B, C, np.eye(4), 0.1 * np.eye(2), 1e-4 * np.eye(4), 1e-4 * np.eye(2), T)
    Wz = np.hstack([np.eye(2), np.zeros((2, 6))])
    s = np.linspace(0, 1, T + 1)
    wbar = np.stack([0.30 + 0.35 * s, 0.135 + 0.0 * s], axis=1)          # 0.055 below box 2
    CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H(), fixed_point_test=True)
    naive = mp.MCProblem(F, G, Wz, wbar)
    P = mp.montecarlo.with_proposal(naive, CC, r2=36.0, max_components=8)
    return P, CC, naive


def mc_flops_per_rollout(P):
    """algorithmic FP64 operations of one rollout: per step the closed-loop update
    2 nz^2 + 2 nz q, the workspace map 2 dw nz, q/2 Box-Muller pairs (log 34, sincos 46, sqrt + 4) and the K
    mixture inner products 2 q K; per rollout K exp (36 each) for the weight."""
    nz, q, dw, K, T = P.F.shape[1], P.G.shape[2], P.Wz.shape[0], P.K, P.F.shape[0]
    per_step = 2 * nz * nz + 2 * nz * q + 2 * dw * nz + (q // 2) * (34 + 46 + 6) + 2 * q * K
    return T * per_step + 36 * K + 2 * K


def c5(mp, orc, fx, args):
    """Monte-Carlo collision probability, LQG-tracked double integrator past ISRR_2H, 1e8 rollouts"""
    P, CC, naive_problem = c5_problem(mp)
    T = P.F.shape[0]
    n = int(1e8 * args.scale)
    mp.collision_probability(P, CC, 100_000)
    t, res = timed(lambda: mp.collision_probability(P, CC, n))
    O = orc.Obstacles2D(fx.ISRR_2H, fixed_point_test=True)
    spec = orc.McSpec(P.F, P.G, P.Wz, P.wbar, P.alpha, P.mu if P.K else None)
    nq = 20000
    t_cpu, ref = timed(lambda: orc.mc_run(spec, O, 20240605, 0, nq, per_rollout=True))
    naive = mp.collision_probability(naive_problem, CC, min(n, 10_000_000), seed=7)
    lib = mp.load()
    peak = pipe_peak(lib, 0)
    ops = float(mc_flops_per_rollout(P)) * n
    # parity of the sample against the oracle: per-rollout hit bits and weights byte-equal
    dev = mp.collision_probability(P, CC, nq, per_rollout=True)
    same = bool(np.array_equal(dev["hit"], np.asarray(ref["hit"]).astype(bool)) and dev["w"].tobytes() == ref["w"].tobytes())
    if not same:
        raise RuntimeError("PARITY FAILURE: Monte-Carlo rollouts differ from the oracle")
    return dict(config="C5", workload="Monte-Carlo collision probability, LQG-tracked double integrator past ISRR_2H, T=%d, K=%d mixture components, %d rollouts" % (T, P.K, n),
                rollouts=n, T=T, K=P.K, p=res["p"], se=res["se"], hits=res["hits"],
                naive_p=naive["p"], naive_se=naive["se"], gpu_s=t, rollouts_per_s=n / t,
                metric="rollouts_per_sec", value=n / t, unit="rollouts/s", parity_checked=nq,
                roofline=dict(kernel="mc_rollout_kernel", bound="fp64", achieved=ops / t / 1e9, peak=peak / 1e9, unit="GFLOP/s",
                              frac=ops / t / peak, traffic=None,
                              peak_kind="measured live: mpb200_pipe_peak(DADD_DMUL)", algorithmic_flops_per_launch=ops),
                cpu_baseline=dict(value=nq / t_cpu, unit="rollouts/s", cores=1, kind="port",
                                  sample="%d rollouts of the same problem (oracle/mc.c)" % nq))


def f1(mp, orc, fx, args):
    """batched free-state sampling on the device (the bulk of sample_free!), C2's space and
    obstacle set with the intended point test (78% of the square is free)."""
    N = int(1_000_000 * args.scale)
    CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H(), fixed_point_test=True)
    SS = mp.UnitHypercube(2)
    CC.handle()
    mp.MetricNN.sample_free(CC, SS, 1000, seed=1).close()      # warm: module load, first allocations
    ts = []
    for _ in range(3):   # best of three: the first call also allocates
        t_one, NN = timed(lambda: mp.MetricNN.sample_free(CC, SS, N, seed=2))     # device sampling + D2H of the host copy
        ts.append(t_one)
        if _ < 2:
            NN.close()
    t_all = min(ts)
    used = NN.candidates
    NN.close()
    tm = []
    for _ in range(2):
        t_one, NM = timed(lambda: mp.MetricNN.sample_free(CC, SS, N, seed=2, order="morton"))
        tm.append(t_one)
        NM.close()
    nq = 200_000
    O = orc.Obstacles2D(fx.ISRR_2H, fixed_point_test=True)
    t_cpu, (V, _) = timed(lambda: orc.sample_free(O, orc.StateSpace([0, 0], [1, 1]), nq, 2))
    return dict(config="F1", N=N, candidates=used, acceptance=N / used, gpu_s=t_all, samples_per_s=N / t_all,
                candidates_per_s=used / t_all, morton_order_gpu_s=min(tm), note="includes the D2H of the 16 MB host copy",
                cpu_port=dict(sample=nq, samples_per_s=nq / t_cpu, cores=1))


def f4(mp, orc, fx, args):
    """chopped-metric car spaces.  N = 200k SE2 states in the unit square past ISRR_2H, turning radius
    0.01, r = 0.025 (~390 (x, y) candidates per column): Reeds-Shepp table (one direction, 48 path evaluations per
    candidate), Dubins tables (both directions, 6 words each) and the arc-waypoint edge checks of the Dubins table."""
    N = int(200_000 * args.scale)
    rturn, r = 0.01, 0.025
    rng = np.random.Generator(np.random.PCG64(20240605))
    V = np.column_stack([rng.random(N), rng.random(N), rng.uniform(0, 2 * np.pi, N)])
    CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H())
    lib = mp.load()
    global SYNC
    SYNC = lib.mpb200_synchronize
    from mpb200 import _lib
    out = dict(config="F4", workload="car spaces: N=%d SE2 states, ISRR_2H, turning radius %.3g, r=%.3g" % (N, rturn, r), N=N)
    q = 32
    So = orc.StateSpace([0, 0, 0], [1, 1, 2 * np.pi], ("view", [1, 2]))
    for kind in ("reedsshepp", "dubins"):
        SS = (mp.ReedsSheppMetricSpace if kind == "reedsshepp" else mp.DubinsQuasiMetricSpace)(rturn)
        mp.setup_steering(SS, r)
        car = orc.SimpleCar(kind, rturn)
        if kind == "reedsshepp":
            NN = mp.MetricNN(V, SS.dist, V[0])
            NN.handle()
            t_nn, nnz = timed(lambda: NN.build_table(r), reps=2)
            table, evals = NN.table, 1
        else:
            NN = mp.QuasiMetricNN(V, SS.dist, V[0])
            NN.handle()
            t_nn, (nF, nnz) = timed(lambda: NN.build_tables(r), reps=2)
            table, evals = NN.tableB, 2
        ph = [lib.mpb200_last_ms_of(_lib.OP_TABLE, k) for k in range(3)]
        cand = int(NN_candidates(lib, NN))
        t_e, (_, checks) = timed(lambda: NN.car_edges_free(CC, SS, fetch=False), reps=2)
        t_cpu, ref = timed(lambda: car.inball(V, r, kind == "reedsshepp", q0=0, q1=q))
        t_cpu_e, _ = timed(lambda: car.edges_free_csc(orc.Obstacles2D(fx.ISRR_2H), So, V, ref[0], ref[1], 0))
        parity = table_parity(mp, NN, table, ref, 0, q)
        out[kind] = dict(nnz=int(nnz), mean_degree=nnz / N, candidate_pairs=cand, table_gpu_s=t_nn, cost_kernel_ms=ph[1],
                         steer_evaluations_per_s=evals * cand / (ph[1] / 1e3) if ph[1] > 0 else None,
                         nn_queries_per_s=evals * N / t_nn, edges_gpu_s=t_e, edges_per_s=nnz / t_e, segment_checks=int(checks),
                         parity_checked=parity,
                         cpu_baseline=dict(value=q / t_cpu, unit="queries/s", cores=1, kind="port",
                                           sample="oracle/cars.c inball on %d of %d columns (one direction); edges %.3g/s"
                                                  % (q, N, len(ref[1]) / max(t_cpu_e, 1e-9))))
        NN.close()
    out.update(metric="nn_queries_per_sec", value=out["reedsshepp"]["nn_queries_per_s"], unit="queries/s")
    return out


def NN_candidates(lib, NN):
    """entries of the (x, y) candidate table behind the last car build = pairs the cost kernel evaluated"""
    import ctypes
    n = ctypes.c_int64(0)
    lib.mpb200_car_last_candidates(NN.handle(), ctypes.byref(n))
    return n.value


def kn(mp, orc, fx, args):
    """k-nearest connections on C2's sample set: k from fmt.jl:6, r-ball table grown until every column
    holds k entries, k-selection, mutual neighbourhoods (device-resident; the host fetch is not timed)."""
    import math
    from mpb200 import nearneighbors as nnm
    N = int(1_000_000 * args.scale)
    V = fx.uniform_samples(N, 2, 20240602)
    k = min(int(math.ceil(4 * (math.e / 2) * math.log(N))), N - 1)
    NN = mp.MetricNN(V)
    NN.handle()
    lib = mp.load()
    global SYNC
    SYNC = lib.mpb200_synchronize
    NN.table_knn, NN.table_mknn = nnm.DeviceTable("knn"), nnm.DeviceTable("mknn")
    state = {}

    def run():
        import time as _tm
        r = nnm._knn_radius_guess(V, k)
        rounds = 0
        tb = ts = 0.0
        while True:
            rounds += 1
            SYNC(); t0 = _tm.perf_counter()
            NN.build_table(r)
            SYNC(); t1 = _tm.perf_counter()
            short = nnm._short_columns(NN.table, k)
            tb += t1 - t0
            if short == 0:
                break
            r *= 1.3
        SYNC(); t1 = _tm.perf_counter()
        nnm._table_knn(NN.table, k, NN.table_knn)
        SYNC(); t2 = _tm.perf_counter()
        ts = t2 - t1
        nnm._table_union_transpose(NN.table_knn, NN.table_knn, NN.table_mknn)
        SYNC(); t3 = _tm.perf_counter()
        state.update(r=r, rounds=rounds, nnz_ball=NN.table.nnz, nnz_mutual=NN.table_mknn.nnz,
                     ball_tables_s=tb, k_selection_s=ts, mutual_s=t3 - t2)
    t_gpu, _ = timed(run, reps=2)
    q = 64
    # oracle: brute force k-selection for q columns over all N samples
    import time as _t
    t0 = _t.perf_counter()
    cols = []
    for c in range(q):
        d = np.sqrt((V[c, 0] - V[:, 0]) ** 2 + (V[c, 1] - V[:, 1]) ** 2)
        cols.append(orc.knn_from_values(d, k, c))
    t_cpu = _t.perf_counter() - t0
    D = NN.fetch_table(NN.table_knn, "knn")
    for c in range(q):
        got = D.rowval[D.colptr[c] - 1:D.colptr[c + 1] - 1] - 1
        if not np.array_equal(got, cols[c]):
            raise RuntimeError("PARITY FAILURE: k-nearest column %d differs from the brute-force selection" % c)
    out = dict(config="KNN", workload="k-nearest connections: C2's %d samples, k=%d (fmt.jl:6)" % (N, k), N=N, k=k,
               gpu_s=t_gpu, nn_queries_per_s=N / t_gpu, metric="nn_queries_per_sec", value=N / t_gpu, unit="queries/s",
               parity_checked=q, cpu_baseline=dict(value=q / t_cpu, unit="queries/s", cores=1, kind="port",
                                                   sample="numpy brute-force selection on %d of %d columns" % (q, N)), **state)
    NN.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C3,C4,C5,F1,F4,KNN")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    import mpb200
    from oracle import oracle as orc
    import fixtures as fx
    mpb200.init(int(os.environ.get("LOCAL_RANK", "0")))
    table = dict(C1=c1, C3=c3, C4=c4, C5=c5, F1=f1, F4=f4, KNN=kn)
    for name in args.configs.split(","):
        out = table[name](mpb200, orc, fx, args)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
