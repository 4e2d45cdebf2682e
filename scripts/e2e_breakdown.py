"""Where the end-to-end (host buffers in, host buffers out) time of one C2 step goes (debug aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpb200
from mpb200 import _lib
from bench import fmt_radius, make_samples
lib = _lib.load()
N = 1_000_000
r = fmt_radius(N, 2)
V_host = torch.from_numpy(np.ascontiguousarray(make_samples(N, False))).pin_memory()
V = V_host.numpy()
CC = mpb200.PointRobot2D(mpb200.obstaclesets.ISRR_2H()); SS = mpb200.UnitHypercube(2); CC.handle()
pool = _lib.PinnedPool(reuse=True)
def T():
    lib.mpb200_synchronize(); return time.perf_counter()
# raw PCIe numbers with the library's pinned buffers
big = pool.array("probe", 220_000_000 // 8, np.int64)
dev = torch.empty(220_000_000 // 8, dtype=torch.int64, device="cuda")
hp = torch.from_numpy(big)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); hp.copy_(dev, non_blocking=True); torch.cuda.synchronize()
    print("D2H 220 MB into library-pinned buffer: %.2f ms (%.1f GB/s)" % ((time.perf_counter() - t0) * 1e3, 0.22 / (time.perf_counter() - t0)))
for it in range(4):
    t = [T()]
    NN = mpb200.MetricNN(V); NN.pool = pool; NN.handle(); t.append(T())
    nnz, checks = NN.build_table_checked(r, CC, SS); t.append(T())
    D = NN.fetch_table(NN.table); t.append(T())
    Eb = NN.fetch_edge_bits(); t.append(T())
    Fb = NN.points_free(CC, SS); t.append(T())
    NN.pool = _lib.PinnedPool(reuse=True); NN.close(); t.append(T())
    names = ["create+H2D", "build_checked", "fetch_table", "fetch_edge_bits", "points_free+D2H", "close"]
    print("  ".join("%s %.2f" % (n, (b - a) * 1e3) for n, a, b in zip(names, t[:-1], t[1:])), " total(ex close) %.2f ms" % ((t[-2] - t[0]) * 1e3))
