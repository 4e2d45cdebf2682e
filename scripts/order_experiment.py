"""How much of the fill / count time is the order of the samples?  Same 1M samples, three index orders:
as drawn (random), sorted by x (stripes), Morton order (spatially coherent).  Debug aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpb200
from mpb200 import _lib
from bench import fmt_radius, make_samples
lib = _lib.load()
N = 1_000_000
r = fmt_radius(N, 2)
V0 = make_samples(N)

def morton(V, bits=16):
    q = np.minimum((V * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    def spread(x):
        x = (x | (x << 16)) & 0x0000FFFF0000FFFF
        x = (x | (x << 8)) & 0x00FF00FF00FF00FF
        x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0F
        x = (x | (x << 2)) & 0x3333333333333333
        x = (x | (x << 1)) & 0x5555555555555555
        return x
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1)

CC = mpb200.PointRobot2D(mpb200.obstaclesets.ISRR_2H()); SS = mpb200.UnitHypercube(2); CC.handle()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
_lib.check(lib.mpb200_set_stream(_lib.c_vp(stream.cuda_stream)))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for name, V in (("as drawn", V0), ("sorted by x", V0[np.argsort(V0[:, 0], kind="stable")]),
                ("morton", V0[np.argsort(morton(V0), kind="stable")])):
    NN = mpb200.MetricNN(np.ascontiguousarray(V)); NN.handle()
    ph = np.zeros(5); pe = []; pp = []
    for it in range(12):
        flush.zero_()
        NN.points_free(CC, SS, fetch=False); nnz = NN.build_table(r); NN.edges_free(NN.table, CC, SS, fetch=False, count=False)
        if it >= 2:
            ph += [lib.mpb200_last_ms_of(_lib.OP_TABLE, k) for k in range(5)]
            pp.append(lib.mpb200_last_ms_of(_lib.OP_POINTS, 1)); pe.append(lib.mpb200_last_ms_of(_lib.OP_EDGES, 1))
    ph /= 10
    print("%-12s nnz %d  front %.3f  fill %.3f  points %.3f  edges %.3f  total %.3f ms" %
          (name, nnz, ph[1], ph[4], np.mean(pp), np.mean(pe), ph[0] + np.mean(pp) + np.mean(pe)), flush=True)
    NN.close()
