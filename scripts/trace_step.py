"""Per-launch gaps of one C2 step (debug aid): MPB200_TRACE=1 MPB200_NO_GRAPH=1 python scripts/trace_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MPB200_TRACE", "1")
os.environ.setdefault("MPB200_NO_GRAPH", "1")
import numpy as np, torch
import mpb200
from mpb200 import _lib
from bench import fmt_radius, make_samples
lib = _lib.load()
G = int(os.environ.get("SHARDS", "1"))   # emulate rank 0 of a G-GPU weak-scaling run on one GPU
N = 1_000_000 * G
r = fmt_radius(N, 2)
V = make_samples(N)
if G > 1:
    V = V[np.argsort(V[:, 0], kind="stable")]
CC = mpb200.PointRobot2D(mpb200.obstaclesets.ISRR_2H()); SS = mpb200.UnitHypercube(2); CC.handle()
NN = mpb200.MetricNN(V)
if G > 1:
    NN.set_query_range(0, 1_000_000)
NN.handle()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
_lib.check(lib.mpb200_set_stream(_lib.c_vp(stream.cuda_stream)))
for it in range(4):
    flush.zero_()
    torch.cuda.synchronize()
    lib.mpb200_synchronize()   # dumps + resets the trace
    if it == 3: sys.stderr.write("[trace] ---- step\n")
    NN.points_free(CC, SS, fetch=False); NN.build_table(r); NN.edges_free(NN.table, CC, SS, fetch=False, count=False)
lib.mpb200_synchronize()
