"""Small end-to-end exercise of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpb200
mpb200.init(0)
rng = np.random.Generator(np.random.PCG64(1))
SS = mpb200.UnitHypercube(2)
CC = mpb200.PointRobot2D(mpb200.obstaclesets.ISRR_2H(), fixed_point_test=True)
# grid r-ball (d = 2, 3), incl. shard mode, big columns, cell-ordered validity passes
for d, N, r in ((2, 6000, 0.03), (3, 3000, 0.12)):
    V = rng.random((N, d))
    NN = mpb200.MetricNN(V)
    NN.precompute(r)
    if d == 2:
        NN.points_free(CC, SS); NN.edges_free(NN.table, CC, SS)
        NN.set_query_range(1000, 4000); NN.build_table(r); NN.points_free(CC, SS); NN.edges_free(NN.table, CC, SS)
    NN.close()
V = rng.random((500, 2)) * 0.05
NN = mpb200.MetricNN(V); NN.precompute(1.0); NN.edges_free(NN.table, CC, SS); NN.close()       # big-column path
# x-sorted shard
V = rng.random((8000, 2)); V = V[np.argsort(V[:, 0])]
NN = mpb200.MetricNN(V); NN.set_query_range(2000, 5000); NN.precompute(0.03); NN.close()
# d >= 4: tensor-core prefilter + CUDA-core fallback dimension
for d in (6, 16):
    V = rng.random((1500, d)); NN = mpb200.MetricNN(V); NN.precompute(0.9 if d == 6 else 1.4); NN.close()
# boxes
B = mpb200.PointRobotNDBoxes([mpb200.BoxBounds(np.array([[0.2, 0.4], [0.2, 0.4], [0.0, 1.0]]))])
V = rng.random((3000, 3)); NN = mpb200.MetricNN(V); NN.precompute(0.1)
NN.points_free(B, mpb200.UnitHypercube(3)); NN.edges_free(NN.table, B, mpb200.UnitHypercube(3)); NN.close()
# LQ
DI = mpb200.DoubleIntegrator(2)
V = np.hstack([rng.random((800, 2)), rng.random((800, 2)) * 2 - 1])
Q = mpb200.QuasiMetricNN(V, DI.dist); Q.precompute(0.8); Q.lq_edges_free(CC, DI); Q.close()
# sampling (both orders) and Monte Carlo
mpb200.MetricNN.sample_free(CC, SS, 5000, seed=1).close()
mpb200.MetricNN.sample_free(CC, SS, 5000, seed=1, order="morton").close()
Bx = mpb200.PointRobotNDBoxes([mpb200.BoxBounds(np.array([0.5, -10.0]), np.array([10.0, 10.0]))])
P = mpb200.MCProblem(np.eye(2)[None], (np.eye(2) * 0.1)[None], np.eye(2), np.array([[0.2, 0.0]] * 2), [0.3, 0.7], [[3.0, 0.0]])
mpb200.collision_probability(P, Bx, 20000, seed=3)
print("sanitize run complete")
