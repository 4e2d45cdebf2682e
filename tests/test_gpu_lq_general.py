"""GPU parity of the GENERAL linear-affine ControlNN path (csrc/lq_general.cu) against its oracle
(oracle/lq_general.c): steer, both neighbour tables, and the swept edge validity along the optimal trajectory --
bit-exact -- for the two systems pinned to the reference's SymPy construction (tests/golden/lq_general.json):
a drifting double integrator with anisotropic R, and a triple integrator."""
import json
import os

import numpy as np
import pytest

import fixtures as fx
from conftest import unpack_bits

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lq_general.json")))["systems"]


def _pair(mp, orc, name):
    g = GOLD[name]
    A, B, c, R = (np.array(g[k]) for k in ("A", "B", "c", "R"))
    n = A.shape[0]
    if name == "di2_drift":
        lo, hi = np.array([0, 0, -1.5, -1.5]), np.array([1, 1, 1.5, 1.5])
        C = np.hstack([np.eye(2), np.zeros((2, 2))])
    else:
        lo, hi = np.array([0, -1.0, -2.0]), np.array([1, 1.0, 2.0])
        C = np.array([[1.0, 0, 0], [0, 1.0, 0]])        # workspace = (position, velocity) plane for the 2-D checker
    SS = mp.linearquadratic.LinearQuadraticQuasiMetricSpace(lo, hi, A, B, c, R, C)
    L = orc.LinearQuadraticGeneral(A, B, c, R)
    So = orc.StateSpace(lo, hi, ("matrix", C))
    return SS, L, So, lo, hi, g["cases"], n


@pytest.mark.parametrize("name", sorted(GOLD))
def test_steer_matches_oracle_and_golden(gpu, orc, name):
    mp = gpu
    SS, L, So, lo, hi, cases, n = _pair(mp, orc, name)
    X0 = np.array([c["x0"] for c in cases])
    X1 = np.array([c["x1"] for c in cases])
    for r in (0.5, 1.2):
        cost, topt = mp.linearquadratic.steer_batch(SS.dist, X0, X1, r)
        for k in range(len(cases)):
            ec, et = L.steer(X0[k], X1[k], r)
            assert cost[k] == ec and topt[k] == et                     # bit-exact vs the oracle
    # and the oracle's own pin: the reference's closures at the golden radius
    for k, c in enumerate(cases):
        cost, topt = mp.linearquadratic.steer_batch(SS.dist, X0[k], X1[k], c["r"])
        assert abs(topt[0] - c["topt"]) <= 1e-7 * max(1.0, c["r"])
        assert abs(cost[0] - c["cost_at_topt"]) <= 1e-8 * max(1.0, abs(c["cost_at_topt"]))
    assert mp.linearquadratic.steer(SS.dist, X0[0], X0[0], 1.0) == (0.0, 0.0)


@pytest.mark.parametrize("name,N,r", [("di2_drift", 3000, 0.9), ("triple", 2500, 1.3)])
def test_tables_and_edges_match_oracle(gpu, orc, name, N, r):
    mp = gpu
    SS, L, So, lo, hi, _, n = _pair(mp, orc, name)
    rng = np.random.Generator(np.random.PCG64(11 + n))
    V = lo + rng.random((N, n)) * (hi - lo)
    V[40] = V[7]                                                       # duplicate state
    NN = mp.QuasiMetricNN(V, SS.dist)
    NN.set_query_range(100, N - 50)                                    # a shard, too
    cF, cB = NN.precompute(r)
    for cache, forwards in ((cF, True), (cB, False)):
        ref = L.inball(V, r, forwards, 100, N - 50)
        assert np.array_equal(cache.D.colptr, ref[0]) and np.array_equal(cache.D.rowval, ref[1])
        assert cache.D.nzval.tobytes() == ref[2].tobytes()
        assert cache.D.nnz > N
    CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H())
    O = orc.Obstacles2D(fx.ISRR_2H)
    bits, checks = NN.lq_edges_free(CC, SS)
    refB = L.inball(V, r, False, 100, N - 50)
    exp, cnt = L.edges_free_csc(O, So, r, V, refB[0], refB[1], 100)
    assert np.array_equal(unpack_bits(bits, cB.D.nnz), np.asarray(exp).astype(bool)) and checks == cnt
    assert 0.02 < np.asarray(exp).mean() < 0.98
    # state-level form (fmt.jl:75 passes states) and is_free_path over consecutive pairs
    got = mp.linearquadratic.lq_motions_free(V[:400], V[400:800], CC, SS, r)
    for k in range(0, 400, 7):
        assert got[k] == L.is_free_motion(O, So, r, V[k], V[400 + k])[0]
    NN.close()


def test_create_accepts_nilpotent_and_rejects_the_rest(gpu):
    mp = gpu
    LQ = mp.linearquadratic.LinearQuadratic
    LQ(np.array([[0.0, 1.0], [0.0, 0.0]]), np.array([[0.0], [1.0]]), np.array([0.0, -9.81]), np.eye(1)).handle()   # 1-D DI with gravity
    with pytest.raises(NotImplementedError):                     # linearquadratic.jl:96, raised by the host mirror
        LQ(np.array([[0.0, 1.0], [-1.0, 0.0]]), np.array([[0.0], [1.0]]), np.zeros(2), np.eye(1))
    with pytest.raises(mp.MPB200Error, match="positive definite"):
        LQ(np.array([[0.0, 1.0], [0.0, 0.0]]), np.array([[0.0], [1.0]]), np.array([0.0, 1.0]), -np.eye(1)).handle()
    with pytest.raises(mp.MPB200Error, match="controllable"):
        LQ(np.array([[0.0, 1.0], [0.0, 0.0]]), np.array([[1.0], [0.0]]), np.zeros(2), np.eye(1)).handle()
