"""GPU parity for the linear-quadratic path (K5 ControlNN tables, K9 swept LQ edges, steer batches)
against the oracle (oracle/lq.c): identical operation order on both sides -> bit-identical."""
import numpy as np
import pytest

import fixtures as fx
from conftest import unpack_bits

pytestmark = pytest.mark.gpu


def di_samples(N, seed, vmax=1.5):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.hstack([rng.random((N, 2)), (rng.random((N, 2)) * 2 - 1) * vmax])


def test_steer_batch_bit_exact(gpu, orc):
    mp = gpu
    SS = mp.DoubleIntegrator(2)
    L = orc.DoubleIntegratorLQ(2)
    V = di_samples(20_000, 1)
    W = V + (np.random.Generator(np.random.PCG64(2)).random(V.shape) - 0.5) * [0.3, 0.3, 0.8, 0.8]
    W[:50] = V[:50]                                  # identical states -> (0, 0)
    for r in (0.5, 1.0):
        cost, topt = mp.steer_batch(SS.dist, V, W, r)
        exp = np.array([L.steer(v, w, r) for v, w in zip(V[:4000], W[:4000])])
        assert cost[:4000].tobytes() == exp[:, 0].tobytes()
        assert topt[:4000].tobytes() == exp[:, 1].tobytes()
        assert (cost[:50] == 0).all() and (topt[:50] == 0).all()
        assert (topt <= r).all() and (topt[50:] > 0).all()


def test_lq_create_takes_other_nilpotent_systems_and_rejects_the_rest(gpu, orc):
    mp = gpu
    A = np.zeros((4, 4)); A[0, 1] = 1; A[1, 2] = 1; A[2, 3] = 1      # nilpotent chain, not [0 I;0 0]: the numeric path
    B = np.zeros((4, 2)); B[3, 0] = 1; B[2, 1] = 1
    d = mp.LinearQuadratic(A, B, np.zeros(4), np.eye(2))
    d.handle()
    rng = np.random.Generator(np.random.PCG64(4))
    X0, X1 = rng.random((64, 4)) - 0.5, rng.random((64, 4)) - 0.5
    cost, topt = mp.linearquadratic.steer_batch(d, X0, X1, 0.8)
    L = orc.LinearQuadraticGeneral(A, B, np.zeros(4), np.eye(2))
    for k in range(64):
        assert (cost[k], topt[k]) == L.steer(X0[k], X1[k], 0.8)
    with pytest.raises(NotImplementedError):                          # linearquadratic.jl:96
        mp.LinearQuadratic(np.eye(2), np.ones((2, 1)), np.zeros(2), np.eye(1))


@pytest.mark.parametrize("rho", [1.0, 2.5])
def test_k5_tables_match_oracle(gpu, orc, rho):
    mp = gpu
    N, r = 3000, 0.8
    V = di_samples(N, 7)
    V[100] = V[50]                                   # duplicate states
    SS = mp.DoubleIntegrator(2, r=rho)
    L = orc.DoubleIntegratorLQ(2, rho * np.eye(2))
    NN = mp.QuasiMetricNN(V, SS.dist)
    cF, cB = NN.precompute(r)
    for cache, fwd in ((cF, True), (cB, False)):
        ref = L.inball(V, r, fwd)
        assert np.array_equal(cache.D.colptr, ref[0])
        assert np.array_equal(cache.D.rowval, ref[1])
        assert cache.D.nzval.tobytes() == ref[2].tobytes()
    assert cF.D.nnz == cB.D.nnz > 1000
    # column views
    v = 77
    col = mp.inballF(NN, v, r)
    assert np.array_equal(col.nzind, cF.D.rowval[cF.D.colptr[v - 1] - 1:cF.D.colptr[v] - 1])
    NN.close()


def test_k5_general_R_and_shards(gpu, orc):
    mp = gpu
    R = np.array([[2.0, 0.3], [0.3, 1.0]])
    N, r = 1500, 0.9
    V = di_samples(N, 11)
    SSd = mp.DoubleIntegrator(2)
    dist = mp.LinearQuadratic(SSd.dist.A, SSd.dist.B, SSd.dist.c, R)
    L = orc.DoubleIntegratorLQ(2, R)
    full = L.inball(V, r, False)
    got = []
    for q0, q1 in ((0, 700), (700, 1500)):
        NN = mp.QuasiMetricNN(V, dist)
        NN.set_query_range(q0, q1)
        _, cB = NN.precompute(r)
        ref = L.inball(V, r, False, q0, q1)
        assert np.array_equal(cB.D.colptr, ref[0]) and np.array_equal(cB.D.rowval, ref[1])
        assert cB.D.nzval.tobytes() == ref[2].tobytes()
        got.append(cB.D.rowval.copy())
        NN.close()
    assert np.array_equal(np.concatenate(got), full[1])


@pytest.mark.parametrize("checker", ["sat2d", "boxes"])
def test_k9_lq_edges_match_oracle(gpu, orc, checker):
    mp = gpu
    N, r = 2500, 0.8
    V = di_samples(N, 21)
    SS = mp.DoubleIntegrator(2)
    L = orc.DoubleIntegratorLQ(2)
    C = np.hstack([np.eye(2), np.zeros((2, 2))])
    So = orc.StateSpace([0, 0, -1.5, -1.5], [1, 1, 1.5, 1.5], ("matrix", C))
    if checker == "sat2d":
        CC, O = mp.PointRobot2D(mp.obstaclesets.ISRR_2H()), orc.Obstacles2D(fx.ISRR_2H)
    else:
        CC, O = mp.PointRobotNDBoxes([mp.BoxBounds(b) for b in fx.BOXES2D]), orc.Boxes(fx.BOXES2D)
    NN = mp.QuasiMetricNN(V, SS.dist)
    _, cB = NN.precompute(r)
    CC.count = 0
    bits, checks = NN.lq_edges_free(CC, SS)
    exp, cnt = L.edges_free_csc(O, So, r, V, cB.D.colptr, cB.D.rowval)
    got = unpack_bits(bits, cB.D.nnz)
    assert np.array_equal(got, exp.astype(bool))
    assert checks == cnt == CC.count
    assert 0.05 < got.mean() < 0.98
    # state-level batch (fmt.jl:75 called with states)
    cols = np.repeat(np.arange(N), np.diff(cB.D.colptr))
    sel = np.arange(0, cB.D.nnz, 7)
    ok = mp.lq_motions_free(V[cB.D.rowval[sel] - 1], V[cols[sel]], CC, SS, r)
    assert np.array_equal(ok, got[sel])
    mp.setup_steering(SS, r)                         # linearquadratic.jl:34
    assert mp.is_free_motion(V[cB.D.rowval[0] - 1], V[cols[0]], CC, SS) == bool(got[0])
    NN.close()


def test_k5_single_sweep_slab_overflow_falls_back(gpu, orc):
    mp = gpu
    rng = np.random.Generator(np.random.PCG64(8))
    spread = np.hstack([rng.random((1500, 2)), (rng.random((1500, 2)) * 2 - 1) * 1.5])
    cluster = np.hstack([0.5 + 0.01 * rng.random((400, 2)), 0.05 * (rng.random((400, 2)) - 0.5)])
    V = np.vstack([spread, cluster])
    r = 0.6
    SS = mp.DoubleIntegrator(2)
    L = orc.DoubleIntegratorLQ(2)
    NN = mp.QuasiMetricNN(V, SS.dist)
    cF, cB = NN.precompute(r)
    for cache, fwd in ((cF, True), (cB, False)):
        ref = L.inball(V, r, fwd)
        assert np.array_equal(cache.D.colptr, ref[0]) and np.array_equal(cache.D.rowval, ref[1])
        assert cache.D.nzval.tobytes() == ref[2].tobytes()
    deg = np.diff(cB.D.colptr)
    assert 2 * deg[:1024].max() + 64 < deg.max()
    NN.close()
