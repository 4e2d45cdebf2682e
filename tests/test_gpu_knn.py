"""GPU parity of the k-nearest connections (csrc/knn.cu) against the oracle's restatement of the same specification
(the reference exports knn* / mutualknn* and defines none of them: parity unpinned): k-NN and mutual tables byte-equal
for 2-D / 3-D (grid path) and 6-D (all-pairs path), lattice ties, k >= N - 1, the steering-cost tables, and FMT* with
connections = :K against the same planner over the oracle's tables."""
import numpy as np
import pytest

import fixtures as fx
from conftest import unpack_bits
from oracle_fmt import fmt_oracle

pytestmark = pytest.mark.gpu


def _same(D, ref):
    return (np.array_equal(D.colptr, ref[0]) and np.array_equal(D.rowval, ref[1])
            and np.asarray(D.nzval).tobytes() == np.ascontiguousarray(ref[2]).tobytes())


@pytest.mark.parametrize("d,N,k", [(2, 3000, 25), (3, 2000, 15), (6, 1200, 10), (2, 40, 39), (2, 50, 200)])
def test_knn_and_mutual_match_oracle(gpu, orc, d, N, k):
    mp = gpu
    V = fx.uniform_samples(N, d, 100 + d + N)
    NN = mp.MetricNN(V)
    cK, cM = NN.precompute_knn(k)
    ref = orc.knn_brute(V, k)
    assert _same(cK.D, ref)
    assert _same(cM.D, orc.union_transpose(ref, ref, N))
    kk = min(k, N - 1)
    v = N // 2 + 1
    col = mp.knn(NN, v, kk)
    assert len(col.nzind) == kk and np.array_equal(col.nzind, ref[1][ref[0][v - 1] - 1:ref[0][v] - 1])
    f = np.zeros(N, dtype=bool); f[::2] = True
    assert np.array_equal(mp.mutualknnF(NN, v, kk, f).nzind, mp.mutualknn(NN, v, kk).nzind[f[mp.mutualknn(NN, v, kk).nzind - 1]])
    NN.close()


def test_knn_lattice_ties(gpu, orc):
    mp = gpu
    g = np.stack(np.meshgrid(np.arange(30.0), np.arange(30.0)), -1).reshape(-1, 2) / 32.0
    NN = mp.MetricNN(g)
    cK, cM = NN.precompute_knn(6)
    ref = orc.knn_brute(g, 6)
    assert _same(cK.D, ref) and _same(cM.D, orc.union_transpose(ref, ref, len(g)))
    NN.close()


def test_knn_edges_validity_on_the_knn_table(gpu, orc):
    mp = gpu
    N, k = 4000, 20
    V = fx.uniform_samples(N, 2, 77)
    NN = mp.MetricNN(V)
    cK, _ = NN.precompute_knn(k)
    CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H())
    SS = mp.UnitHypercube(2)
    bits, checks = NN.edges_free(NN.table_knn, CC, SS)
    exp, cnt = orc.edges_free_csc(orc.Obstacles2D(fx.ISRR_2H), orc.StateSpace([0, 0], [1, 1]), V, cK.D.colptr, cK.D.rowval)
    assert np.array_equal(unpack_bits(bits, cK.D.nnz), exp.astype(bool)) and checks == cnt == N * k
    NN.close()


def test_lq_knn_tables(gpu, orc):
    mp = gpu
    N, k = 900, 8
    SS = mp.DoubleIntegrator(2)
    rng = np.random.Generator(np.random.PCG64(3))
    V = SS.lo + rng.random((N, 4)) * (SS.hi - SS.lo)
    NN = mp.QuasiMetricNN(V, SS.dist)
    cF, cB, cM = NN.precompute_knn(k, 0.6)
    L = orc.DoubleIntegratorLQ(2)
    r = cB.r
    TF, TB = L.inball(V, r, True), L.inball(V, r, False)
    assert np.diff(TF[0]).min() >= k and np.diff(TB[0]).min() >= k
    kF, kB = orc.knn_of_table(*TF, k), orc.knn_of_table(*TB, k)
    assert _same(cF.D, kF) and _same(cB.D, kB)
    assert _same(cM.D, orc.union_transpose(kF, kB, N))
    NN.close()


@pytest.mark.parametrize("edge_checks", ["table", "lazy"])
def test_fmt_k_nearest_connections(gpu, orc, edge_checks):
    """fmtstar!(P, N; connections = :K): same tree, path and cost as the reference's loop over the oracle's k-NN tables"""
    mp = gpu
    N = 1500
    O = orc.Obstacles2D(fx.ISRR_2H)
    So = orc.StateSpace([0, 0], [1, 1])
    cand = fx.uniform_samples(3 * N, 2, 20240601)
    V = np.vstack([[0.1, 0.1], cand[orc.states_free(O, So, cand)][:N - 2], [0.9, 0.9]])
    SS = mp.UnitHypercube(2)
    CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H())
    P = mp.MPProblem(SS, [0.1, 0.1], mp.PointGoal([0.9, 0.9]), CC, V=mp.MetricNN(V, SS.dist, V[0]))
    status, cost, _ = mp.fmtstar(P, connections="K", rm=1.0, edge_checks=edge_checks)
    k = P.solution.metadata["k"]
    assert k == min(int(np.ceil(4 * (np.e / 2) * np.log(N))), N - 1)
    K = orc.knn_brute(V, k)
    M = orc.union_transpose(K, K, N)
    col = lambda T, v: (T[1][T[0][v - 1] - 1:T[0][v] - 1], T[2][T[0][v - 1] - 1:T[0][v] - 1])

    def edge(y0, x0):
        ok, n = orc.motions_free_straight(O, So, V[y0:y0 + 1], V[x0:x0 + 1])
        return bool(ok[0]), n

    ref = fmt_oracle(V, np.all(V == [0.9, 0.9], axis=1), lambda v: col(M, v), lambda v: col(K, v),
                     lambda i: bool(orc.states_free(O, So, V[i:i + 1])[0]), edge)
    assert ref["solved"] and status == "solved"
    assert P.solution.metadata["path"] == ref["path"] and np.array_equal(P.solution.metadata["tree"], ref["tree"])
    assert abs(cost - ref["cost"]) <= 1e-12 * ref["cost"]
    assert P.solution.metadata["collision_checks"] == ref["checks"]
    P.V.close()
