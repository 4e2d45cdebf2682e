"""GPU parity of the chopped-metric car spaces (csrc/cars.cu) against oracle/cars.c, bit for bit: steering cost and
control, the forward / backward neighbour tables (nearneighbors.jl:185-198 over simplecars.jl), the waypoint collision
checks (statespaces.jl:153-158 over simplecars.jl:55-82) with both obstacle kinds, sharded query ranges, and FMT* over a
Reeds-Shepp and a Dubins space against the reference's loop run over the oracle's tables."""
import math

import numpy as np
import pytest

import fixtures as fx
from conftest import unpack_bits
from oracle_fmt import fmt_oracle

pytestmark = pytest.mark.gpu
TWO_PI = 2 * math.pi
KINDS = ["reedsshepp", "dubins"]


def _states(rng, n, lo=0.0, hi=1.0):
    return np.column_stack([rng.uniform(lo, hi, n), rng.uniform(lo, hi, n), rng.uniform(0, TWO_PI, n)])


def _space(mp, kind, rturn, speed=1.0):
    return (mp.ReedsSheppMetricSpace if kind == "reedsshepp" else mp.DubinsQuasiMetricSpace)(rturn, speed)


def _same(D, ref):
    return (np.array_equal(D.colptr, ref[0]) and np.array_equal(D.rowval, ref[1])
            and np.asarray(D.nzval).tobytes() == np.ascontiguousarray(ref[2]).tobytes())


@pytest.mark.parametrize("kind", KINDS)
def test_steer_is_bit_identical(gpu, orc, kind):
    mp = gpu
    rng = np.random.Generator(np.random.PCG64(11))
    n = 6000
    V, W = _states(rng, n, 0, 3), _states(rng, n, 0, 3)
    # degenerate pairs: identical states, same position, aligned, pure arcs, axis-aligned headings
    V[:6] = [[0, 0, 0], [1, 1, 1], [0, 0, 0], [0, 0, 0], [0, 0, 0], [0.5, 0.5, math.pi / 2]]
    W[:6] = [[0, 0, 0], [1, 1, 2], [3, 0, 0], [1, 1, math.pi / 2], [0, 0, math.pi], [0.5, 1.5, math.pi / 2]]
    SS = _space(mp, kind, 0.7, 1.3)
    cost, nseg, segs = mp.car_steer_batch(SS, V, W)
    car = orc.SimpleCar(kind, 0.7, 1.3)
    for i in range(n):
        c, s = car.steer(V[i], W[i])
        assert cost[i].tobytes() == np.float64(c).tobytes(), i
        assert nseg[i] == len(s) and segs[i, :nseg[i]].tobytes() == s.tobytes(), i
    assert mp.steering_control(SS, V[7], W[7]) == [(t, (a, b)) for t, a, b in car.steer(V[7], W[7])[1]]


@pytest.mark.parametrize("kind", KINDS)
def test_tables_match_oracle(gpu, orc, kind):
    mp = gpu
    rng = np.random.Generator(np.random.PCG64(12))
    N, rturn, r = 2500, 0.06, 0.17
    V = _states(rng, N)
    SS = _space(mp, kind, rturn)
    mp.setup_steering(SS, r)
    car = orc.SimpleCar(kind, rturn)
    if kind == "reedsshepp":
        NN = mp.MetricNN(V, SS.dist, V[0])
        D = NN.precompute(r).D
        assert _same(D, car.inball(V, r, True))
        assert D.nnz > 5 * N
        col = mp.inballF(NN, 7, r)
        assert np.array_equal(col.nzind, mp.inballB(NN, 7, r).nzind)           # MetricNN: both are inball!
    else:
        NN = mp.QuasiMetricNN(V, SS.dist, V[0])
        cF, cB = NN.precompute(r)
        assert _same(cF.D, car.inball(V, r, True)) and _same(cB.D, car.inball(V, r, False))
        assert cF.D.nnz == cB.D.nnz > 2 * N                                    # B is the transpose pattern of F
    NN.close()


def test_tables_sharded_query_range_and_chop_value(gpu, orc):
    mp = gpu
    rng = np.random.Generator(np.random.PCG64(13))
    N, rturn, r = 1800, 0.05, 0.2
    V = _states(rng, N)
    SS = _space(mp, "dubins", rturn)
    SS.dist.chopval = 0.15                                                     # chop below r: primitivetypes.jl:95-100
    car = orc.SimpleCar("dubins", rturn)
    NN = mp.QuasiMetricNN(V, SS.dist, V[0])
    NN.set_query_range(500, 1300)
    cF, cB = NN.precompute(r)
    assert _same(cF.D, car.inball(V, r, True, chopval=0.15, q0=500, q1=1300))
    assert _same(cB.D, car.inball(V, r, False, chopval=0.15, q0=500, q1=1300))
    assert cF.D.nzval.max() <= 0.15
    NN.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("obst", ["sat2d", "boxes"])
def test_edge_validity_and_check_counts(gpu, orc, kind, obst):
    mp = gpu
    rng = np.random.Generator(np.random.PCG64(14))
    N, rturn, r = 1500, 0.04, 0.2
    V = _states(rng, N)
    SS = _space(mp, kind, rturn, 0.8)
    mp.setup_steering(SS, r)
    car = orc.SimpleCar(kind, rturn, 0.8)
    So = orc.StateSpace(SS.lo, SS.hi, ("view", [1, 2]))
    if obst == "sat2d":
        CC, O = mp.PointRobot2D(fx.product_shape(mp, fx.ISRR_POLY_WITH_SPIKE)), orc.Obstacles2D(fx.ISRR_POLY_WITH_SPIKE)
    else:
        CC, O = mp.PointRobotNDBoxes([mp.BoxBounds(b) for b in fx.BOXES2D]), orc.Boxes(fx.BOXES2D)
    if kind == "reedsshepp":
        NN = mp.MetricNN(V, SS.dist, V[0])
        D = NN.precompute(r).D
    else:
        NN = mp.QuasiMetricNN(V, SS.dist, V[0])
        D = NN.precompute(r)[1].D
    bits, checks = NN.car_edges_free(CC, SS)
    exp, cnt = car.edges_free_csc(O, So, V, D.colptr, D.rowval)
    got = unpack_bits(bits, D.nnz)
    assert np.array_equal(got, exp.astype(bool)) and checks == cnt
    assert 0.2 < got.mean() < 0.98
    # the state-level batch (what the lazy planner and is_free_path use)
    sel = rng.integers(0, D.nnz, 400)
    cols = np.searchsorted(D.colptr, sel + 1, side="right") - 1
    ys = D.rowval[sel] - 1
    CC.count = 0
    ok = mp.car_motions_free(V[ys], V[cols], CC, SS)
    assert np.array_equal(ok, got[sel])
    assert CC.count == sum(car.is_free_motion(O, So, V[y], V[x])[1] for y, x in zip(ys, cols))
    assert mp.is_free_motion(V[ys[0]], V[cols[0]], CC, SS) == bool(got[sel[0]])
    NN.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("edge_checks", ["table", "lazy"])
def test_fmt_over_a_car_space_matches_the_oracle(gpu, orc, kind, edge_checks):
    mp = gpu
    N, rturn, r = 1500, 0.05, 0.25
    rng = np.random.Generator(np.random.PCG64(15))
    SS = _space(mp, kind, rturn)
    So = orc.StateSpace(SS.lo, SS.hi, ("view", [1, 2]))
    B = orc.Boxes(fx.BOXES2D)
    cand = _states(rng, 3 * N)
    cand = cand[orc.states_free(B, So, cand)][:N - 2]
    init, goal = np.array([0.1, 0.1, 0.0]), np.array([0.9, 0.9, 0.0])
    V = np.vstack([init, cand, goal])
    CC = mp.PointRobotNDBoxes([mp.BoxBounds(b) for b in fx.BOXES2D])
    NNcls = mp.MetricNN if kind == "reedsshepp" else mp.QuasiMetricNN
    P = mp.MPProblem(SS, init, mp.StateGoal(goal), CC, V=NNcls(V, SS.dist, init))
    status, cost, _ = mp.fmtstar(P, r=r, edge_checks=edge_checks)
    car = orc.SimpleCar(kind, rturn)
    TF = car.inball(V, r, True)
    TB = TF if kind == "reedsshepp" else car.inball(V, r, False)        # MetricNN serves both from inball!
    col = lambda T, v: (T[1][T[0][v - 1] - 1:T[0][v] - 1], T[2][T[0][v - 1] - 1:T[0][v] - 1])
    ref = fmt_oracle(V, np.all(V == goal, axis=1), lambda v: col(TF, v), lambda v: col(TB, v),
                     lambda i: bool(orc.states_free(B, So, V[i:i + 1])[0]),
                     lambda y0, x0: car.is_free_motion(B, So, V[y0], V[x0]))
    assert ref["solved"] and status == "solved"
    assert P.solution.metadata["path"] == ref["path"] and np.array_equal(P.solution.metadata["tree"], ref["tree"])
    assert abs(cost - ref["cost"]) <= 1e-12 * ref["cost"]
    assert cost >= math.hypot(0.8, 0.8)
    assert mp.is_free_path(V[np.asarray(ref["path"]) - 1], CC, SS)
    P.V.close()


def test_degenerate_sample_sets_and_argument_errors(gpu, orc):
    mp = gpu
    SS = _space(mp, "dubins", 0.1)
    car = orc.SimpleCar("dubins", 0.1)
    for V in (np.zeros((1, 3)), np.array([[0.2, 0.2, 0.0], [0.25, 0.2, 0.0]]), np.array([[0.5, 0.5, 1.0]] * 3)):
        NN = mp.QuasiMetricNN(V, SS.dist, V[0])
        cF, cB = NN.precompute(0.3)
        assert _same(cF.D, car.inball(V, 0.3, True)) and _same(cB.D, car.inball(V, 0.3, False))
        NN.close()
    bad = mp.DubinsQuasiMetricSpace(-1.0)
    NN = mp.QuasiMetricNN(np.zeros((4, 3)), bad.dist)
    with pytest.raises(mp.MPB200Error, match="turning radius"):
        NN.precompute(0.1)
    NN.close()
    NN2 = mp.QuasiMetricNN(np.zeros((4, 2)), SS.dist)
    with pytest.raises(mp.MPB200Error, match="SE2"):
        NN2.precompute(0.1)
    NN2.close()


@pytest.mark.parametrize("kind", KINDS)
def test_fmtstar_samples_and_solves_a_car_problem_from_scratch(gpu, kind):
    """fmtstar!(P, N) on a car space with nothing precomputed: defaultNN (statespaces.jl:166-170), sample_free!
    (SE2 states, goal samples), tables, checks, path: the reference's notebook flow"""
    mp = gpu
    SS = _space(mp, kind, 0.05)
    CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H(), fixed_point_test=True)
    P = mp.MPProblem(SS, [0.1, 0.1, 0.0], mp.BallGoal([0.9, 0.9], 0.08), CC)
    status, cost, _ = mp.fmtstar(P, 2500, r=0.2, ensure_goal_ct=5, seed=3)
    assert type(P.V).__name__ == ("MetricNN" if kind == "reedsshepp" else "QuasiMetricNN") and len(P.V) == 2500
    assert status == "solved" and math.hypot(0.8, 0.8) - 0.08 <= cost < 4.0
    path = P.solution.metadata["path"]
    assert path[0] == 1 and mp.is_free_path(P.V.V[np.asarray(path) - 1], CC, SS)
    P.V.close()


@pytest.mark.parametrize("kind", KINDS)
def test_fmt_k_nearest_connections_over_a_car_space(gpu, orc, kind):
    """connections = :K on the car spaces: tables = k cheapest of the chopped-metric balls at the radius the growth loop
    ends with (chop value = that radius), mutual forward neighbourhoods, same planner loop over the oracle's tables"""
    mp = gpu
    N, rturn, r0, k = 1200, 0.05, 0.12, 12
    rng = np.random.Generator(np.random.PCG64(16))
    SS = _space(mp, kind, rturn)
    So = orc.StateSpace(SS.lo, SS.hi, ("view", [1, 2]))
    B = orc.Boxes(fx.BOXES2D)
    cand = _states(rng, 3 * N)
    cand = cand[orc.states_free(B, So, cand)][:N - 2]
    init, goal = np.array([0.1, 0.1, 0.0]), np.array([0.9, 0.9, 0.0])
    V = np.vstack([init, cand, goal])
    CC = mp.PointRobotNDBoxes([mp.BoxBounds(b) for b in fx.BOXES2D])
    NNcls = mp.MetricNN if kind == "reedsshepp" else mp.QuasiMetricNN
    P = mp.MPProblem(SS, init, mp.StateGoal(goal), CC, V=NNcls(V, SS.dist, init))
    status, cost, _ = mp.fmtstar(P, r=r0, connections="K", k=k)
    r = P.solution.metadata["r"]
    assert r > r0                                          # the ball had to grow to hold k neighbours everywhere
    car = orc.SimpleCar(kind, rturn)
    TF = car.inball(V, r, True)
    TB = TF if kind == "reedsshepp" else car.inball(V, r, False)
    assert np.diff(TF[0]).min() >= k and np.diff(TB[0]).min() >= k
    kF, kB = orc.knn_of_table(*TF, k), orc.knn_of_table(*TB, k)
    M = orc.union_transpose(kF, kB, len(V))
    col = lambda T, v: (T[1][T[0][v - 1] - 1:T[0][v] - 1], T[2][T[0][v - 1] - 1:T[0][v] - 1])
    ref = fmt_oracle(V, np.all(V == goal, axis=1), lambda v: col(M, v), lambda v: col(kB, v),
                     lambda i: bool(orc.states_free(B, So, V[i:i + 1])[0]),
                     lambda y0, x0: car.is_free_motion(B, So, V[y0], V[x0]))
    assert status == ("solved" if ref["solved"] else "failed")
    assert P.solution.metadata["path"] == ref["path"] and np.array_equal(P.solution.metadata["tree"], ref["tree"])
    if ref["solved"]:
        assert abs(cost - ref["cost"]) <= 1e-12 * ref["cost"]
    P.V.close()
