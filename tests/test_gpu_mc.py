"""GPU parity for K10 (Monte-Carlo collision probability) against the oracle's specification:
identical per-rollout hit bits and bit-identical importance weights; sums to 1e-12."""
import math

import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu


def di_lqg_problem(mp, T=30, dt=0.05):
    """2-D double integrator tracking a straight nominal path past ISRR_2H's box 2"""
    A = np.block([[np.eye(2), dt * np.eye(2)], [np.zeros((2, 2)), np.eye(2)]])
    B = np.vstack([0.5 * dt * dt * np.eye(2), dt * np.eye(2)])
    C = np.hstack([np.eye(2), np.zeros((2, 2))])
    F, G = mp.montecarlo.lqg_closed_loop(A, B, C, np.eye(4), 0.1 * np.eye(2), 1e-4 * np.eye(4), 1e-4 * np.eye(2), T)
    Wz = np.hstack([np.eye(2), np.zeros((2, 6))])
    wbar = np.stack([np.linspace(0.30, 0.62, T + 1), np.full(T + 1, 0.165)], axis=1)   # passes 0.025 under box 2
    return mp.MCProblem(F, G, Wz, wbar)


def to_spec(orc, P):
    return orc.McSpec(P.F, P.G, P.Wz, P.wbar, P.alpha, P.mu if P.K else None, P.swept)


@pytest.mark.parametrize("checker", ["boxes", "sat2d_fixed"])
@pytest.mark.parametrize("swept", [False, True])
def test_k10_matches_oracle_per_rollout(gpu, orc, checker, swept):
    mp = gpu
    if checker == "boxes":
        CC, O = mp.PointRobotNDBoxes([mp.BoxBounds(b) for b in fx.BOXES2D]), orc.Boxes(fx.BOXES2D)
    else:
        CC = mp.PointRobot2D(mp.obstaclesets.ISRR_2H(), fixed_point_test=True)
        O = orc.Obstacles2D(fx.ISRR_2H, fixed_point_test=True)
    P0 = di_lqg_problem(mp)
    P0.swept = swept
    P = mp.montecarlo.with_proposal(P0, CC, r2=25.0)
    assert P.K >= 1 and abs(P.alpha.sum() - 1) < 1e-12
    n = 20_000
    got = mp.collision_probability(P, CC, n, seed=20240605, first=12345, per_rollout=True)
    exp = orc.mc_run(to_spec(orc, P), O, 20240605, 12345, n, per_rollout=True)
    assert np.array_equal(got["hit"], exp["hit"].astype(bool))
    assert got["w"].tobytes() == exp["w"].tobytes()
    assert got["hits"] == exp["hits"] and got["n"] == n
    for k in ("S1", "S2", "S0"):
        assert abs(got[k] - exp[k]) <= 1e-12 * max(1.0, abs(exp[k]))
    assert 0 < got["hits"] < n


def test_k10_known_answer_and_shard_invariance(gpu, orc):
    mp = gpu
    CC = mp.PointRobotNDBoxes([mp.BoxBounds(np.array([0.5, -10.0]), np.array([10.0, 10.0]))])
    delta, sigma = 0.3, 0.1
    F, G = np.eye(2)[None], (np.eye(2) * sigma)[None]
    wbar = np.array([[0.5 - delta, 0.0]] * 2)
    P = mp.MCProblem(F, G, np.eye(2), wbar, [0.3, 0.7], [[delta / sigma, 0.0]])
    n = 1_000_000
    whole = mp.collision_probability(P, CC, n, seed=9)
    exact = 0.5 * math.erfc(delta / sigma / math.sqrt(2))            # Phi(-3) = 1.35e-3
    assert abs(whole["p"] - exact) < 4 * whole["se"] and whole["se"] < 0.01 * exact
    # shard the rollout ids over "8 GPUs": sums agree with the single launch to summation order
    parts = [mp.collision_probability(P, CC, n // 8, seed=9, first=g * (n // 8)) for g in range(8)]
    tot = mp.montecarlo.combine(parts)
    assert tot["hits"] == whole["hits"] and tot["n"] == n
    for k in ("S1", "S2", "S0"):
        assert abs(tot[k] - whole[k]) <= 1e-12 * abs(whole[k])
    # a given launch geometry is bit-reproducible
    again = mp.collision_probability(P, CC, n, seed=9)
    assert (again["S1"], again["S2"], again["S0"]) == (whole["S1"], whole["S2"], whole["S0"])
    # naive MC (no shifted component) agrees within its own error
    naive = mp.collision_probability(mp.MCProblem(F, G, np.eye(2), wbar), CC, n, seed=10)
    assert abs(naive["p"] - whole["p"]) < 4 * math.hypot(naive["se"], whole["se"])
    assert naive["S0"] == n
