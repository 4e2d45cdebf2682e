"""Pins the oracle's double-integrator steering restatement (oracle/lq.c) against golden vectors
produced by re-running the reference's SymPy construction (tests/golden/gen_lq_golden.py,
linearquadratic.jl:94-157) and checks topt_newton / steer_pairwise semantics."""
import json
import os

import numpy as np

import fixtures as fx

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lq_di2.json")))


def test_closures_match_sympy_golden(orc):
    L = orc.DoubleIntegratorLQ(2)
    for c in GOLD["cases"]:
        got = L.cost_terms(c["x0"], c["x1"], c["t"])
        # tolerance: 1e-11 relative to the magnitude of the terms (cancellation in alpha/t^3 - beta/t^2)
        scale = max(1.0, abs(c["cost"]))
        assert abs(got[0] - c["cost"]) <= 1e-11 * scale
        assert abs(got[1] - c["dcost"]) <= 1e-10 * max(1.0, abs(c["dcost"]), scale / c["t"])
        assert abs(got[2] - c["ddcost"]) <= 1e-10 * max(1.0, abs(c["ddcost"]), scale / c["t"] ** 2)
        xs = L.state(c["x0"], c["x1"], c["t"], c["s"])
        assert np.allclose(xs, c["xofs"], rtol=0, atol=1e-11 * max(1.0, np.abs(c["xofs"]).max()))


def test_state_boundary_conditions(orc):
    L = orc.DoubleIntegratorLQ(2)
    for c in GOLD["cases"][:40]:
        x0, x1, t = np.array(c["x0"]), np.array(c["x1"]), c["t"]
        assert np.array_equal(L.state(x0, x1, t, 0.0), x0)            # x(0) = x exactly
        assert np.allclose(L.state(x0, x1, t, t), x1, atol=1e-12)     # x(t) = y up to rounding


def test_topt_newton_matches_reference_iteration(orc):
    L = orc.DoubleIntegratorLQ(2)
    n_inner = 0
    for c in GOLD["cases"]:
        cost, t = L.steer(c["x0"], c["x1"], c["r"])
        # same safeguarded-Newton iteration on closures that agree to ~1e-12 -> same iterate sequence;
        # the stopping test is on |dcost| <= 1e-6, so t agrees far tighter than the tolerance
        assert abs(t - c["topt"]) <= 1e-9 * max(1.0, c["r"])
        assert abs(cost - c["cost_at_topt"]) <= 1e-10 * max(1.0, abs(c["cost_at_topt"]))
        n_inner += t < c["r"]
    assert 20 < n_inner < len(GOLD["cases"])                           # both branches exercised


def test_steer_identical_states(orc):
    L = orc.DoubleIntegratorLQ(2)
    assert L.steer([.1, .2, .3, .4], [.1, .2, .3, .4], 1.0) == (0.0, 0.0)   # linearquadratic.jl:192


def test_lq_inball_semantics(orc):
    rng = np.random.Generator(np.random.PCG64(5))
    V = np.hstack([rng.random((300, 2)), (rng.random((300, 2)) - 0.5) * 1.6])
    V[17] = V[3]                                                        # duplicate state
    L = orc.DoubleIntegratorLQ(2)
    r = 0.9
    F = L.inball(V, r, True)
    B = L.inball(V, r, False)
    N = len(V)
    colsF = np.repeat(np.arange(1, N + 1), np.diff(F[0]))
    colsB = np.repeat(np.arange(1, N + 1), np.diff(B[0]))
    # DSF = Dmat' and DSB = Dmat (linearquadratic.jl:73-74): same entries, transposed
    f = sorted(zip(colsF.tolist(), F[1].tolist(), F[2].tolist()))
    b = sorted(zip(B[1].tolist(), colsB.tolist(), B[2].tolist()))
    assert f == b and len(f) >= 40
    assert not (F[1] == colsF).any() and (F[2] <= r).all()
    # brute check of one column through steer()
    q = 5
    exp = [(j + 1, L.steer(V[q], V[j], r)[0]) for j in range(N) if j != q]
    exp = [(j, c) for j, c in exp if c <= r and L.cost_terms(V[q], V[j - 1], r)[1] > 0 or (V[q] == V[j - 1]).all()]
    got = list(zip(F[1][F[0][q] - 1:F[0][q + 1] - 1].tolist(), F[2][F[0][q] - 1:F[0][q + 1] - 1].tolist()))
    assert got == [(j, c) for j, c in exp if c <= r]
    # the duplicate pair is a neighbour at cost 0 in both directions (Q9 analogue for distinct indices)
    col = F[1][F[0][3] - 1:F[0][4] - 1].tolist()
    slow = 12 * (V[3, 2:] ** 2).sum() < r * r                          # prefilter dc(r) = 1 - gamma/r^2 > 0
    assert (18 in col) == bool(slow)


def test_lq_edge_waypoints(orc):
    L = orc.DoubleIntegratorLQ(2)
    B = orc.Boxes(fx.BOXES2D)
    C = np.hstack([np.eye(2), np.zeros((2, 2))])
    S = orc.StateSpace([0, 0, -1.5, -1.5], [1, 1, 1.5, 1.5], ("matrix", C))
    # straight slow motion in free space: 4 segment checks, free
    ok, cnt = L.is_free_motion(B, S, 1.0, [0.1, 0.1, 0.3, 0.0], [0.3, 0.1, 0.3, 0.0])
    assert ok and cnt == 4
    # through box 2 ([.4,.5]x[.19,.35])
    ok, cnt = L.is_free_motion(B, S, 1.0, [0.35, 0.25, 0.5, 0.0], [0.6, 0.25, 0.5, 0.0])
    assert not ok and 1 <= cnt <= 4
    # start state outside the velocity bounds: no segment check at all (statespaces.jl:155)
    ok, cnt = L.is_free_motion(B, S, 1.0, [0.1, 0.1, 2.0, 0.0], [0.3, 0.1, 0.3, 0.0])
    assert not ok and cnt == 0
