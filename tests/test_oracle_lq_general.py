"""Pins the oracle's GENERAL linear-affine steering restatement (oracle/lq_general.c: numeric G(t), Cholesky
solves) against golden vectors produced by re-running the reference's SymPy construction
(tests/golden/gen_lq_general_golden.py, linearquadratic.jl:94-157) for a drifting double integrator and a
triple integrator, and against the double-integrator closed form (oracle/lq.c)."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lq_general.json")))["systems"]


def _system(orc, name):
    g = GOLD[name]
    return orc.LinearQuadraticGeneral(np.array(g["A"]), np.array(g["B"]), np.array(g["c"]), np.array(g["R"])), g["cases"]


@pytest.mark.parametrize("name", sorted(GOLD))
def test_closures_match_sympy_golden(orc, name):
    L, cases = _system(orc, name)
    for c in cases:
        got = L.cost_terms(c["x0"], c["x1"], c["t"])
        scale = max(1.0, abs(c["cost"]))
        # G(t) is ill-conditioned like t^-(2n-2): the numeric solve loses what the symbolic inverse does not
        assert abs(got[0] - c["cost"]) <= 1e-9 * scale
        assert abs(got[1] - c["dcost"]) <= 1e-9 * max(1.0, abs(c["dcost"]), scale / c["t"])
        assert abs(got[2] - c["ddcost"]) <= 1e-9 * max(1.0, abs(c["ddcost"]), scale / c["t"] ** 2)
        xs = L.state(c["x0"], c["x1"], c["t"], c["s"])
        assert np.allclose(xs, c["xofs"], rtol=0, atol=1e-9 * max(1.0, np.abs(c["xofs"]).max()))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_topt_newton_and_boundary_conditions(orc, name):
    L, cases = _system(orc, name)
    inner = 0
    for c in cases:
        cost, t = L.steer(c["x0"], c["x1"], c["r"])
        assert abs(t - c["topt"]) <= 1e-7 * max(1.0, c["r"])
        assert abs(cost - c["cost_at_topt"]) <= 1e-8 * max(1.0, abs(c["cost_at_topt"]))
        inner += t < c["r"]
        x0, x1 = np.array(c["x0"]), np.array(c["x1"])
        assert np.allclose(L.state(x0, x1, c["t"], 0.0), x0, atol=1e-12)
        assert np.allclose(L.state(x0, x1, c["t"], c["t"]), x1, atol=1e-9)
    assert 5 < inner < len(cases)


def test_general_path_reproduces_the_double_integrator_closed_form(orc):
    d = 2
    A = np.block([[np.zeros((d, d)), np.eye(d)], [np.zeros((d, 2 * d))]])
    B = np.vstack([np.zeros((d, d)), np.eye(d)])
    R = np.array([[1.5, 0.2], [0.2, 0.8]])
    G, D = orc.LinearQuadraticGeneral(A, B, np.zeros(4), R), orc.DoubleIntegratorLQ(2, R)
    rng = np.random.Generator(np.random.PCG64(3))
    for _ in range(200):
        x = rng.random(4) * [1, 1, 3, 3] - [0, 0, 1.5, 1.5]
        y = x + (rng.random(4) - 0.5) * [0.4, 0.4, 1.0, 1.0]
        t = 0.1 + rng.random()
        a, b = G.cost_terms(x, y, t), D.cost_terms(x, y, t)
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9)
        assert np.allclose(G.steer(x, y, 0.9), D.steer(x, y, 0.9), rtol=1e-8, atol=1e-9)


def test_setup_rejects_what_the_reference_rejects(orc):
    with pytest.raises(ValueError):      # linearquadratic.jl:96 -- nilpotent A only
        orc.LinearQuadraticGeneral(np.array([[0.0, 1.0], [-1.0, 0.0]]), np.array([[0.0], [1.0]]), np.zeros(2), np.eye(1))
    with pytest.raises(ValueError):      # R must be positive definite
        orc.LinearQuadraticGeneral(np.array([[0.0, 1.0], [0.0, 0.0]]), np.array([[0.0], [1.0]]), np.zeros(2), -np.eye(1))


def test_general_inball_semantics(orc):
    L, _ = _system(orc, "triple")
    rng = np.random.Generator(np.random.PCG64(8))
    V = np.array([0, -1, -2]) + rng.random((250, 3)) * np.array([1, 2, 4])
    V[11] = V[5]
    r = 1.4
    F, Bk = L.inball(V, r, True), L.inball(V, r, False)
    assert F[0][-1] == Bk[0][-1] > 1                     # DSF = DSB'
    pairs_f = {(q, j - 1) for q in range(250) for j in F[1][F[0][q] - 1:F[0][q + 1] - 1]}
    pairs_b = {(j - 1, q) for q in range(250) for j in Bk[1][Bk[0][q] - 1:Bk[0][q + 1] - 1]}
    assert pairs_f == pairs_b
    # duplicate states cost 0 (linearquadratic.jl:192) but, with drift, are stored only if they pass the dense
    # prefilter dcost(r) > 0 (:213) -- the restatement keeps that order
    assert ((5, 11) in pairs_f) == (L.cost_terms(V[5], V[11], r)[1] > 0)
    for q in (0, 100, 249):
        rows = F[1][F[0][q] - 1:F[0][q + 1] - 1]
        assert np.all(np.diff(rows) > 0) and q + 1 not in rows
        for j, c in zip(rows, F[2][F[0][q] - 1:F[0][q + 1] - 1]):
            assert c <= r and abs(L.steer(V[q], V[j - 1], r)[0] - c) == 0.0
