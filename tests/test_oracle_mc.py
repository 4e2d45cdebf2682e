"""Known-answer tests for the Monte-Carlo estimator's specification (oracle/mc.c; parity unpinned:
the estimator is absent from the reference) and for its building blocks."""
import math

import numpy as np

import fixtures as fx


def test_philox_random123_known_answers(orc):
    # Random123 kat_vectors for philox4x32-10
    assert [hex(x) for x in orc.philox4x32_10([0] * 4, [0] * 2)] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    assert [hex(x) for x in orc.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)] == \
        ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(x) for x in orc.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_deterministic_elementary_functions(orc):
    xs = np.random.default_rng(0).random(4000) * 0.999 + 1e-12
    for x in xs:
        assert abs(orc.det_log(x) - math.log(x)) <= 1e-15 * max(1.0, abs(math.log(x)))
        y = (x - 0.5) * 300
        assert abs(orc.det_exp(y) - math.exp(y)) <= 1e-15 * math.exp(y)
        s, c = orc.det_sincos2pi(x)
        assert abs(s - math.sin(2 * math.pi * x)) <= 2e-15 and abs(c - math.cos(2 * math.pi * x)) <= 2e-15
    assert orc.det_sincos2pi(0.0) == (0.0, 1.0) and orc.det_sincos2pi(0.25) == (1.0, 0.0)
    assert orc.det_exp(0.0) == 1.0 and orc.det_log(1.0) == 0.0


def half_plane_problem(orc, delta, sigma, T=1, shift=None, alpha0=0.3):
    """obstacle = box [0.5, 10] x [-10, 10]; nominal sits at x = 0.5 - delta; w_T = wbar + sigma*sum eps"""
    F = np.stack([np.eye(2)] * T)
    G = np.stack([np.eye(2) * sigma] * T)
    wbar = np.tile([0.5 - delta, 0.0], (T + 1, 1))
    if shift is None:
        return orc.McSpec(F, G, np.eye(2), wbar, [1.0], None)
    return orc.McSpec(F, G, np.eye(2), wbar, [alpha0, 1 - alpha0], [shift])


def Phi(x):
    return 0.5 * math.erfc(-x / math.sqrt(2))


def test_half_plane_single_step_known_answer(orc):
    B = orc.Boxes([(np.array([0.5, -10.0]), np.array([10.0, 10.0]))])
    delta, sigma, n = 0.2, 0.1, 200_000
    exact = Phi(-delta / sigma)                                   # 0.02275
    naive = orc.mc_run(half_plane_problem(orc, delta, sigma), B, 7, 0, n)
    p = naive["S1"] / n
    se = math.sqrt((naive["S2"] / n - p * p) / n)
    assert abs(p - exact) < 4 * se and naive["hits"] == round(naive["S1"]) and naive["S0"] == n
    # importance sampling: shift the x-noise onto the boundary (mu = delta/sigma)
    IS = orc.mc_run(half_plane_problem(orc, delta, sigma, shift=[delta / sigma, 0.0]), B, 7, 0, n)
    p2 = IS["S1"] / n
    se2 = math.sqrt((IS["S2"] / n - p2 * p2) / n)
    assert abs(p2 - exact) < 4 * se2 and se2 < 0.5 * se
    assert abs(IS["S0"] / n - 1) < 0.02                           # weights average to 1


def test_rare_event_needs_importance_sampling(orc):
    B = orc.Boxes([(np.array([0.5, -10.0]), np.array([10.0, 10.0]))])
    delta, sigma, n = 0.45, 0.1, 100_000
    exact = Phi(-delta / sigma)                                   # 3.4e-6
    IS = orc.mc_run(half_plane_problem(orc, delta, sigma, shift=[delta / sigma, 0.0]), B, 11, 0, n)
    p = IS["S1"] / n
    se = math.sqrt((IS["S2"] / n - p * p) / n)
    assert abs(p - exact) < 4 * se and se < 0.05 * exact


def test_multi_step_random_walk_and_rollout_independence(orc):
    B = orc.Boxes([(np.array([0.5, -10.0]), np.array([10.0, 10.0]))])
    T, sigma, delta = 4, 0.1, 0.3
    spec = half_plane_problem(orc, delta, sigma, T=T)
    a = orc.mc_run(spec, B, 3, 0, 3000, per_rollout=True)
    b1 = orc.mc_run(spec, B, 3, 0, 1000, per_rollout=True)
    b2 = orc.mc_run(spec, B, 3, 1000, 2000, per_rollout=True)
    # counter-based RNG: any partition of the id range reproduces the same rollouts
    assert np.array_equal(a["hit"], np.concatenate([b1["hit"], b2["hit"]]))
    assert a["hits"] == b1["hits"] + b2["hits"]
    # reflection-principle bound: P(max_t S_t >= delta) between P(S_T >= delta) and 2 P(S_T >= delta)
    n = 200_000
    r = orc.mc_run(spec, B, 5, 0, n)
    p = r["S1"] / n
    lo = Phi(-delta / (sigma * math.sqrt(T)))
    assert lo * 0.95 < p < 2 * lo * 1.05


def test_sat2d_event_and_swept_variant(orc):
    O = orc.Obstacles2D(("compound", [("circle", (0.5, 0.0), 0.1)]))
    F = np.stack([np.eye(2)] * 2)
    G = np.stack([np.eye(2) * 1e-3] * 2)
    wbar = np.array([[0.0, 0.0], [0.3, 0.0], [0.9, 0.0]])         # steps over the circle between t=1 and t=2
    pt = orc.mc_run(orc.McSpec(F, G, np.eye(2), wbar, [1.0], None, swept=False), O, 1, 0, 2000)
    sw = orc.mc_run(orc.McSpec(F, G, np.eye(2), wbar, [1.0], None, swept=True), O, 1, 0, 2000)
    assert pt["hits"] == 0 and sw["hits"] == 2000
