"""Pins oracle/cars.c (the restatement of src/statespaces/simplecars.jl) without a GPU:
elementary routines against libm, mod2piF against Julia's mod semantics, Dubins against an INDEPENDENT geometric
construction (tangent lines between turning circles), closed-form known answers, and for both metrics the consistency
properties that any correct implementation must have (the returned control, propagated, arrives at the target; its
durations add up to the cost; Reeds-Shepp is symmetric, invariant under rigid motions and never longer than the four
Dubins-type bounds; triangle inequality)."""
import math

import numpy as np
import pytest

from oracle import oracle as orc

TWO_PI = 2 * math.pi


def _rand_states(rng, n, box=3.0):
    return np.column_stack([rng.uniform(0, box, n), rng.uniform(0, box, n), rng.uniform(0, TWO_PI, n)])


def test_elementary_functions_are_within_a_few_ulp_of_libm():
    L = orc.lib()
    rng = np.random.default_rng(1)
    for x in rng.uniform(-60, 60, 20000):
        assert abs(L.orc_det_sin(x) - math.sin(x)) <= 2.3e-16
        assert abs(L.orc_det_cos(x) - math.cos(x)) <= 2.3e-16
    for y, x in rng.normal(size=(20000, 2)):
        assert abs(L.orc_det_atan2(y, x) - math.atan2(y, x)) <= 9e-16
    for u in np.concatenate([rng.uniform(-1, 1, 20000), [-1.0, 1.0, 0.0]]):
        assert abs(L.orc_det_acos(u) - math.acos(u)) <= 9e-16
    # quadrant conventions of atan2, including the axes
    for y, x in [(0.0, 1.0), (1.0, 0.0), (0.0, -1.0), (-1.0, 0.0), (-0.0, -1.0), (1.0, 1.0), (-1.0, -1.0), (0.0, 0.0)]:
        assert L.orc_det_atan2(y, x) == pytest.approx(math.atan2(y, x), abs=5e-16)


def test_mod2pi_is_julias_mod():
    L = orc.lib()
    rng = np.random.default_rng(2)
    for x in rng.uniform(-100, 100, 20000):
        m = L.orc_mod2pi(x)
        assert 0.0 <= m <= TWO_PI and m == np.mod(x, TWO_PI)      # numpy's mod has the same floored definition
    assert L.orc_mod2pi(0.0) == 0.0 and math.copysign(1.0, L.orc_mod2pi(-0.0)) == 1.0
    assert L.orc_mod2pi(TWO_PI) == 0.0 and L.orc_mod2pi(-TWO_PI) == 0.0
    assert L.orc_mod2pi(-1e-20) == TWO_PI                       # r + y rounds to y, as in Julia


# ---- an independent Dubins: tangent construction between the turning circles --------------------------------------
def _dubins_geometric(v, w, r):
    best = math.inf

    def centre(s, side):      # side +1 = left circle, -1 = right circle
        return np.array([s[0] - side * r * math.sin(s[2]), s[1] + side * r * math.cos(s[2])])

    def arc(a0, a1, side):    # angle travelled from heading a0 to heading a1 turning left (+1) / right (-1)
        return (side * (a1 - a0)) % TWO_PI

    for s1 in (1, -1):
        for s2 in (1, -1):
            c1, c2 = centre(v, s1), centre(w, s2)
            d = c2 - c1
            D = math.hypot(*d)
            if s1 == s2:
                if D < 1e-12:
                    cand = [v[2]]
                else:
                    cand = [math.atan2(d[1], d[0])]
                length = D
            else:
                if D < 2 * r:
                    continue
                length = math.sqrt(max(D * D - 4 * r * r, 0.0))
                # heading of the inner tangent
                cand = [math.atan2(d[1], d[0]) + s1 * math.asin(min(1.0, 2 * r / D))]
            for th in cand:
                tot = r * arc(v[2], th, s1) + length + r * arc(th, w[2], s2)
                best = min(best, tot)
    for s in (1, -1):         # CCC: the middle circle is tangent to both
        c1, c2 = centre(v, s), centre(w, s)
        d = c2 - c1
        D = math.hypot(*d)
        if D < 4 * r and D > 1e-12:
            base = math.atan2(d[1], d[0])
            alpha = math.acos(D / (4 * r))
            for sg in (1, -1):
                c3 = c1 + 2 * r * np.array([math.cos(base + sg * alpha), math.sin(base + sg * alpha)])
                # tangent points and headings
                p1 = (c1 + c3) / 2
                p2 = (c2 + c3) / 2
                h1 = math.atan2(p1[1] - c1[1], p1[0] - c1[0]) + s * math.pi / 2
                h2 = math.atan2(p2[1] - c2[1], p2[0] - c2[0]) + s * math.pi / 2
                tot = r * (arc(v[2], h1, s) + arc(h1, h2, -s) + arc(h2, w[2], s))
                best = min(best, tot)
    return best


def test_dubins_matches_an_independent_geometric_construction():
    rng = np.random.default_rng(3)
    car = orc.SimpleCar("dubins", 0.6, 1.0)
    V, W = _rand_states(rng, 3000), _rand_states(rng, 3000)
    worst = 0.0
    for v, w in zip(V, W):
        c, _ = car.steer(v, w)
        g = _dubins_geometric(v, w, 0.6)
        worst = max(worst, abs(c - g) / max(1.0, g))
    assert worst < 1e-9


def test_known_answers():
    D, RS = orc.SimpleCar("dubins", 1.0), orc.SimpleCar("reedsshepp", 1.0)
    assert D.steer([0, 0, 0], [3, 0, 0])[0] == pytest.approx(3.0, abs=1e-14)
    # pure arcs are DEGENERATE for the reference's closed forms (atan2(~0, ~0), mod2piF of an angle that is 0 up to
    # rounding): the result is right modulo a full turn, as it would be in the reference -- reproduced, not fixed
    for tgt, ang in (([1, 1, math.pi / 2], math.pi / 2), ([0, 2, math.pi], math.pi), ([0, -2, math.pi], math.pi)):
        c = D.steer([0, 0, 0], tgt)[0]
        assert min((c - ang) % TWO_PI, TWO_PI - (c - ang) % TWO_PI) < 1e-12
    # composite words built forwards: left phi, straight p, left/right psi
    for s2, name in ((1, "LSL"), (-1, "LSR")):
        x = np.zeros(3)
        for u in ((0.5, 1.0, 1.0), (1.2, 1.0, 0.0), (0.7, 1.0, float(s2))):
            x = D.propagate(x, u)
        assert D.steer([0, 0, 0], x)[0] == pytest.approx(2.4, abs=1e-12), name
        assert RS.steer([0, 0, 0], x)[0] == pytest.approx(2.4, abs=1e-12), name
    assert D.steer([0, 0, 0], [-1, 0, 0])[0] > 2 * math.pi - 1e-9                                   # cannot reverse
    assert RS.steer([0, 0, 0], [-1, 0, 0])[0] == pytest.approx(1.0, abs=1e-14)                      # can
    assert RS.steer([0, 0, 0], [2.5, 0, 0])[0] == pytest.approx(2.5, abs=1e-14)
    assert RS.steer([0, 0, 0], [0, 0, math.pi])[0] == pytest.approx(math.pi, abs=1e-12)             # turn in place: C|C|C
    # the turning radius scales lengths
    for car_kind in ("dubins", "reedsshepp"):
        a, b = orc.SimpleCar(car_kind, 1.0), orc.SimpleCar(car_kind, 0.25)
        v, w = np.array([0.3, 0.2, 1.0]), np.array([1.4, 2.0, 4.0])
        assert b.steer([0.25 * v[0], 0.25 * v[1], v[2]], [0.25 * w[0], 0.25 * w[1], w[2]])[0] == \
            pytest.approx(0.25 * a.steer(v, w)[0], rel=1e-13)


@pytest.mark.parametrize("kind", ["dubins", "reedsshepp"])
def test_the_steering_control_arrives_and_its_durations_add_up(kind):
    rng = np.random.default_rng(4)
    car = orc.SimpleCar(kind, 0.7, 1.3)
    for v, w in zip(_rand_states(rng, 4000), _rand_states(rng, 4000)):
        c, segs = car.steer(v, w)
        assert len(segs) == 3 if kind == "dubins" else 3 <= len(segs) <= 5
        x = v.copy()
        for u in segs:
            assert u[0] >= 0 and abs(u[1]) in (0.0, 1.3) and abs(u[2]) in (0.0, 1 / 0.7)
            x = car.propagate(x, u)
        assert abs(x[0] - w[0]) < 1e-12 and abs(x[1] - w[1]) < 1e-12
        assert abs(math.remainder(x[2] - w[2], TWO_PI)) < 1e-12
        assert sum(u[0] * abs(u[1]) for u in segs) == pytest.approx(c, rel=1e-12, abs=1e-12)   # length = sum |speed| dt


def test_reeds_shepp_metric_properties():
    rng = np.random.default_rng(5)
    RS, D = orc.SimpleCar("reedsshepp", 0.5), orc.SimpleCar("dubins", 0.5)
    A, B, C = _rand_states(rng, 1500), _rand_states(rng, 1500), _rand_states(rng, 1500)
    flip = lambda s: np.array([s[0], s[1], (s[2] + math.pi) % TWO_PI])
    for a, b, c in zip(A, B, C):
        ab = RS.steer(a, b)[0]
        assert ab == pytest.approx(RS.steer(b, a)[0], rel=1e-10, abs=1e-12)                      # symmetric
        assert ab >= math.hypot(a[0] - b[0], a[1] - b[1]) - 1e-12                                # lower bound used by the chop
        ub = min(D.steer(a, b)[0], D.steer(b, a)[0], D.steer(flip(a), flip(b))[0], D.steer(flip(b), flip(a))[0])
        assert ab <= ub + 1e-10                                                                   # forward-only / reverse-only paths
        assert ab <= RS.steer(a, c)[0] + RS.steer(c, b)[0] + 1e-10                               # triangle inequality
        # rigid motion invariance
        th = rng.uniform(0, TWO_PI)
        R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        mv = lambda s: np.concatenate([R @ s[:2] + [0.3, -0.2], [(s[2] + th) % TWO_PI]])
        assert RS.steer(mv(a), mv(b))[0] == pytest.approx(ab, rel=1e-9, abs=1e-11)


def test_chopped_metric_and_inball():
    """primitivetypes.jl:95-100 and nearneighbors.jl:185-198"""
    rng = np.random.default_rng(6)
    car = orc.SimpleCar("dubins", 0.1)
    V = np.column_stack([rng.random(300), rng.random(300), rng.uniform(0, TWO_PI, 300)])
    r = 0.3
    for fw in (True, False):
        cp, rv, nz = car.inball(V, r, fw)
        for q in (0, 17, 299):
            exp = []
            for i in range(300):
                if i == q or (V[q, 0] - V[i, 0]) ** 2 + (V[q, 1] - V[i, 1]) ** 2 > r * r:
                    continue
                d = car.steer(V[q], V[i])[0] if fw else car.steer(V[i], V[q])[0]
                if d <= r:
                    exp.append((i + 1, d))
            got = list(zip(rv[cp[q] - 1:cp[q + 1] - 1], nz[cp[q] - 1:cp[q + 1] - 1]))
            assert got == exp
    assert car.chopped([0, 0, 0], [0.5, 0, 0], 0.4) == math.inf            # lower bound beyond the chop value
    assert car.chopped([0, 0, 0], [-0.05, 0, 0], 0.4) == math.inf          # exact length (a loop) beyond it
    assert car.chopped([0, 0, 0], [0.3, 0, 0], 0.4) == pytest.approx(0.3, abs=1e-15)


def test_is_free_motion_waypoints_and_count():
    """statespaces.jl:153-158 over collision_waypoints (simplecars.jl:71-82, statespaces.jl:134-142)"""
    car = orc.SimpleCar("dubins", 0.1)
    S = orc.StateSpace([0, 0, 0], [1, 1, TWO_PI], ("view", [1, 2]))
    free = orc.Obstacles2D(("compound", []))
    # straight ahead: LSL with two empty arcs -> waypoints [v, v, end, w] -> 3 swept segments
    ok, n = car.is_free_motion(free, S, [0.1, 0.5, 0.0], [0.9, 0.5, 0.0])
    assert ok and n == 3
    wall = orc.Obstacles2D(("circle", (0.5, 0.5), 0.1))
    ok, n = car.is_free_motion(wall, S, [0.1, 0.5, 0.0], [0.9, 0.5, 0.0])
    assert not ok and n == 2                                                  # the first (degenerate) segment passes
    # a left U-turn of radius 0.1: heading change pi -> 12 arc waypoints after the start of the arc
    ok, n = car.is_free_motion(free, S, [0.5, 0.4, 0.0], [0.5, 0.6, math.pi])
    assert ok and n >= 12
    # ... which leaves the state space when started next to the boundary (bounds-checked waypoints)
    ok, _ = car.is_free_motion(free, S, [0.95, 0.4, 0.0], [0.95, 0.6, math.pi])
    assert not ok
    # right turns get NO intermediate waypoints (floor of a negative heading change, simplecars.jl:75-77): the chord
    # of the U-turn is swept instead of the arc, so an obstacle inside the arc but off the chord is missed
    bump = orc.Obstacles2D(("circle", (0.6, 0.5), 0.02))
    assert car.is_free_motion(bump, S, [0.5, 0.6, 0.0], [0.5, 0.4, math.pi])[0]          # right U-turn: missed
    assert not car.is_free_motion(bump, S, [0.5, 0.4, 0.0], [0.5, 0.6, math.pi])[0]      # left U-turn: caught


def test_golden_vectors_from_the_reference_formulas_in_40_digit_arithmetic():
    """tests/golden/cars.json (gen_cars_golden.py: the reference's formulas and sweep order re-run in mpmath)"""
    import json
    import os
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cars.json")))
    rt = G["turning_radius"]
    compared_controls = 0
    for case in G["cases"]:
        v = np.array([float.fromhex(x) for x in case["v"]])
        w = np.array([float.fromhex(x) for x in case["w"]])
        for kind in ("dubins", "reedsshepp"):
            g = case[kind]
            c, segs = orc.SimpleCar(kind, rt).steer(v, w)
            gc = float(g["cost"])
            # float64 evaluation of lengths of a few units: absolute 1e-12 ... except next to the formulas' own
            # discontinuities (mod2pi of an angle within rounding of 0 / 2pi adds a full turn: reproduced, not compared)
            if abs(c - gc) > 1e-11 * max(1.0, gc):
                assert abs(abs(c - gc) - TWO_PI * rt) < 1e-9 or g["gap"] < 1e-9, (kind, case["v"], case["w"], c, gc)
                continue
            if g["gap"] > 1e-7:          # the winner is unambiguous: same word, same directions, same durations
                ctrl = g["control"]
                assert len(segs) == len(ctrl)
                for s, (t, u1, u2) in zip(segs, ctrl):
                    assert abs(s[0] - float(t)) < 1e-10 and s[1] == float(u1) and abs(s[2] - float(u2)) < 1e-12
                compared_controls += 1
    assert compared_controls > 350
