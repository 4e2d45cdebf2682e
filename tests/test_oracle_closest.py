"""Pins the oracle's closest / closeR restatement (oracle/closest.c; SAT2D.jl:208-285, boxesND.jl:61-86): every
result is the minimiser of a strictly convex problem, so it is checked against an independent brute-force search
over a dense sampling of each obstacle, plus exact known answers."""
import numpy as np
import pytest

import fixtures as fx


def _Ws(rng, n, d):
    out = np.empty((n, d, d))
    for i in range(n):
        A = rng.standard_normal((d, d))
        out[i] = A @ A.T + 0.3 * np.eye(d)
    return out


def _brute_polygon(p, pts, W, m=4001):
    best = np.inf
    pts = np.asarray(pts)
    for i in range(len(pts)):
        a, b = pts[i], pts[(i + 1) % len(pts)]
        t = np.linspace(0, 1, m)[:, None]
        D = a + t * (b - a) - p
        best = min(best, np.einsum("ij,jk,ik->i", D, W, D).min())
    return best


def test_polygon_and_circle_against_brute_force(orc):
    rng = np.random.Generator(np.random.PCG64(21))
    O = orc.Obstacles2D(fx.ISRR_POLY)               # 3 polygons + 1 circle
    parts = fx.ISRR_POLY[1]
    n = 60
    P = rng.random((n, 2)) * 1.2 - 0.1
    Ws = _Ws(rng, n, 2)
    count, d2, shape, x, all_d2, all_x = orc.close_points(O, P, Ws, 1e9, want_all=True)
    assert (count == 4).all()
    for i in range(n):
        for s, part in enumerate(parts):
            g = all_x[i, s] - P[i]
            assert abs(g @ Ws[i] @ g - all_d2[i, s]) <= 1e-9 * max(1.0, all_d2[i, s])     # d2 is the W-distance of x
            if part[0] == "polygon":
                ref = _brute_polygon(P[i], part[1], Ws[i])
                assert all_d2[i, s] <= ref * (1 + 1e-9) + 1e-12 and all_d2[i, s] >= ref * (1 - 2e-3) - 1e-9
            else:
                c, r = np.array(part[1]), part[2]
                th = np.linspace(0, 2 * np.pi, 200001)
                D = c + r * np.stack([np.cos(th), np.sin(th)], axis=1) - P[i]
                ref = np.einsum("ij,jk,ik->i", D, Ws[i], D).min()
                inside = np.linalg.norm(P[i] - c) < r
                if not inside:                          # the Newton iteration targets the boundary point from outside
                    assert abs(all_d2[i, s] - ref) <= 1e-6 * max(ref, 1e-6)
                    assert abs(np.linalg.norm(all_x[i, s] - c) - r) <= 1e-6
        # closeR: ascending, stable, all below the radius
        assert np.all(np.diff(d2[i, :count[i]]) >= 0)
        assert sorted(shape[i, :count[i]].tolist()) == [0, 1, 2, 3]


def test_closeR_cut_and_known_answers(orc):
    O = orc.Obstacles2D(("compound", [("circle", (0.5, 0.5), 0.1), fx.box2d([0.7, 0.9], [0.1, 0.3])]))
    P = np.array([[0.9, 0.5], [0.5, 0.5 + 0.35]])
    Ws = np.stack([np.eye(2), np.diag([1.0, 4.0])])
    count, d2, shape, x = orc.close_points(O, P, Ws, 0.1)
    # point 0: circle at Euclidean distance 0.3 (d2 = 0.09 < 0.1); box corner (0.9, 0.3) at d2 = 0.04 -> box first
    assert count[0] == 2 and shape[0, 0] == 1 and shape[0, 1] == 0
    assert abs(d2[0, 0] - 0.04) < 1e-15 and np.allclose(x[0, 0], [0.9, 0.3])
    assert abs(d2[0, 1] - 0.09) < 1e-9 and np.allclose(x[0, 1], [0.6, 0.5], atol=1e-6)
    # point 1: W = diag(1, 4): the circle's top point is 0.25 away in y -> d2 = 4 * 0.0625 = 0.25 > 0.1: cut
    assert count[1] == 0


@pytest.mark.parametrize("d", [2, 3, 4])
def test_box_closest_against_projected_gradient(orc, d):
    rng = np.random.Generator(np.random.PCG64(5 + d))
    M = 6
    lo = rng.random((M, d)) * 0.6
    hi = lo + 0.05 + rng.random((M, d)) * 0.3
    B = orc.Boxes([(lo[k], hi[k]) for k in range(M)])
    n = 40
    P = rng.random((n, d)) * 1.4 - 0.2
    Ws = _Ws(rng, n, d)
    count, d2, shape, x, all_d2, all_x = orc.close_points(B, P, Ws, 1e9, want_all=True)
    for i in range(n):
        W = Ws[i]
        step = 1.0 / np.linalg.eigvalsh(W).max()
        for k in range(M):
            v = np.clip(P[i], lo[k], hi[k])
            for _ in range(4000):                      # projected gradient on the convex QP
                v = np.clip(v - step * (W @ (v - P[i])), lo[k], hi[k])
            ref = (v - P[i]) @ W @ (v - P[i])
            assert abs(all_d2[i, k] - ref) <= 1e-7 * max(ref, 1e-9) + 1e-12
            assert np.all(all_x[i, k] >= lo[k]) and np.all(all_x[i, k] <= hi[k])
            if np.all((P[i] >= lo[k]) & (P[i] <= hi[k])):
                assert all_d2[i, k] == 0.0
        assert np.all(np.diff(d2[i, :count[i]]) >= 0) and count[i] == M


def test_matches_the_host_version_used_so_far(orc, mp):
    """montecarlo.closest (numpy) and the oracle agree to rounding on polygons and boxes"""
    rng = np.random.Generator(np.random.PCG64(77))
    O = orc.Obstacles2D(fx.ISRR_2H)
    P = rng.random((30, 2))
    Ws = _Ws(rng, 30, 2)
    _, _, _, _, all_d2, all_x = orc.close_points(O, P, Ws, 1e9, want_all=True)
    shapes = mp.obstaclesets.ISRR_2H().parts
    for i in range(30):
        for s, shp in enumerate(shapes):
            hd2, hx = mp.montecarlo.closest(P[i], shp, Ws[i])
            assert abs(hd2 - all_d2[i, s]) <= 1e-10 * max(1.0, hd2) and np.allclose(hx, all_x[i, s], atol=1e-9)
