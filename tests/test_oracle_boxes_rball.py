"""Known-answer tests for the oracle's N-d box checker (boxesND.jl:42-56) and r-ball query
(nearneighbors.jl:138-150,179-183)."""
import numpy as np

import fixtures as fx


def _slab_hits(v, w, lo, hi):
    t0, t1 = 0.0, 1.0
    for i in range(len(v)):
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


def test_box_point_test_is_correct_unlike_polygons(orc):
    B = orc.Boxes(fx.BOXES2D)
    assert not B.points_free([[0.08, 0.43]])[0]      # inside box 1 (the 2-D SAT checker says free: Q1)
    assert not B.points_free([[0.16, 0.5]])[0]       # closed box: corner collides
    assert B.points_free([[0.1, 0.1]])[0]
    P = fx.uniform_samples(20000, 2, 1)
    F = orc.Obstacles2D(fx.ISRR_2H, fixed_point_test=True)
    assert (B.points_free(P) == F.points_free(P)).all()


def test_box_segments_match_slab_test_when_start_is_outside_q2(orc):
    for boxes, d in ((fx.BOXES2D, 2), (fx.BOXES3D, 3)):
        B = orc.Boxes(boxes)
        P = fx.uniform_samples(60000, d, 21)
        V, W = P[:30000], P[30000:]
        W = V + (W - 0.5) * 0.4
        free_v = B.points_free(V)
        got = B.segments_free(V, W)
        los = [b[:, 0] for b in boxes]
        his = [b[:, 1] for b in boxes]
        exp = np.array([not any(_slab_hits(v, w, lo, hi) for lo, hi in zip(los, his)) for v, w in zip(V, W)])
        assert (got[free_v] == exp[free_v]).all()           # exact whenever v is outside every box
        # Q2: starting inside a box and leaving through a `lo` face can be missed
        missed = (~free_v) & got & (~exp)
        assert missed.any()
        assert not ((~free_v) & (~got) & exp).any()          # never a false collision


def test_box_segment_degenerate_direction_ieee(orc):
    B = orc.Boxes([(np.array([0.4, 0.4]), np.array([0.6, 0.6]))])
    # axis-parallel segments: v_to_w has a zero -> +-Inf/NaN lambdas, comparisons false (Q2)
    assert not B.segments_free([[0.1, 0.5]], [[0.9, 0.5]])[0]    # horizontal through the box
    assert B.segments_free([[0.1, 0.7]], [[0.9, 0.7]])[0]        # horizontal above it
    assert B.segments_free([[0.1, 0.5]], [[0.3, 0.5]])[0]        # stops short (broadphase)
    assert B.segments_free([[0.5, 0.5]], [[0.5, 0.5]])[0] in (True, False)   # zero-length inside: defined, no crash


def test_state_space_wrappers(orc):
    B = orc.Boxes(fx.BOXES2D)
    S = orc.StateSpace([0, 0], [1, 1])
    assert not orc.states_free(B, S, [[1.1, 0.1]])[0]                  # out of bounds
    ok, cnt = orc.motions_free_straight(B, S, [[1.1, 0.1], [0.1, 0.1], [0.1, 0.1]],
                                        [[0.9, 0.1], [1.2, 0.1], [0.9, 0.1]])
    # first waypoint out of bounds -> false without a segment check; the LAST waypoint is never
    # bounds-checked (statespaces.jl:155-157, Q4)
    assert ok.tolist() == [False, True, True] and cnt == 2
    # OutputMatrix / VectorView: 4-d double-integrator state, workspace = position
    C = np.hstack([np.eye(2), np.zeros((2, 2))])
    for s2w in (("matrix", C), ("view", [1, 2])):
        S4 = orc.StateSpace([0, 0, -1.5, -1.5], [1, 1, 1.5, 1.5], s2w)
        got = orc.states_free(B, S4, [[0.08, 0.43, 0, 0], [0.1, 0.1, 0, 0], [0.1, 0.1, 2.0, 0]])
        assert got.tolist() == [False, True, False]


def test_rball_hand_case(orc):
    V = np.array([[0.0, 0.0], [0.3, 0.0], [0.3, 0.4], [1.0, 1.0]])
    colptr, rowval, nzval = orc.rball_brute(V, 0.5)
    assert colptr.tolist() == [1, 3, 5, 7, 7]
    assert rowval.tolist() == [2, 3, 1, 3, 1, 2]
    # |V1-V3| = sqrt(0.09+0.16) = 0.5 exactly at the radius: s <= r*r keeps it
    assert nzval[1] == np.sqrt(0.3 * 0.3 + 0.4 * 0.4)
    assert nzval.tolist() == [0.3, nzval[1], 0.3, 0.4, nzval[1], 0.4]


def test_rball_kdtree_equals_brute(orc):
    for d, N, r in ((2, 3000, 0.05), (3, 2000, 0.12), (5, 1500, 0.4)):
        V = fx.uniform_samples(N, d, 100 + d)
        a = orc.rball_brute(V, r)
        b = orc.KDTree(V).rball(r)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        # size-independent properties: symmetric relation, ascending rows, no self loops
        colptr, rowval, nzval = a
        cols = np.repeat(np.arange(1, N + 1), np.diff(colptr))
        assert not (rowval == cols).any()
        fwd = set(zip(cols.tolist(), rowval.tolist()))
        assert all((j, i) in fwd for i, j in fwd)
        for v in range(N):
            seg = rowval[colptr[v] - 1:colptr[v + 1] - 1]
            assert (np.diff(seg) > 0).all()
        assert (nzval <= r).all()


def test_rball_shard_range(orc):
    V = fx.uniform_samples(500, 2, 9)
    full = orc.rball_brute(V, 0.1)
    part = orc.rball_brute(V, 0.1, q0=100, q1=250)
    lo, hi = full[0][100] - 1, full[0][250] - 1
    assert np.array_equal(part[1], full[1][lo:hi])
    assert np.array_equal(part[0], full[0][100:251] - lo)


def test_brute_predicates_agree_off_the_band(orc):
    V = fx.uniform_samples(1500, 2, 4)
    a = orc.rball_brute(V, 0.07, pred=0)
    b = orc.rball_brute(V, 0.07, pred=1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
