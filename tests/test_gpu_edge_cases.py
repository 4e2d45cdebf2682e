"""GPU parity on degenerate inputs (the reference's own tests hold none; these are the cases its code paths
handle implicitly): single samples, duplicates, r = 0, empty ranges, empty obstacle sets, all-pairs radii."""
import numpy as np
import pytest

import fixtures as fx
from conftest import unpack_bits

pytestmark = pytest.mark.gpu


def _check_table(mp, orc, V, r):
    NN = mp.MetricNN(V)
    D = NN.precompute(r).D
    ec, er, ez = orc.rball_brute(V, r)
    assert np.array_equal(D.colptr, ec) and np.array_equal(D.rowval, er) and np.array_equal(D.nzval, ez)
    NN.close()
    return D


@pytest.mark.parametrize("d", [2, 3, 6])
def test_single_sample_and_pairs(gpu, orc, d):
    mp = gpu
    one = np.full((1, d), 0.25)
    D = _check_table(mp, orc, one, 0.3)
    assert D.nnz == 0 and list(D.colptr) == [1, 1]
    two = np.vstack([one, one])                       # exact duplicates: distance 0 <= r, stored with value 0
    D = _check_table(mp, orc, two, 0.0)
    assert D.nnz == 2 and np.all(D.nzval == 0.0)
    far = np.vstack([one, one + 0.5])
    D = _check_table(mp, orc, far, 0.1)
    assert D.nnz == 0


@pytest.mark.parametrize("d", [2, 3])
def test_all_pairs_radius_takes_the_big_column_path(gpu, orc, d):
    """r larger than the sample cloud: every column holds N-1 rows (beyond the 64-entry register path)."""
    mp = gpu
    V = fx.uniform_samples(700, d, 3) * 0.05
    D = _check_table(mp, orc, V, 1.0)
    assert D.nnz == 700 * 699


def test_collinear_and_lattice_points_at_exact_radius(gpu, orc):
    """ties: lattice points whose squared distance equals r*r exactly (representable) are neighbours (<=)"""
    mp = gpu
    g = np.arange(12, dtype=np.float64) * 0.125
    V = np.array([[x, y] for x in g for y in g])
    for r in (0.125, 0.25, 0.125 * np.sqrt(2.0), 0.375):
        _check_table(mp, orc, V, r)


def test_empty_query_range_and_empty_obstacles(gpu, orc):
    mp = gpu
    V = fx.uniform_samples(5000, 2, 11)
    NN = mp.MetricNN(V)
    NN.set_query_range(1234, 1234)                    # empty shard
    assert NN.build_table(0.05) == 0
    D = NN.fetch_table(NN.table)
    assert list(D.colptr) == [1] and D.nnz == 0
    NN.set_query_range(0, 5000)
    CC = mp.PointRobot2D(mp.Compound2D([]))            # no obstacles: everything inside the bounds is free
    SS = mp.UnitHypercube(2)
    F = unpack_bits(NN.points_free(CC, SS), 5000)
    assert F.all()
    nnz = NN.build_table(0.05)
    bits, checks = NN.edges_free(NN.table, CC, SS)
    assert unpack_bits(bits, nnz).all() and checks == nnz
    NN.close()


def test_everything_colliding(gpu, orc):
    """all samples and all edges inside one convex polygon: the classify pass's whole-column shortcut"""
    mp = gpu
    V = fx.uniform_samples(4000, 2, 12) * 0.5 + 0.25
    spec = ("compound", [("polygon", [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0)])])
    SSp = mp.UnitHypercube(2)
    SSo = orc.StateSpace([0, 0], [1, 1])
    for fixed in (False, True):
        CC = mp.PointRobot2D(fx.product_shape(mp, spec), fixed_point_test=fixed)
        O = orc.Obstacles2D(spec, fixed_point_test=fixed)
        NN = mp.MetricNN(V)
        D = NN.precompute(0.03).D
        F = unpack_bits(NN.points_free(CC, SSp), 4000)
        assert np.array_equal(F, orc.states_free(O, SSo, V))
        bits, checks = NN.edges_free(NN.table, CC, SSp)
        exp, cnt = orc.edges_free_csc(O, SSo, V, D.colptr, D.rowval)
        assert np.array_equal(unpack_bits(bits, D.nnz), exp.astype(bool)) and checks == cnt
        assert not exp.any()
        NN.close()


def test_lq_degenerate_sample_sets(gpu, orc):
    """ControlNN tables on 1 and 2 samples, duplicates included, and with a radius nothing reaches"""
    mp = gpu
    SS = mp.DoubleIntegrator(2)
    L = orc.DoubleIntegratorLQ(2)
    one = np.array([[0.2, 0.3, 0.1, -0.4]])
    cases = [(one, 0.7), (np.vstack([one, one]), 0.7), (np.vstack([one, one + [0.05, 0.0, 0.0, 0.1]]), 0.7),
             (np.vstack([one, one + [0.6, 0.6, 1.0, 1.0]]), 0.05)]
    for V, r in cases:
        NN = mp.QuasiMetricNN(V, SS.dist)
        cF, cB = NN.precompute(r)
        for cache, fwd in ((cF, True), (cB, False)):
            ref = L.inball(V, r, fwd)
            assert np.array_equal(cache.D.colptr, ref[0]) and np.array_equal(cache.D.rowval, ref[1])
            assert cache.D.nzval.tobytes() == ref[2].tobytes()
        NN.close()


def test_lq_many_survivors_per_column_keep_row_order(gpu, orc):
    """a tight cluster: almost every pair passes the candidate test, so the per-warp survivor ring is saturated
    and batches mix many owners -- rows must still come out ascending and bit-identical to the oracle"""
    mp = gpu
    rng = np.random.Generator(np.random.PCG64(31))
    V = np.hstack([0.5 + 0.02 * rng.random((600, 2)), 0.05 * (rng.random((600, 2)) - 0.5)])
    V[5] = V[4]
    SS = mp.DoubleIntegrator(2)
    L = orc.DoubleIntegratorLQ(2)
    NN = mp.QuasiMetricNN(V, SS.dist)
    cF, cB = NN.precompute(0.6)
    for cache, fwd in ((cF, True), (cB, False)):
        ref = L.inball(V, 0.6, fwd)
        assert np.array_equal(cache.D.colptr, ref[0]) and np.array_equal(cache.D.rowval, ref[1])
        assert cache.D.nzval.tobytes() == ref[2].tobytes()
        assert all(np.all(np.diff(cache.D.rowval[cache.D.colptr[c] - 1:cache.D.colptr[c + 1] - 1]) > 0) for c in range(600))
    assert cF.D.nnz > 600 * 300
    NN.close()


def test_mc_zero_and_one_rollout(gpu, orc):
    mp = gpu
    Bx = mp.PointRobotNDBoxes([mp.BoxBounds(np.array([0.5, -10.0]), np.array([10.0, 10.0]))])
    P = mp.MCProblem(np.eye(2)[None], (np.eye(2) * 0.1)[None], np.eye(2), np.array([[0.2, 0.0]] * 2), [0.3, 0.7],
                     [[3.0, 0.0]])
    r0 = mp.collision_probability(P, Bx, 0, seed=3)
    assert r0["n"] == 0 and r0["hits"] == 0 and r0["S1"] == 0.0
    r1 = mp.collision_probability(P, Bx, 1, seed=3, first=41)
    whole = mp.collision_probability(P, Bx, 64, seed=3, per_rollout=True)
    assert r1["n"] == 1
    one = mp.collision_probability(P, Bx, 1, seed=3, first=41, per_rollout=True)
    assert one["hit"][0] == whole["hit"][41] and one["w"][0] == whole["w"][41]


@pytest.mark.parametrize("case", ["sat2d", "sat2d_fixed", "boxes3d", "di_view"])
def test_device_sample_free_matches_the_oracle_stream(gpu, orc, case):
    """mpb200_sample_free: the device-generated sample set is bit-identical to oracle/sample.c --
    same candidates, same acceptance, same order -- whatever the chunking"""
    mp = gpu
    if case.startswith("sat2d"):
        fixed = case.endswith("fixed")
        CC, R = mp.PointRobot2D(mp.obstaclesets.ISRR_2H(), fixed_point_test=fixed), orc.Obstacles2D(fx.ISRR_2H, fixed_point_test=fixed)
        SSp, SSo = mp.UnitHypercube(2), orc.StateSpace([0, 0], [1, 1])
    elif case == "boxes3d":
        CC, R = mp.PointRobotNDBoxes([mp.BoxBounds(b) for b in fx.BOXES3D]), orc.Boxes(fx.BOXES3D)
        SSp, SSo = mp.UnitHypercube(3), orc.StateSpace([0, 0, 0], [1, 1, 1])
    else:  # 4-D double-integrator states, collision checked on the position coordinates only
        CC, R = mp.PointRobot2D(mp.obstaclesets.ISRR_2H(), fixed_point_test=True), orc.Obstacles2D(fx.ISRR_2H, fixed_point_test=True)
        lo, hi = [0, 0, -1.5, -1.5], [1, 1, 1.5, 1.5]
        SSp = mp.BoundedStateSpace(lo, hi, mp.Euclidean(), mp.VectorView(1, 2))
        SSo = orc.StateSpace(lo, hi, ("view", [1, 2]))
    for N, seed in ((1, 3), (777, 4), (60_000, 5)):
        NN = mp.MetricNN.sample_free(CC, SSp, N, seed=seed)
        V, used = orc.sample_free(R, SSo, N, seed)
        assert NN.candidates == used and NN.V.tobytes() == V.tobytes()
        NM = mp.MetricNN.sample_free(CC, SSp, N, seed=seed, order="morton")      # same set, Z-order numbering
        VM, _ = orc.sample_free(R, SSo, N, seed, order=1)
        assert NM.candidates == used and NM.V.tobytes() == VM.tobytes()
        NM.close()
        assert unpack_bits(NN.points_free(CC, SSp), N).all()            # the handle really holds those samples
        if N == 777 and SSp.dim <= 3:
            D = NN.precompute(0.08).D
            ec, er, ez = orc.rball_brute(V, 0.08)
            assert np.array_equal(D.colptr, ec) and np.array_equal(D.rowval, er) and np.array_equal(D.nzval, ez)
        NN.close()


def test_device_sample_free_with_no_free_space_fails_loudly(gpu):
    mp = gpu
    CC = mp.PointRobotNDBoxes([mp.BoxBounds(np.array([-1.0, -1.0]), np.array([2.0, 2.0]))])   # covers the whole square
    with pytest.raises(mp.MPB200Error):
        mp.MetricNN.sample_free(CC, mp.UnitHypercube(2), 100, seed=1)


@pytest.mark.parametrize("name", ["ISRR_2H", "ISRR_POLY", "TRI_BALLS"])
@pytest.mark.parametrize("fixed", [False, True])
def test_edge_validity_shortcuts_next_to_obstacle_boundaries(gpu, orc, name, fixed):
    """The classify pass settles whole columns from the r-box of the column point (clear of every cull box ->
    all free; inside one convex polygon by a margin -> all colliding).  Samples are placed ON and within
    1e-13 .. r of every obstacle AABB edge, polygon vertex and state-space bound, where those shortcuts must
    hand over to the exact per-edge test: bits and check counts have to stay identical to the oracle's."""
    mp = gpu
    spec = fx.ALL_2D[name]
    CC = mp.PointRobot2D(fx.product_shape(mp, spec), fixed_point_test=fixed)
    O = orc.Obstacles2D(spec, fixed_point_test=fixed)
    SSp, SSo = mp.UnitHypercube(2), orc.StateSpace([0, 0], [1, 1])
    rng = np.random.Generator(np.random.PCG64(17))
    xs, ys = [0.0, 1.0], [0.0, 1.0]
    for s in spec[1]:
        if s[0] == "polygon":
            P = np.array(s[1]); xs += list(P[:, 0]); ys += list(P[:, 1])
        else:
            (cx, cy), rad = s[1], s[2]; xs += [cx - rad, cx, cx + rad]; ys += [cy - rad, cy, cy + rad]
    for r in (1e-3, 0.02):
        offs = np.array([0.0, 1e-13, -1e-13, 1e-9, -1e-9, r, -r, r * (1 + 1e-9), -r * (1 + 1e-9), 0.5 * r, -0.5 * r])
        gx = np.unique(np.clip(np.add.outer(np.array(xs), offs).ravel(), 0.0, 1.0))
        gy = np.unique(np.clip(np.add.outer(np.array(ys), offs).ravel(), 0.0, 1.0))
        V = np.array([[x, y] for x in gx for y in gy])
        V = np.vstack([V, V[rng.integers(0, len(V), 4000)] + (rng.random((4000, 2)) - 0.5) * 2 * r])   # neighbours within r
        V = np.clip(V, 0.0, 1.0)
        NN = mp.MetricNN(V)
        D = NN.precompute(r).D
        F = unpack_bits(NN.points_free(CC, SSp), len(V))
        assert np.array_equal(F, orc.states_free(O, SSo, V))
        bits, checks = NN.edges_free(NN.table, CC, SSp)
        exp, cnt = orc.edges_free_csc(O, SSo, V, D.colptr, D.rowval)
        got = unpack_bits(bits, D.nnz)
        assert np.array_equal(got, exp.astype(bool)), "first mismatch at stored entry %d" % int(np.argmax(got != exp.astype(bool)))
        assert checks == cnt and D.nnz > 1000
        NN.close()


@pytest.mark.parametrize("d,boxes", [(2, "BOXES2D"), (3, "BOXES3D")])
def test_box_edge_validity_shortcuts_next_to_faces(gpu, orc, d, boxes):
    """same idea for the N-d box checker: samples on and next to every box face and state-space bound"""
    mp = gpu
    B = getattr(fx, boxes)
    CC, R = mp.PointRobotNDBoxes([mp.BoxBounds(b) for b in B]), orc.Boxes(B)
    SSp, SSo = mp.UnitHypercube(d), orc.StateSpace([0] * d, [1] * d)
    rng = np.random.Generator(np.random.PCG64(23))
    r = 0.01 if d == 2 else 0.03
    offs = np.array([0.0, 1e-13, -1e-13, r, -r, r * (1 + 1e-9), -r * (1 + 1e-9)])
    faces = [np.unique(np.clip(np.add.outer(np.concatenate([[0.0, 1.0]] + [np.asarray(b)[i] for b in B]), offs).ravel(), 0, 1))
             for i in range(d)]
    base = np.stack([f[rng.integers(0, len(f), 6000)] for f in faces], axis=1)      # every coordinate on/near a face
    V = np.clip(np.vstack([base, base[rng.integers(0, len(base), 6000)] + (rng.random((6000, d)) - 0.5) * 2 * r]), 0, 1)
    NN = mp.MetricNN(V)
    D = NN.precompute(r).D
    assert np.array_equal(unpack_bits(NN.points_free(CC, SSp), len(V)), orc.states_free(R, SSo, V))
    bits, checks = NN.edges_free(NN.table, CC, SSp)
    exp, cnt = orc.edges_free_csc(R, SSo, V, D.colptr, D.rowval)
    assert np.array_equal(unpack_bits(bits, D.nnz), exp.astype(bool)) and checks == cnt and D.nnz > 1000
    NN.close()


@pytest.mark.parametrize("N,d,r", [(3000, 2, 2.0), (60000, 2, 0.02), (1_200_000, 2, 0.0011), (40000, 3, 0.06)])
def test_table_fetch_bit_packed_indices_round_trip(gpu, N, d, r):
    """mpb200_table_fetch sends the row indices bit-packed (12 / 16 / 20 / 24 ... bits by sample count) and unpacks them
    on the host: the caller's Int64 array must equal the device array, entry for entry, at every width and for entry
    counts that are not a multiple of the packing group or the chunking"""
    mp = gpu
    from mpb200 import sharding
    V = fx.uniform_samples(N, d, 4242 + N)
    NN = mp.MetricNN(V)
    D = NN.precompute(r).D
    assert D.nnz >= 1 << 20                       # large enough to take the packed path
    cp, rv, nz, _ = sharding.table_device_tensors(NN.table)
    assert np.array_equal(D.rowval, rv.cpu().numpy())
    assert np.array_equal(D.colptr, cp.cpu().numpy())
    assert D.nzval.tobytes() == nz.cpu().numpy().tobytes()
    assert D.rowval.min() >= 1 and D.rowval.max() <= N
    NN.close()
