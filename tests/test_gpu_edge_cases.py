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
