"""GPU parity of the batched closest / closeR kernel (csrc/closest.cu) against its oracle (oracle/closest.c):
bit-exact squared W-distances, closest points, shape order and counts -- 2-D shape sets (circles + polygons) and
box lists in 2, 3 and 4 dimensions -- and the device-built Monte-Carlo proposal against the host-built one."""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu


def _Ws(rng, n, d):
    out = np.empty((n, d, d))
    for i in range(n):
        A = rng.standard_normal((d, d))
        W = A @ A.T + 0.3 * np.eye(d)
        out[i] = 0.5 * (W + W.T)
    return out


@pytest.mark.parametrize("name", ["ISRR_2H", "TRI_BALLS", "ISRR_POLY_WITH_SPIKE"])
def test_close_points_2d_matches_oracle(gpu, orc, name):
    mp = gpu
    spec = fx.ALL_2D[name]
    CC = mp.PointRobot2D(fx.product_shape(mp, spec))
    O = orc.Obstacles2D(spec)
    rng = np.random.Generator(np.random.PCG64(len(name)))
    n = 5000
    P = rng.random((n, 2)) * 1.2 - 0.1
    Ws = _Ws(rng, n, 2)
    Ws[:50] = np.eye(2)                                      # b == 0 branch of the 2x2 eigen-decomposition
    for r2 in (0.05, 1e9):
        got = mp.montecarlo.close_points(P, CC, Ws, r2, want_all=True)
        exp = orc.close_points(O, P, Ws, r2, want_all=True)
        assert np.array_equal(got[0], exp[0])
        for i in np.flatnonzero(exp[0] > 0)[::7]:
            k = exp[0][i]
            assert got[1][i, :k].tobytes() == exp[1][i, :k].tobytes()
            assert np.array_equal(got[2][i, :k], exp[2][i, :k]) and got[3][i, :k].tobytes() == exp[3][i, :k].tobytes()
        assert got[4].tobytes() == exp[4].tobytes() and got[5].tobytes() == exp[5].tobytes()
    assert 0 < (exp[0] > 0).mean()


@pytest.mark.parametrize("d", [2, 3, 4])
def test_close_points_boxes_match_oracle(gpu, orc, d):
    mp = gpu
    rng = np.random.Generator(np.random.PCG64(40 + d))
    M = 9
    lo = rng.random((M, d)) * 0.6
    hi = lo + 0.05 + rng.random((M, d)) * 0.3
    CC = mp.PointRobotNDBoxes([mp.BoxBounds(lo[k], hi[k]) for k in range(M)])
    B = orc.Boxes([(lo[k], hi[k]) for k in range(M)])
    n = 1500
    P = rng.random((n, d)) * 1.4 - 0.2
    P[:40] = 0.5 * (lo[0] + hi[0])                           # inside a box: distance 0
    Ws = _Ws(rng, n, d)
    got = mp.montecarlo.close_points(P, CC, Ws, 0.5, want_all=True)
    exp = orc.close_points(B, P, Ws, 0.5, want_all=True)
    assert np.array_equal(got[0], exp[0])
    assert got[4].tobytes() == exp[4].tobytes() and got[5].tobytes() == exp[5].tobytes()
    for i in range(0, n, 11):
        k = exp[0][i]
        assert got[1][i, :k].tobytes() == exp[1][i, :k].tobytes() and np.array_equal(got[2][i, :k], exp[2][i, :k])
    assert (got[4][:40, 0] == 0).all()


def test_device_built_proposal_equals_host_built(gpu):
    """with_proposal: the mixture built from ONE batched device call == the one built step by step on the host"""
    mp = gpu
    import bench_configs as bc
    P_dev, CC, naive = bc.c5_problem(mp)
    P_host = mp.montecarlo.with_proposal(naive, CC, r2=36.0, max_components=8, device=False)
    assert P_dev.K == P_host.K >= 1
    assert np.allclose(P_dev.alpha, P_host.alpha, rtol=1e-9, atol=1e-12)
    assert np.allclose(P_dev.mu, P_host.mu, rtol=1e-7, atol=1e-9)
