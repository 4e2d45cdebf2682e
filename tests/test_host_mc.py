"""Host-side pieces of the Monte-Carlo path (no GPU): LQG closed loop, noise maps, closest points."""
import numpy as np
import pytest

import fixtures as fx


def test_noise_to_workspace_matches_simulation(mp):
    rng = np.random.default_rng(0)
    dt, T = 0.1, 12
    A = np.block([[np.eye(2), dt * np.eye(2)], [np.zeros((2, 2)), np.eye(2)]])
    B = np.vstack([0.5 * dt * dt * np.eye(2), dt * np.eye(2)])
    C = np.hstack([np.eye(2), np.zeros((2, 2))])
    F, G = mp.montecarlo.lqg_closed_loop(A, B, C, np.eye(4), np.eye(2), 1e-3 * np.eye(4), 1e-3 * np.eye(2), T)
    assert F.shape == (T, 8, 8) and G.shape == (T, 8, 6)
    P = mp.MCProblem(F, G, np.hstack([np.eye(2), np.zeros((2, 6))]), np.zeros((T + 1, 2)))
    Ms = P.noise_to_workspace()
    eps = rng.standard_normal(T * 6)
    z = np.zeros(8)
    for t in range(T):
        z = F[t] @ z + G[t] @ eps[t * 6:(t + 1) * 6]
        assert np.allclose(P.Wz @ z, Ms[t] @ eps, atol=1e-12)
    # the closed loop is stable: deviations stay bounded
    assert np.abs(np.linalg.eigvals(F[0])).max() < 1.0 + 1e-9


def test_closest_points(mp):
    mc = mp.montecarlo
    W = np.array([[4.0, 0.5], [0.5, 1.0]])
    rng = np.random.default_rng(1)
    box = (np.array([0.4, 0.2]), np.array([0.6, 0.5]))
    poly = mp.Box2D([0.4, 0.6], [0.2, 0.5])
    for _ in range(50):
        p = rng.random(2)
        d2b, xb = mc.closest(p, box, W)
        d2p, xp = mc.closest(p, poly, W)
        # brute force over a fine boundary/interior grid
        gx, gy = np.meshgrid(np.linspace(0.4, 0.6, 201), np.linspace(0.2, 0.5, 301))
        D = np.stack([gx.ravel() - p[0], gy.ravel() - p[1]], axis=1)
        best = np.einsum("ij,jk,ik->i", D, W, D).min()
        assert abs(d2b - best) < 2e-3 * max(best, 1e-3) + 1e-9
        inside = (0.4 <= p[0] <= 0.6) and (0.2 <= p[1] <= 0.5)
        if not inside:
            assert abs(d2p - d2b) < 1e-9 and np.allclose(xb, xp, atol=1e-7)
    c = mp.Circle((0.5, 0.5), 0.1)
    d2, x = mc.closest(np.array([0.9, 0.5]), c, np.eye(2))
    assert abs(d2 - 0.09) < 1e-9 and np.allclose(x, [0.6, 0.5], atol=1e-6)


def test_with_proposal_builds_normalised_mixture(mp):
    CC = mp.PointRobotNDBoxes([mp.BoxBounds(b) for b in fx.BOXES2D])
    T = 5
    F = np.stack([np.eye(2)] * T)
    G = np.stack([np.eye(2) * 0.02] * T)
    wbar = np.stack([np.linspace(0.3, 0.6, T + 1), np.full(T + 1, 0.17)], axis=1)
    P = mp.montecarlo.with_proposal(mp.MCProblem(F, G, np.eye(2), wbar), CC, r2=25.0, device=False)   # host geometry (no GPU here)
    assert P.K >= 1 and abs(P.alpha.sum() - 1) < 1e-12 and P.alpha[0] == 0.2
    Ms = P.noise_to_workspace()
    # every shifted mean trajectory touches an obstacle boundary point at some step
    touched = 0
    for k in range(P.K):
        for t in range(T):
            w = wbar[t + 1] + Ms[t] @ P.mu[k]
            if any(abs(w[1] - 0.19) < 1e-6 and 0.4 - 1e-6 <= w[0] <= 0.5 + 1e-6 for _ in [0]):
                touched += 1
    assert touched >= 1


def test_save_load_nn_roundtrip(tmp_path):
    """saveNN / loadNN (names only in the reference, nearneighbors.jl:8): the CSC fields and the validity
    sidecars survive the round trip; inconsistent files are rejected"""
    import mpb200
    rng = np.random.Generator(np.random.PCG64(3))
    n = 50
    counts = rng.integers(0, 6, size=n)
    colptr = np.concatenate([[1], 1 + np.cumsum(counts)]).astype(np.int64)
    nnz = int(colptr[-1] - 1)
    rowval = np.concatenate([np.sort(rng.choice(n, c, replace=False)) + 1 for c in counts]).astype(np.int64) if nnz else np.zeros(0, np.int64)
    nzval = rng.random(nnz)
    cache = mpb200.ImmutableNNC(mpb200.SparseMatrixCSC(n, n, colptr, rowval, nzval), 0.25)
    edge = rng.integers(0, 2**63, size=(nnz + 63) // 64, dtype=np.uint64)
    pts = rng.integers(0, 2**63, size=1, dtype=np.uint64)
    path = str(tmp_path / "nn.npz")
    mpb200.saveNN(path, cache, edge, pts)
    c2, e2, p2 = mpb200.loadNN(path)
    assert c2.r == 0.25 and c2.D.m == n and c2.D.n == n
    assert np.array_equal(c2.D.colptr, colptr) and np.array_equal(c2.D.rowval, rowval)
    assert c2.D.nzval.tobytes() == nzval.tobytes() and np.array_equal(e2, edge) and np.array_equal(p2, pts)
    mpb200.saveNN(path, cache, edge[:-1] if len(edge) > 1 else np.zeros(len(edge) + 1, np.uint64))
    with pytest.raises(ValueError):
        mpb200.loadNN(path)
