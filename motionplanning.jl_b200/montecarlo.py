"""Monte-Carlo (importance-sampled) trajectory collision probability -- host side.

The estimator is NOT part of the reference source (README.md:9-10 links the papers; only the
helper geometry closest/closeR exists, SAT2D.jl:208-285, boxesND.jl:61-86).  Its specification is
SURVEY.md section 11 / DESIGN.md; the rollouts run on the GPU (libmpb200, mc.cu).  This module
prepares the problem the kernel consumes: the discrete LQG closed loop (Riccati recursions,
O(T n^3) on the host), the linear maps from stacked noise to workspace deviation, and the
defensive mixture proposal built from closest obstacle points under the marginal covariance
(`closest` mirrors SAT2D.jl:208-258 for circles and polygons, boxesND.jl:61-70 for boxes).
"""
import ctypes
import math

import numpy as np

from . import _lib
from .shapes2d import Circle, Compound2D, Polygon


class MCProblem:
    """Arrays of mpb200_mc_problem (row-major)."""

    def __init__(self, F, G, Wz, wbar, alpha=(1.0,), mu=None, swept=False):
        self.F = np.ascontiguousarray(F, dtype=np.float64)
        self.G = np.ascontiguousarray(G, dtype=np.float64)
        self.Wz = np.ascontiguousarray(Wz, dtype=np.float64)
        self.wbar = np.ascontiguousarray(wbar, dtype=np.float64)
        self.alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        self.T, self.nz, self.q = self.F.shape[0], self.F.shape[1], self.G.shape[2]
        self.dw = self.Wz.shape[0]
        self.K = len(self.alpha) - 1
        self.mu = (np.ascontiguousarray(np.asarray(mu, dtype=np.float64).reshape(self.K, self.T * self.q))
                   if self.K else np.zeros((0, self.T * self.q)))
        self.swept = bool(swept)
        assert self.wbar.shape == (self.T + 1, self.dw) and self.G.shape[:2] == (self.T, self.nz)

    def desc(self):
        return _lib.McProblem(self.T, self.nz, self.q, self.dw, _lib.ptr(self.F), _lib.ptr(self.G), _lib.ptr(self.Wz),
                              _lib.ptr(self.wbar), self.K, _lib.ptr(self.alpha), _lib.ptr(self.mu), int(self.swept))

    def noise_to_workspace(self):
        """M_t (dw x T q): workspace deviation at step t+1 as a linear map of the stacked noise."""
        T, nz, q = self.T, self.nz, self.q
        Z = np.zeros((nz, T * q))
        out = []
        for t in range(T):
            Z = self.F[t] @ Z
            Z[:, t * q:(t + 1) * q] += self.G[t]
            out.append(self.Wz @ Z)
        return out


def collision_probability(problem, CC, n, seed=20240605, first=0, per_rollout=False):
    """Run rollouts [first, first+n) on the GPU.  Returns dict(S1, S2, S0, n, hits, p, se[, hit, w])."""
    res = _lib.McResult()
    hit = np.empty(n, dtype=np.uint8) if per_rollout else None
    w = np.empty(n) if per_rollout else None
    d = problem.desc()
    _lib.check(_lib.lib().mpb200_mc_collision_probability(ctypes.byref(d), CC.handle(), seed, first, n, ctypes.byref(res),
                                                          _lib.ptr(hit), _lib.ptr(w)))
    out = dict(S1=res.s1, S2=res.s2, S0=res.s0, n=res.n, hits=res.hits)
    out.update(estimate(out))
    if per_rollout:
        out["hit"], out["w"] = hit.astype(bool), w
    return out


def estimate(sums):
    """p = S1/n, se = sqrt((S2/n - p^2)/n)"""
    n = max(int(sums["n"]), 1)
    p = sums["S1"] / n
    var = max(sums["S2"] / n - p * p, 0.0)
    return dict(p=p, se=math.sqrt(var / n))


def combine(parts):
    """Sum shard results (fixed rank order) -- what the all-reduce of the multi-GPU path computes."""
    tot = dict(S1=0.0, S2=0.0, S0=0.0, n=0, hits=0)
    for r in parts:
        for k in tot:
            tot[k] += r[k]
    tot.update(estimate(tot))
    return tot


# ---- LQG closed loop --------------------------------------------------------------------------
def lqg_closed_loop(A, B, C, Qc, Rc, V, W, T, Wz=None, P0=None):
    """Discrete LQG around a nominal trajectory:  x' = A x + B u + v,  y = C x + w,
    u = -L_t xhat (finite-horizon LQR), xhat from the Kalman filter.  z = [dx; dxhat].
    Returns (F, G) with F: T x 2n x 2n, G: T x 2n x (n + ny)."""
    A, B, C = (np.asarray(M, dtype=np.float64) for M in (A, B, C))
    n, ny = A.shape[0], C.shape[0]
    S = np.asarray(Qc, dtype=np.float64).copy()
    L = [None] * T
    for t in reversed(range(T)):                       # Riccati recursion for the LQR gains
        L[t] = np.linalg.solve(Rc + B.T @ S @ B, B.T @ S @ A)
        S = Qc + A.T @ S @ (A - B @ L[t])
    P = np.zeros((n, n)) if P0 is None else np.asarray(P0, dtype=np.float64)
    Vh, Wh = np.linalg.cholesky(V), np.linalg.cholesky(W)
    F = np.zeros((T, 2 * n, 2 * n))
    G = np.zeros((T, 2 * n, n + ny))
    for t in range(T):
        Pm = A @ P @ A.T + V                           # Kalman predict / update for step t+1
        Kt = Pm @ C.T @ np.linalg.inv(C @ Pm @ C.T + W)
        P = (np.eye(n) - Kt @ C) @ Pm
        ABL = A - B @ L[t]
        F[t] = np.block([[A, -B @ L[t]], [Kt @ C @ A, ABL - Kt @ C @ A]])
        G[t] = np.block([[Vh, np.zeros((n, ny))], [Kt @ C @ Vh, Kt @ Wh]])
    return F, G


# ---- closest obstacle points under a weight matrix (host geometry) -----------------------------------
def _closest_polypts(p, pts):
    """SAT2D.jl:240-252"""
    best, vbest = math.inf, pts[0]
    n = len(pts)
    for i in range(n):
        a, b = np.asarray(pts[i]), np.asarray(pts[(i + 1) % n])
        e = b - a
        x = float(e @ (p - a)) / float(e @ e)
        v = a if x < 0 else (a + x * e if x < 1 else b)
        d2 = float((p - v) @ (p - v))
        if d2 < best:
            best, vbest = d2, v
    return best, vbest


def closest(p, shape, W):
    """(squared W-distance, closest point) -- SAT2D.jl:212-258 (circle: Newton on the multiplier;
    polygon: Cholesky transform), boxesND.jl:61-70 (box: bounded least squares, here by enumeration
    of active sets, exact for the small dimensions used)."""
    p = np.asarray(p, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    if isinstance(shape, Polygon):
        Lc = np.linalg.cholesky(W).T                 # chol(W) upper, W = L'L
        _, v = _closest_polypts(Lc @ p, [Lc @ np.asarray(q) for q in shape.points])
        x = np.linalg.solve(Lc, v)
        return float((x - p) @ W @ (x - p)), x
    if isinstance(shape, Circle):
        c = np.asarray(shape.c)
        s, Vv = np.linalg.eigh(W)
        ctop = p - c
        p1, p2 = float(Vv[:, 0] @ ctop), float(Vv[:, 1] @ ctop)
        f = lambda lam: (p1 * s[0] / (lam + s[0])) ** 2 + (p2 * s[1] / (lam + s[1])) ** 2 - shape.r ** 2
        lam = 1.0
        fv = f(lam)
        it = 0
        while abs(fv) > 1e-12 and it < 200:
            fp = -2 / (lam + s[0]) * (p1 * s[0] / (lam + s[0])) ** 2 - 2 / (lam + s[1]) * (p2 * s[1] / (lam + s[1])) ** 2
            al = 1.0
            while True:
                ln = lam - al * fv / fp
                fn = f(ln)
                if abs(fn) < abs(fv) or al < 1e-12:
                    break
                al /= 2
            lam, fv = ln, fn
            it += 1
        x = c + Vv[:, 0] * p1 * s[0] / (lam + s[0]) + Vv[:, 1] * p2 * s[1] / (lam + s[1])
        return float((x - p) @ W @ (x - p)), x
    lo, hi = shape                                     # (lo, hi) box
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    d = len(lo)
    best, xbest = math.inf, np.clip(p, lo, hi)
    for code in range(3 ** d):                         # each coordinate: free / at lo / at hi
        st = [(code // 3 ** i) % 3 for i in range(d)]
        free = [i for i in range(d) if st[i] == 0]
        x = np.where(np.array(st) == 1, lo, hi).astype(np.float64)
        if free:
            fixed = [i for i in range(d) if st[i] != 0]
            rhs = -W[np.ix_(free, fixed)] @ (x[fixed] - p[fixed]) if fixed else np.zeros(len(free))
            x[free] = p[free] + np.linalg.solve(W[np.ix_(free, free)], rhs)
        if np.all(x >= lo - 1e-12) and np.all(x <= hi + 1e-12):
            d2 = float((x - p) @ W @ (x - p))
            if d2 < best:
                best, xbest = d2, np.clip(x, lo, hi)
    return best, xbest


def closeR(p, CC, W, r2):
    """closeR(p, CC, W, r2): closest points of every basic obstacle within squared W-distance r2,
    sorted ascending (SAT2D.jl:280-285, boxesND.jl:83-86)."""
    shapes = []
    if hasattr(CC, "boxes"):
        shapes = [(b.lo, b.hi) for b in CC.boxes]
    else:
        def walk(s):
            if isinstance(s, Compound2D):
                for q in s.parts:
                    walk(q)
            else:
                shapes.append(s)
        walk(CC.obstacles)
    cps = [closest(p, s, W) for s in shapes]
    return sorted([c for c in cps if c[0] < r2], key=lambda c: c[0])


def close_points(P, CC, Ws, r2, want_all=False):
    """closeR(p_i, CC, W_i, r2) for a whole batch on the GPU (mpb200_close_points, csrc/closest.cu): P is n x dw,
    Ws n x dw x dw (one SPD weight matrix per point).  Returns (count[n], d2[n,S], shape[n,S], x[n,S,dw]) -- per point
    the basic shapes closer than r2 in ascending squared W-distance -- and, with want_all, closest() for every
    basic shape as (all_d2[n,S], all_x[n,S,dw])."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    Ws = np.ascontiguousarray(Ws, dtype=np.float64)
    n, dw = P.shape
    S = max(len(CC.boxes) if hasattr(CC, "boxes") else CC.n_basic_shapes(), 1)
    count = np.zeros(n, dtype=np.int32)
    d2 = np.zeros((n, S))
    shape = np.full((n, S), -1, dtype=np.int32)
    x = np.zeros((n, S, dw))
    all_d2 = np.zeros((n, S)) if want_all else None
    all_x = np.zeros((n, S, dw)) if want_all else None
    _lib.check(_lib.lib().mpb200_close_points(CC.handle(), _lib.ptr(P), _lib.ptr(Ws), n, dw, float(r2), _lib.ptr(count),
                                              _lib.ptr(d2), _lib.ptr(shape), _lib.ptr(x), _lib.ptr(all_d2),
                                              _lib.ptr(all_x)))
    return (count, d2, shape, x) + ((all_d2, all_x) if want_all else ())


def with_proposal(problem, CC, r2=16.0, alpha0=0.2, max_components=16, device=True):
    """Defensive mixture: one component per (step, close obstacle point), shifted by the minimum-
    norm stacked noise whose mean trajectory touches that point; alpha_k ~ Phi(-sqrt(d2_k)).
    The close points of ALL steps come from one batched GPU call (device=True); device=False walks the steps
    with the host geometry above, as the reference's closeR would."""
    Ms = problem.noise_to_workspace()
    comps = []
    steps, Wts = [], []
    for t, M in enumerate(Ms):
        Sigma = M @ M.T
        if np.linalg.matrix_rank(Sigma) < Sigma.shape[0]:
            continue
        steps.append(t)
        Wts.append(np.linalg.inv(Sigma))
    if device and steps:
        Wts = [0.5 * (W + W.T) for W in Wts]                       # exactly symmetric, as the kernel reads W[0], W[1], W[3]
        P = np.stack([problem.wbar[t + 1] for t in steps])
        count, d2s, _, xs = close_points(P, CC, np.stack(Wts), r2)
        for i, t in enumerate(steps):
            for k in range(int(count[i])):
                mu = Ms[t].T @ Wts[i] @ (xs[i, k] - problem.wbar[t + 1])
                comps.append((float(d2s[i, k]), mu))
    else:
        for t, Wt in zip(steps, Wts):
            M = Ms[t]
            for d2, x in closeR(problem.wbar[t + 1], CC, Wt, r2):
                mu = M.T @ Wt @ (x - problem.wbar[t + 1])
                comps.append((d2, mu))
    comps.sort(key=lambda c: c[0])
    comps = comps[:max_components]
    if not comps:
        return MCProblem(problem.F, problem.G, problem.Wz, problem.wbar, (1.0,), None, problem.swept)
    wts = np.array([0.5 * math.erfc(math.sqrt(max(d2, 0.0)) / math.sqrt(2)) + 1e-12 for d2, _ in comps])
    alpha = np.concatenate([[alpha0], (1 - alpha0) * wts / wts.sum()])
    return MCProblem(problem.F, problem.G, problem.Wz, problem.wbar, alpha, np.stack([m for _, m in comps]),
                     problem.swept)
