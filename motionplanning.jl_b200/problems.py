"""Goals, MPProblem, MPSolution and batched free-space sampling.

Mirror of src/goals.jl (RectangleGoal :10-14,92-93; BallGoal :17-22,96-104; PointGoal/StateGoal
:47,72; is_goal_pt / sample_goal), src/problems.jl:5-44 (MPSolution, MPProblem) and
src/sampling.jl:3-45 (sample_free_goal, sample_free!).  Where the reference rejection-samples one
state at a time through is_free_state (sampling.jl:23-37), sample_free draws a batch on the host
and filters it with one GPU call (mpb200_states_free) -- SURVEY 8(f) rank 1.  Julia's rand stream
cannot be reproduced, so samples come from NumPy PCG64(seed); they are inputs, not results.
"""
import numpy as np

from .nearneighbors import MetricNN, QuasiMetricNN
from .statespaces import Euclidean, state2workspace, states_free, volume


class RectangleGoal:
    """goals.jl:10-14"""

    def __init__(self, lo, hi):
        self.lo, self.hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)


class BallGoal:
    """goals.jl:17-22"""

    def __init__(self, center, radius):
        self.center, self.radius = np.asarray(center, dtype=np.float64), float(radius)


class PointGoal:
    """goals.jl:47 -- ConvexHullWorkspaceGoal of one point"""

    def __init__(self, pt):
        self.pt = np.asarray(pt, dtype=np.float64)


class StateGoal:
    """goals.jl:72 -- ConvexHullStateSpaceGoal of one state"""

    def __init__(self, s):
        self.s = np.asarray(s, dtype=np.float64)


def is_goal_pt(v, G, SS):
    """goals.jl:92,96,107-110,127-129"""
    if isinstance(G, StateGoal):
        return bool(np.all(np.asarray(v) == G.s))
    w = state2workspace(v, SS)
    if isinstance(G, RectangleGoal):
        return bool(np.all((G.lo <= w) & (w <= G.hi)))
    if isinstance(G, BallGoal):
        return bool(np.sqrt(np.sum((w - G.center) ** 2)) <= G.radius)
    return bool(np.all(w == G.pt))


def goal_mask(V, G, SS):
    """is_goal_pt for every row of V (used by the planner's termination test)."""
    V = np.asarray(V, dtype=np.float64)
    if isinstance(G, StateGoal):
        return np.all(V == G.s, axis=1)
    if SS.s2w.kind == 0:
        Wk = V
    elif SS.s2w.kind == 1:
        Wk = V[:, [i - 1 for i in SS.s2w.inds]]
    else:
        Wk = V @ SS.s2w.C.T
    if isinstance(G, RectangleGoal):
        return np.all((G.lo <= Wk) & (Wk <= G.hi), axis=1)
    if isinstance(G, BallGoal):
        return np.sqrt(np.sum((Wk - G.center) ** 2, axis=1)) <= G.radius
    return np.all(Wk == G.pt, axis=1)


def _workspace2state(w, SS, rng):
    """statespaces.jl:62-70: embed a workspace point in a (random) state"""
    if SS.s2w.kind == 0:
        return np.asarray(w, dtype=np.float64)
    v = SS.lo + rng.random(SS.dim) * (SS.hi - SS.lo)
    if SS.s2w.kind == 1:
        v[[i - 1 for i in SS.s2w.inds]] = w
        return v
    C = SS.s2w.C
    return v + np.linalg.lstsq(C, np.asarray(w) - C @ v, rcond=None)[0]


def sample_goal(G, SS, rng):
    """goals.jl:89-90,93,97-104,122-125"""
    if isinstance(G, StateGoal):
        return G.s.copy()
    if isinstance(G, RectangleGoal):
        w = G.lo + (G.hi - G.lo) * rng.random(len(G.lo))
    elif isinstance(G, BallGoal):
        while True:
            w = G.center + 2 * G.radius * (rng.random(len(G.center)) - 0.5)
            if np.sqrt(np.sum((w - G.center) ** 2)) <= G.radius:
                break
    else:
        w = G.pt.copy()
    return _workspace2state(w, SS, rng)


class MPSolution:
    """problems.jl:5-10"""

    def __init__(self, status, cost, elapsed, metadata):
        self.status, self.cost, self.elapsed, self.metadata = status, cost, elapsed, metadata


class MPProblem:
    """problems.jl:12-44 -- state space, init, goal, collision checker, sample set V."""

    def __init__(self, SS, init, goal, CC, V=None):
        self.SS = SS
        self.init = np.asarray(init, dtype=np.float64)
        self.goal = goal
        self.CC = CC
        self.V = V if V is not None else defaultNN(SS, self.init)
        self.status = "not yet solved"
        self.solution = None


def defaultNN(SS, init):
    """statespaces.jl:163-170"""
    init = np.asarray(init, dtype=np.float64)
    if isinstance(SS.dist, Euclidean) or getattr(SS.dist, "symmetric", False):   # Metric / ChoppedMetric{<:Metric}
        return MetricNN(init.reshape(1, -1), SS.dist, init)
    return QuasiMetricNN(init.reshape(1, -1), SS.dist, init)


def sample_free_goal(P, rng):
    """sampling.jl:3-9"""
    while True:
        v = sample_goal(P.goal, P.SS, rng)
        if states_free(v, P.CC, P.SS)[0]:
            return v


def sample_free(P, N, ensure_goal=True, ensure_goal_ct=5, seed=0, batch=None, device=False):
    """sample_free!(P, N; ensure_goal_ct): sampling.jl:11-45.  Sample 1 is the init state, the last
    ensure_goal_ct samples are goal samples; the rest are uniform states that pass is_free_state,
    accepted in draw order.  Returns volume(SS) like the reference (:44).
    device=True draws the uniform bulk with mpb200_sample_free (Philox candidate stream, generated and
    filtered on the GPU) instead of the host generator + batched validity."""
    if N <= 0:
        return volume(P.SS)
    rng = np.random.Generator(np.random.PCG64(seed))
    SS = P.SS
    V = P.V.V
    have_init = len(V) > 0 and np.all(V[0] == P.init)
    W = np.empty((N, SS.dim))
    count = 0
    if not have_init:
        W[0] = P.init
        count = 1
    if device and N - count > 0:
        from .nearneighbors import MetricNN
        bulk = MetricNN.sample_free(P.CC, SS, N - count, seed=seed)
        W[count:] = bulk.V
        bulk.close()
        count = N
    batch = batch or max(1024, int(1.5 * N))
    while count < N:
        cand = SS.lo + rng.random((batch, SS.dim)) * (SS.hi - SS.lo)
        ok = states_free(cand, P.CC, SS)
        take = cand[ok][:N - count]
        W[count:count + len(take)] = take
        count += len(take)
    if ensure_goal:
        for i in range(1, min(ensure_goal_ct, N - 1) + 1):
            W[N - i] = sample_free_goal(P, rng)
    old = P.V
    P.V = type(old)(np.vstack([old.V, W]) if len(old.V) else W, old.dist, old.init)   # addpoints, nearneighbors.jl:108
    old.close()
    return volume(SS)
