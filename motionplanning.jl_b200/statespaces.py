"""State spaces and the (CollisionChecker, StateSpace) validity wrappers.

Mirror of src/statespaces.jl:29-60 (BoundedStateSpace, Identity/VectorView/OutputMatrix),
:150-160 (is_free_state / is_free_motion / is_free_path), src/statespaces/geometric.jl:8-20
(Euclidean spaces) and src/statespaces/linearquadratic.jl:41-53 (DoubleIntegrator).
The predicates themselves run on the GPU through libmpb200; this file only carries the
host-side description of a space across the C ABI (mpb200_space_desc).
"""
import ctypes
import math

import numpy as np

from . import _lib


# ---- State2Workspace (statespaces.jl:45-60) ------------------------------------------------
class Identity:
    kind = 0


class VectorView:
    kind = 1

    def __init__(self, *inds):
        if len(inds) == 1 and hasattr(inds[0], "__len__"):
            inds = tuple(inds[0])
        self.inds = tuple(int(i) for i in inds)  # 1-based, as in the reference


class OutputMatrix:
    kind = 2

    def __init__(self, C):
        self.C = np.asarray(C, dtype=np.float64)


# ---- metrics ---------------------------------------------------------------------------------
class Euclidean:
    """Distances.Euclidean as used by geometric.jl"""
    symmetric = True


class BoundedStateSpace:
    """statespaces.jl:29-34"""

    def __init__(self, lo, hi, dist, s2w):
        self.lo = np.ascontiguousarray(lo, dtype=np.float64)
        self.hi = np.ascontiguousarray(hi, dtype=np.float64)
        self.dist = dist
        self.s2w = s2w
        self._keep = None

    @property
    def dim(self):
        return len(self.lo)

    @property
    def workspace_dim(self):
        if self.s2w.kind == 0:
            return self.dim
        if self.s2w.kind == 1:
            return len(self.s2w.inds)
        return self.s2w.C.shape[0]

    def desc(self):
        """mpb200_space_desc for this space (arrays kept alive on self)."""
        inds = C = None
        dw = self.workspace_dim
        if self.s2w.kind == 1:
            inds = np.asarray([i - 1 for i in self.s2w.inds], dtype=np.int32)
        elif self.s2w.kind == 2:
            C = np.asfortranarray(self.s2w.C).ravel(order="F").copy()
        self._keep = (inds, C)
        return _lib.SpaceDesc(self.dim, _lib.ptr(self.lo), _lib.ptr(self.hi), self.s2w.kind, dw,
                              _lib.ptr(inds), _lib.ptr(C))


def volume(SS):
    """statespaces.jl:41"""
    return float(np.prod(SS.hi - SS.lo))


def dim(SS):
    return SS.dim


def state2workspace(v, SS):
    """statespaces.jl:57-60 (host convenience; the device applies the same map)."""
    v = np.asarray(v, dtype=np.float64)
    if SS.s2w.kind == 0:
        return v
    if SS.s2w.kind == 1:
        return v[[i - 1 for i in SS.s2w.inds]]
    return SS.s2w.C @ v


def BoundedEuclideanStateSpace(lo, hi):
    """geometric.jl:10-11"""
    return BoundedStateSpace(lo, hi, Euclidean(), Identity())


def UnitHypercube(d):
    """geometric.jl:12"""
    return BoundedEuclideanStateSpace(np.zeros(d), np.ones(d))


def _unbounded(d):
    return BoundedStateSpace(np.full(d, -math.inf), np.full(d, math.inf), Euclidean(), Identity())


# ---- validity wrappers (statespaces.jl:150-160) --------------------------------------------------
def _states(v):
    a = np.ascontiguousarray(v, dtype=np.float64)
    return a.reshape(1, -1) if a.ndim == 1 else a


def states_free(V, CC, SS=None):
    """Batch of is_free_state(v, CC, SS): V is n x d (row per state). Returns bool[n]."""
    V = _states(V)
    SS = SS if SS is not None else _unbounded(V.shape[1])
    out = np.empty(V.shape[0], dtype=np.uint8)
    d = SS.desc()
    _lib.check(_lib.lib().mpb200_states_free(_lib.ptr(V), V.shape[0], V.shape[1], CC.handle(), ctypes.byref(d),
                                             _lib.ptr(out)))
    return out.astype(bool)


def segments_free(V, W, CC, SS=None):
    """Batch of straight-edge is_free_motion(v, w, CC, SS); bumps CC.count like the reference."""
    V, W = _states(V), _states(W)
    SS = SS if SS is not None else _unbounded(V.shape[1])
    out = np.empty(V.shape[0], dtype=np.uint8)
    d = SS.desc()
    _lib.check(_lib.lib().mpb200_segments_free(_lib.ptr(V), _lib.ptr(W), V.shape[0], V.shape[1], CC.handle(),
                                               ctypes.byref(d), _lib.ptr(out)))
    # CC.count += 1 per segment test that actually runs: the wrapper short-circuits on
    # in_state_space(wps[1]) (statespaces.jl:155-157)
    CC.count += int(np.count_nonzero(np.all((SS.lo <= V) & (V <= SS.hi), axis=1)))
    return out.astype(bool)


def is_free_state(v, CC, SS=None):
    """statespaces.jl:151-152 / robots2D.jl:12 / boxesND.jl:25"""
    return bool(states_free(v, CC, SS)[0])


def is_free_motion(v, w, CC, SS=None):
    """statespaces.jl:153-158 for straight-line (Euclidean) waypoints / robots2D.jl:13 / boxesND.jl:26"""
    if SS is not None and not isinstance(SS.dist, Euclidean):
        from .simplecars import car_motions_free, is_car_metric
        if is_car_metric(SS.dist):
            return bool(car_motions_free(v, w, CC, SS)[0])
        from .linearquadratic import lq_is_free_motion
        return lq_is_free_motion(v, w, CC, SS)
    return bool(segments_free(v, w, CC, SS)[0])


def is_free_path(path, CC, SS=None):
    """statespaces.jl:159-160 / robots2D.jl:15-20 (one batch instead of a sequential loop)."""
    P = _states(path)
    if P.shape[0] < 2:
        return True
    if SS is not None and not isinstance(SS.dist, Euclidean):
        # is_free_motion(p[i], p[i+1], CC, SS) per pair: the waypoints of the optimal trajectory, not a chord
        from .simplecars import car_motions_free, is_car_metric
        if is_car_metric(SS.dist):
            return bool(np.all(car_motions_free(P[:-1], P[1:], CC, SS)))
        from .linearquadratic import lq_motions_free
        return bool(np.all(lq_motions_free(P[:-1], P[1:], CC, SS)))
    return bool(np.all(segments_free(P[:-1], P[1:], CC, SS)))
