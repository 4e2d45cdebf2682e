"""Linear-quadratic ("drift") state spaces: LinearQuadratic metric, DoubleIntegrator, steer.

Mirror of src/statespaces/linearquadratic.jl: LinearQuadratic (:28-39), DoubleIntegrator (:46-53),
steer / steer_pairwise (:191-225), collision_waypoints (:85-88).  The 2BVP is solved on the GPU
(libmpb200): the double-integrator family through its closed form (lq.cu), every other linear-affine
system with nilpotent A -- drift c, integrator chains, general B and R -- numerically (lq_general.cu);
as in the reference only nilpotent A is possible (:94-98).
"""
import ctypes

import numpy as np

from . import _lib
from .statespaces import BoundedStateSpace, OutputMatrix


class LinearQuadratic:
    """linearquadratic.jl:28-39 -- quasimetric  cost = min_t  t + int u'Ru ; cmax is the steering radius."""
    symmetric = False

    def __init__(self, A, B, c, R, cmax=1.0):
        self.A = np.asarray(A, dtype=np.float64)
        self.B = np.asarray(B, dtype=np.float64)
        self.c = np.asarray(c, dtype=np.float64)
        self.R = np.asarray(R, dtype=np.float64)
        self.cmax = float(cmax)
        n = self.A.shape[0]
        if np.abs(np.linalg.matrix_power(self.A, n)).max() > 0:   # linearquadratic.jl:96
            raise NotImplementedError("TODO: implement more cases than nilpotent A! (e.g. diagonalizable)")
        self._handle = None

    def handle(self):
        if self._handle is None:
            lib = _lib.lib()
            h = _lib.c_vp()
            A = np.asfortranarray(self.A).ravel(order="F").copy()
            B = np.asfortranarray(self.B).ravel(order="F").copy()
            R = np.asfortranarray(self.R).ravel(order="F").copy()
            _lib.check(lib.mpb200_lq_create(_lib.ptr(A), _lib.ptr(B), _lib.ptr(self.c), _lib.ptr(R), self.A.shape[0],
                                            self.B.shape[1], ctypes.byref(h)))
            self._handle = h
        return self._handle

    def close(self):
        if self._handle is not None:
            _lib.load().mpb200_lq_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def setup_steering(d, r):
    """linearquadratic.jl:34 / statespaces.jl:73-75"""
    if hasattr(d, "dist"):
        d = d.dist
    if isinstance(d, LinearQuadratic):
        d.cmax = float(r)
    elif hasattr(d, "chopval"):                            # setup_steering(d::ChoppedPreMetric, r), statespaces.jl:75
        d.chopval = float(r)


def LinearQuadraticQuasiMetricSpace(lo, hi, A, B, c, R, C):
    """linearquadratic.jl:42-45"""
    return BoundedStateSpace(lo, hi, LinearQuadratic(A, B, c, R), OutputMatrix(C))


def DoubleIntegrator(d, lo=None, hi=None, vmax=1.5, r=1.0):
    """linearquadratic.jl:46-53"""
    lo = np.zeros(d) if lo is None else np.asarray(lo, dtype=np.float64)
    hi = np.ones(d) if hi is None else np.asarray(hi, dtype=np.float64)
    A = np.block([[np.zeros((d, d)), np.eye(d)], [np.zeros((d, 2 * d))]])
    B = np.vstack([np.zeros((d, d)), np.eye(d)])
    c = np.zeros(2 * d)
    R = r * np.eye(d)
    C = np.hstack([np.eye(d), np.zeros((d, d))])
    return LinearQuadraticQuasiMetricSpace(np.concatenate([lo, -vmax * np.ones(d)]),
                                           np.concatenate([hi, vmax * np.ones(d)]), A, B, c, R, C)


def _pairs(v, w):
    V = np.ascontiguousarray(v, dtype=np.float64)
    W = np.ascontiguousarray(w, dtype=np.float64)
    if V.ndim == 1:
        V, W = V.reshape(1, -1), W.reshape(1, -1)
    return V, W


def steer_batch(d, V, W, r=None):
    """steer(L, v, w, r) for row-paired states -> (cost[n], topt[n])"""
    V, W = _pairs(V, W)
    r = d.cmax if r is None else float(r)
    cost = np.empty(V.shape[0])
    topt = np.empty(V.shape[0])
    _lib.check(_lib.lib().mpb200_lq_steer(d.handle(), _lib.ptr(V), _lib.ptr(W), V.shape[0], r, _lib.ptr(cost),
                                          _lib.ptr(topt)))
    return cost, topt


def steer(d, v, w, r=None):
    """linearquadratic.jl:35,191-195 -> (cost, t*)"""
    c, t = steer_batch(d, v, w, r)
    return float(c[0]), float(t[0])


def evaluate(d, v, w):
    """linearquadratic.jl:37"""
    return steer(d, v, w)[0]


def lq_motions_free(V, W, CC, SS, r=None):
    """Batch of is_free_motion(v, w, CC, SS) for a LinearQuadratic space; bumps CC.count."""
    V, W = _pairs(V, W)
    d = SS.dist
    r = d.cmax if r is None else float(r)
    out = np.empty(V.shape[0], dtype=np.uint8)
    checks = _lib.c_i64(0)
    desc = SS.desc()
    _lib.check(_lib.lib().mpb200_lq_motions_free(d.handle(), r, _lib.ptr(V), _lib.ptr(W), V.shape[0], CC.handle(),
                                                 ctypes.byref(desc), _lib.ptr(out), ctypes.byref(checks)))
    CC.count += checks.value
    return out.astype(bool)


def lq_is_free_motion(v, w, CC, SS):
    """statespaces.jl:153-158 with collision_waypoints of linearquadratic.jl:85-88"""
    return bool(lq_motions_free(v, w, CC, SS)[0])
