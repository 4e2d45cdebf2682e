"""Chopped-metric car spaces: Reeds-Shepp (metric) and Dubins (quasimetric) over SE2 states (x, y, theta).

Mirror of src/statespaces/simplecars.jl: ReedsSheppExact / DubinsExact (:5-27), ReedsSheppMetricSpace /
DubinsQuasiMetricSpace (:29-40), steering_control / propagate / collision_waypoints (:55-82), and of the chopped,
lower-bounded metrics of src/primitivetypes.jl:74-107.  Steering, the neighbour tables and the waypoint collision
checks run on the GPU (csrc/cars.cu) through mpb200_car_*; this file carries the host-side description.
"""
import ctypes
import math

import numpy as np

from . import _lib
from .statespaces import BoundedStateSpace, Euclidean, VectorView

REEDS_SHEPP, DUBINS = 0, 1


class ReedsSheppExact:
    """simplecars.jl:5-12 -- metric: length of the shortest forward/backward path with turning radius r"""
    symmetric = True
    kind = REEDS_SHEPP

    def __init__(self, r=1.0, s=1.0):
        self.r, self.s = float(r), float(s)


class DubinsExact:
    """simplecars.jl:13-20 -- quasimetric: length of the shortest forward-only path with turning radius r"""
    symmetric = False
    kind = DUBINS

    def __init__(self, r=1.0, s=1.0):
        self.r, self.s = float(r), float(s)


class ChoppedMetric:
    """primitivetypes.jl:79-83 -- m(v, w) <= chopval ? m(v, w) : Inf, with lowerbound(v, w) <= m(v, w)"""
    symmetric = True

    def __init__(self, m, lowerbound=None, chopval=math.inf):
        self.m = m
        self.lowerbound = lowerbound if lowerbound is not None else Euclidean()
        self.chopval = float(chopval)


class ChoppedQuasiMetric(ChoppedMetric):
    """primitivetypes.jl:88-92"""
    symmetric = False


def ReedsSheppMetricSpace(r, s=1.0, lo=(0.0, 0.0), hi=(1.0, 1.0)):
    """simplecars.jl:29-34"""
    return BoundedStateSpace(np.concatenate([lo, [0.0]]), np.concatenate([hi, [2 * math.pi]]),
                             ChoppedMetric(ReedsSheppExact(r, s), Euclidean(), math.inf), VectorView(1, 2))


def DubinsQuasiMetricSpace(r, s=1.0, lo=(0.0, 0.0), hi=(1.0, 1.0)):
    """simplecars.jl:35-40"""
    return BoundedStateSpace(np.concatenate([lo, [0.0]]), np.concatenate([hi, [2 * math.pi]]),
                             ChoppedQuasiMetric(DubinsExact(r, s), Euclidean(), math.inf), VectorView(1, 2))


def is_car_metric(d):
    return isinstance(d, ChoppedMetric) and isinstance(d.m, (ReedsSheppExact, DubinsExact))


def _car(d):
    """the SimpleCarMetric behind a (possibly chopped) distance; statespaces.jl:143-146 forwards through .m"""
    d = getattr(d, "dist", d)
    d = getattr(d, "m", d)
    if not isinstance(d, (ReedsSheppExact, DubinsExact)):
        raise TypeError("not a simple-car metric")
    return d


def _pairs(v, w):
    V = np.ascontiguousarray(v, dtype=np.float64)
    W = np.ascontiguousarray(w, dtype=np.float64)
    if V.ndim == 1:
        V, W = V.reshape(1, -1), W.reshape(1, -1)
    if V.shape[1] != 3 or W.shape != V.shape:
        raise ValueError("SE2 states are (x, y, theta) rows")
    return V, W


def car_steer_batch(d, V, W):
    """reedsshepp / dubins for row-paired states -> (cost[n], nseg[n], segments[n, 5, 3]); a segment is a
    StepControl (duration, (signed speed, signed curvature)) (simplecars.jl:85-87)"""
    m = _car(d)
    V, W = _pairs(V, W)
    n = V.shape[0]
    cost = np.empty(n)
    nseg = np.empty(n, dtype=np.int32)
    segs = np.empty((n, 5, 3))
    _lib.check(_lib.lib().mpb200_car_steer(m.kind, m.r, m.s, _lib.ptr(V), _lib.ptr(W), n, _lib.ptr(cost), _lib.ptr(nseg),
                                           _lib.ptr(segs)))
    return cost, nseg, segs


def steering_control(d, v, w):
    """simplecars.jl:69-70 -> list of (t, (u1, u2))"""
    _, nseg, segs = car_steer_batch(d, v, w)
    return [(float(t), (float(u1), float(u2))) for t, u1, u2 in segs[0, :nseg[0]]]


def evaluate(d, v, w):
    """simplecars.jl:23-24 for the exact metrics; primitivetypes.jl:95-100 when chopped"""
    cost = float(car_steer_batch(d, v, w)[0][0])
    d = getattr(d, "dist", d)
    if isinstance(d, ChoppedMetric):
        v, w = np.asarray(v, dtype=np.float64), np.asarray(w, dtype=np.float64)
        if math.hypot(v[0] - w[0], v[1] - w[1]) > d.chopval or not cost <= d.chopval:
            return math.inf
    return cost


def propagate(v, u):
    """simplecars.jl:55-68 (host convenience for drawing paths; the collision checks use the device code)"""
    t, (s, invr) = u
    x, y, th = (float(a) for a in v)
    if abs(t * s * invr) > 10 * np.finfo(np.float64).eps:
        return np.array([x + (math.sin(th + t * s * invr) - math.sin(th)) / invr,
                         y + (math.cos(th) - math.cos(th + t * s * invr)) / invr,
                         math.fmod(math.fmod(th + t * s * invr, 2 * math.pi) + 2 * math.pi, 2 * math.pi)])
    return np.array([x + t * s * math.cos(th), y + t * s * math.sin(th),
                     math.fmod(math.fmod(th + t * s * invr, 2 * math.pi) + 2 * math.pi, 2 * math.pi)])


def car_motions_free(V, W, CC, SS):
    """Batch of is_free_motion(v, w, CC, SS) for a car space (statespaces.jl:153-158 over the arc waypoints of
    simplecars.jl:71-82); bumps CC.count by the segment tests the reference would have run."""
    m = _car(SS)
    V, W = _pairs(V, W)
    out = np.empty(V.shape[0], dtype=np.uint8)
    checks = _lib.c_i64(0)
    desc = SS.desc()
    _lib.check(_lib.lib().mpb200_car_motions_free(m.kind, m.r, m.s, _lib.ptr(V), _lib.ptr(W), V.shape[0], CC.handle(),
                                                  ctypes.byref(desc), _lib.ptr(out), ctypes.byref(checks)))
    CC.count += checks.value
    return out.astype(bool)
