"""Collision checkers backed by libmpb200: PointRobot2D, PointRobotNDBoxes.

Mirror of the reference's SweptCollisionChecker plug-in interface
(src/collisioncheckers.jl:4-6; robots2D.jl:5-31; boxesND.jl:5-38): is_free_state,
is_free_motion, is_free_path, inflate, addobstacle, addblocker and the mutable `count`
field FMT* resets and reports (fmt.jl:12,106).  State-level calls launch a (tiny) batch
kernel; the batched entry points (`points_free`, `edges_free`) are what the planner uses.
"""
import ctypes

import numpy as np

from . import _lib
from . import shapes2d
from .shapes2d import Circle, Compound2D


class CollisionChecker:
    kind = None

    def __init__(self):
        self.count = 0
        self._handle = None

    # -- device handle -------------------------------------------------------------
    def handle(self):
        if self._handle is None:
            self._handle = self._create()
        return self._handle

    def _create(self):
        raise NotImplementedError

    def close(self):
        if self._handle is not None:
            _lib.load().mpb200_obstacles_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SweptCollisionChecker(CollisionChecker):
    pass


class PointRobot2D(SweptCollisionChecker):
    """robots2D.jl:5-10 -- point robot among a Shape2D (pretty much always a Compound2D)."""
    kind = 0
    workspace_dim = 2

    def __init__(self, obstacles, fixed_point_test=False):
        super().__init__()
        self.obstacles = obstacles
        self.fixed_point_test = bool(fixed_point_test)
        self.packed = shapes2d.pack_obstacles(obstacles, fixed_point_test)

    def n_basic_shapes(self):
        """circles + polygons of the shape tree (the S of mpb200_close_points)"""
        return int(self.packed["n_shapes"])

    def _create(self):
        lib = _lib.lib()
        p = self.packed
        d = _lib.ObstaclesDesc(
            p["n_gates"], _lib.ptr(p["gate_parent"]), _lib.ptr(p["gate_aabb"]),
            p["n_shapes"], _lib.ptr(p["shape_kind"]), _lib.ptr(p["shape_gate"]), _lib.ptr(p["shape_off"]),
            _lib.ptr(p["data"]), p["flags"])
        h = _lib.c_vp()
        _lib.check(lib.mpb200_obstacles2d_create(ctypes.byref(d), ctypes.byref(h)))
        return h


class BoxBounds:
    """boxesND.jl:5-13"""

    def __init__(self, lo, hi=None):
        if hi is None:  # BoxBounds(lohi::Matrix): column 1 = lo, column 2 = hi
            lohi = np.asarray(lo, dtype=np.float64)
            lo, hi = lohi[:, 0], lohi[:, 1]
        self.lo = np.ascontiguousarray(lo, dtype=np.float64)
        self.hi = np.ascontiguousarray(hi, dtype=np.float64)


class PointRobotNDBoxes(SweptCollisionChecker):
    """boxesND.jl:15-23 -- point robot among axis-aligned N-d boxes."""
    kind = 1

    def __init__(self, boxes):
        super().__init__()
        bl = []
        for b in boxes:
            bl.append(b if isinstance(b, BoxBounds) else BoxBounds(b))
        self.boxes = bl
        self.d = len(bl[0].lo) if bl else 0
        self.workspace_dim = self.d
        self.lo = np.ascontiguousarray(np.stack([b.lo for b in bl]) if bl else np.zeros((0, 0)), dtype=np.float64)
        self.hi = np.ascontiguousarray(np.stack([b.hi for b in bl]) if bl else np.zeros((0, 0)), dtype=np.float64)

    def _create(self):
        lib = _lib.lib()
        h = _lib.c_vp()
        _lib.check(lib.mpb200_boxes_create(_lib.ptr(self.lo), _lib.ptr(self.hi), len(self.boxes), max(self.d, 1),
                                           ctypes.byref(h)))
        return h


# ---- reference-named free functions ----------------------------------------------------
def inflate(CC, eps, roundcorners=True):
    """robots2D.jl:21-22 / boxesND.jl:30 (host-side construction)."""
    if eps <= 0:
        return CC
    if isinstance(CC, PointRobot2D):
        return PointRobot2D(shapes2d.inflate(CC.obstacles, eps, roundcorners), CC.fixed_point_test)
    return PointRobotNDBoxes([BoxBounds(b.lo - float(eps), b.hi + float(eps)) for b in CC.boxes])


def addobstacle(CC, o):
    """robots2D.jl:23 (nests a new Compound2D) / boxesND.jl:31"""
    if isinstance(CC, PointRobot2D):
        return PointRobot2D(Compound2D(CC.obstacles, o), CC.fixed_point_test)
    return PointRobotNDBoxes(CC.boxes + [o if isinstance(o, BoxBounds) else BoxBounds(o)])


def addblocker(CC, p, r):
    """robots2D.jl:24 / boxesND.jl:32"""
    if isinstance(CC, PointRobot2D):
        return addobstacle(CC, Circle(p, r))
    p = np.asarray(p, dtype=np.float64)
    return addobstacle(CC, BoxBounds(p - r, p + r))
