"""Host-side 2-D shape constructors: Circle, Polygon, Box2D, Compound2D.

Mirrors src/collisioncheckers/SAT2D.jl:12-98 (constructors) and :189-204 (inflate).  These run
once per obstacle set on the host, exactly as in the reference; their output (points, unit
normals, per-normal extrema, AABBs) is what the device predicates consume -- the GPU never
re-derives a normal.  Arithmetic is plain IEEE double in the reference's operation order
(Python floats), so the tables are bit-identical to a straight restatement of the Julia code
(`normalize` = StaticArrays' inv(norm(v))*v).
"""
import math

import numpy as np


class Shape2D:
    pass


class Circle(Shape2D):
    """SAT2D.jl:12-25"""

    def __init__(self, c, r):
        r = float(r)
        if r <= 0:
            raise ValueError("Radius must be positive")  # SAT2D.jl:19
        self.c = (float(c[0]), float(c[1]))
        self.r = r
        self.xrange = (self.c[0] - r, self.c[0] + r)
        self.yrange = (self.c[1] - r, self.c[1] + r)

    def record(self):
        return [self.c[0], self.c[1], self.r, self.xrange[0], self.xrange[1], self.yrange[0], self.yrange[1]]


def _project_nextrema(points, n):
    """vec2Dutils.jl:18-27"""
    dmin, dmax = math.inf, -math.inf
    for p in points:
        d = p[0] * n[0] + p[1] * n[1]
        if d < dmin:
            dmin = d
        if d > dmax:
            dmax = d
    return (dmin, dmax)


class Polygon(Shape2D):
    """SAT2D.jl:29-51 -- convex polygon, reordered counter-clockwise, outward unit normals."""

    def __init__(self, points):
        pts = [(float(p[0]), float(p[1])) for p in points]
        n = len(pts)
        if n < 3:
            raise ValueError("Polygons need at least 3 points! Try Line?")  # SAT2D.jl:39
        s = 0.0
        for i in range(n):
            j = (i + 1) % n
            s += (pts[j][0] - pts[i][0]) * (pts[j][1] + pts[i][1])
        if s > 0:
            pts.reverse()  # SAT2D.jl:41
        edges = [(pts[(i + 1) % n][0] - pts[i][0], pts[(i + 1) % n][1] - pts[i][1]) for i in range(n)]
        normals = []
        for e in edges:
            p1, p2 = e[1], -e[0]  # perp, vec2Dutils.jl:6
            inv = 1.0 / math.sqrt(p1 * p1 + p2 * p2)
            normals.append((inv * p1, inv * p2))
        ang = [math.atan2(nn[1], nn[0]) for nn in normals]
        ang.append(ang[0])
        if any(-math.pi <= ang[i + 1] - ang[i] <= 0 for i in range(n)):
            raise ValueError("Polygon must be convex")  # SAT2D.jl:45
        self.points = pts
        self.edges = edges
        self.normals = normals
        self.xrange = (min(p[0] for p in pts), max(p[0] for p in pts))
        self.yrange = (min(p[1] for p in pts), max(p[1] for p in pts))
        self.nextrema = [_project_nextrema(pts, nn) for nn in normals]

    def record(self):
        out = [self.xrange[0], self.xrange[1], self.yrange[0], self.yrange[1]]
        for p in self.points:
            out += [p[0], p[1]]
        for nn in self.normals:
            out += [nn[0], nn[1]]
        for e in self.nextrema:
            out += [e[0], e[1]]
        return out


def Box2D(xr, yr):
    """SAT2D.jl:53-56"""
    return Polygon([(xr[0], yr[0]), (xr[1], yr[0]), (xr[1], yr[1]), (xr[0], yr[1])])


class Compound2D(Shape2D):
    """SAT2D.jl:82-96 -- parts plus the union AABB; an empty compound has AABB (0,0)-(0,0)."""

    def __init__(self, *parts):
        if len(parts) == 1 and isinstance(parts[0], (list, tuple)):
            parts = tuple(parts[0])
        self.parts = list(parts)
        if not self.parts:
            self.xrange = (0.0, 0.0)
            self.yrange = (0.0, 0.0)
        else:
            self.xrange = (min(p.xrange[0] for p in self.parts), max(p.xrange[1] for p in self.parts))
            self.yrange = (min(p.yrange[0] for p in self.parts), max(p.yrange[1] for p in self.parts))


def _cross(a, b):
    return a[0] * b[1] - a[1] * b[0]


def inflate(shape, eps, roundcorners=True):
    """SAT2D.jl:189-204 (host-side obstacle construction)."""
    eps = float(eps)
    if isinstance(shape, Circle):
        return Circle(shape.c, shape.r + eps)
    if isinstance(shape, Compound2D):
        return Compound2D([inflate(p, eps, roundcorners) for p in shape.parts])
    P = shape
    n = len(P.points)
    if not roundcorners:
        def push_out(n0, n1):
            cr = _cross(n0, n1)
            if abs(cr) < 1e-6:
                return n0
            return ((n1[1] - n0[1]) / cr, (-n1[0] + n0[0]) / cr)  # (perp(n1) - perp(n0)) / cross
        pts = []
        for i in range(n):
            v = push_out(P.normals[(i - 1) % n], P.normals[i])
            pts.append((P.points[i][0] + eps * v[0], P.points[i][1] + eps * v[1]))
        return Polygon(pts)
    pts = []
    for i in range(n):
        n0, n1 = P.normals[(i - 1) % n], P.normals[i]
        pts.append((P.points[i][0] + eps * n0[0], P.points[i][1] + eps * n0[1]))
        pts.append((P.points[i][0] + eps * n1[0], P.points[i][1] + eps * n1[1]))
    return Compound2D([Polygon(pts)] + [Circle(p, eps) for p in P.points])


def pack_obstacles(shape, fixed_point_test=False):
    """Flatten a shape tree into the arrays of mpb200_obstacles2d_desc (include/mpb200.h).

    Compound2D nodes become AABB gates (parent-before-child); basic shapes keep the index of
    their enclosing gate, so  gate-chain && shape-test  equals the reference's recursive
    `AABBseparated(C,S) && return false; @any [colliding(P,S) for P in C.parts]`.
    """
    gate_parent, gate_aabb = [], []
    kinds, gates, offs, data = [], [], [0], []

    def walk(s, parent):
        if isinstance(s, Compound2D):
            g = len(gate_parent)
            gate_parent.append(parent)
            gate_aabb.extend([s.xrange[0], s.xrange[1], s.yrange[0], s.yrange[1]])
            for p in s.parts:
                walk(p, g)
        elif isinstance(s, (Circle, Polygon)):
            kinds.append(0 if isinstance(s, Circle) else 1)
            gates.append(parent)
            data.extend(s.record())
            offs.append(len(data))
        else:
            raise TypeError("obstacles must be Circle, Polygon or Compound2D (got %r)" % type(s))

    walk(shape, -1)
    return {
        "n_gates": len(gate_parent),
        "gate_parent": np.asarray(gate_parent, dtype=np.int32),
        "gate_aabb": np.asarray(gate_aabb, dtype=np.float64),
        "n_shapes": len(kinds),
        "shape_kind": np.asarray(kinds, dtype=np.int32),
        "shape_gate": np.asarray(gates, dtype=np.int32),
        "shape_off": np.asarray(offs, dtype=np.int32),
        "data": np.asarray(data, dtype=np.float64),
        "flags": 1 if fixed_point_test else 0,
    }
