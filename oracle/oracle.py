"""ctypes front-end of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.  See mp_oracle.h for
the parity-pinning statement.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
c_i64, c_i32, c_dbl, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_double, ctypes.c_void_p
P = ctypes.POINTER
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        res = subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return _SO


class Obs2D(ctypes.Structure):
    _fields_ = [("n_gates", c_i32), ("gate_parent", c_vp), ("gate_aabb", c_vp), ("n_shapes", c_i32),
                ("shape_kind", c_vp), ("shape_gate", c_vp), ("shape_off", c_vp), ("data", c_vp), ("flags", c_i32)]


class Space(ctypes.Structure):
    _fields_ = [("n", c_i32), ("lo", c_vp), ("hi", c_vp), ("s2w_kind", c_i32), ("dw", c_i32), ("inds", c_vp),
                ("C", c_vp)]


class Checker(ctypes.Structure):
    _fields_ = [("kind", c_i32), ("obs2d", P(Obs2D)), ("box_lo", c_vp), ("box_hi", c_vp), ("M", c_i32), ("d", c_i32)]


class McProblem(ctypes.Structure):
    _fields_ = [("T", c_i32), ("nz", c_i32), ("q", c_i32), ("dw", c_i32), ("F", c_vp), ("G", c_vp), ("Wz", c_vp),
                ("wbar", c_vp), ("K", c_i32), ("alpha", c_vp), ("mu", c_vp), ("swept", c_i32)]


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        L = _lib
        L.orc_circle_build.argtypes = [c_dbl, c_dbl, c_dbl, c_vp]
        L.orc_polygon_build.argtypes = [c_vp, ctypes.c_int, c_vp]
        L.orc_point_colliding_2d.argtypes = [P(Obs2D), c_dbl, c_dbl]
        L.orc_line_colliding_2d.argtypes = [P(Obs2D), c_dbl, c_dbl, c_dbl, c_dbl]
        L.orc_points_free_2d.argtypes = [P(Obs2D), c_vp, c_i64, c_vp]
        L.orc_segments_free_2d.argtypes = [P(Obs2D), c_vp, c_vp, c_i64, c_vp]
        L.orc_points_free_boxes.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_i64, c_vp]
        L.orc_segments_free_boxes.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_i64, c_vp]
        L.orc_states_free.argtypes = [P(Checker), P(Space), c_vp, c_i64, c_vp]
        L.orc_edges_free_csc.argtypes = [P(Checker), P(Space), c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, P(c_i64)]
        L.orc_is_free_motion_straight.argtypes = [P(Checker), P(Space), c_vp, c_vp, P(c_i64)]
        L.orc_rball_brute.argtypes = [c_vp, c_i64, ctypes.c_int, c_dbl, ctypes.c_int, c_i64, c_i64, c_vp, c_vp, c_vp]
        L.orc_kdtree_build.restype = c_vp
        L.orc_kdtree_build.argtypes = [c_vp, c_i64, ctypes.c_int, ctypes.c_int]
        L.orc_kdtree_free.argtypes = [c_vp]
        L.orc_rball_kdtree.argtypes = [c_vp, c_dbl, c_i64, c_i64, c_vp, c_vp, c_vp]
        L.orc_philox4x32_10.argtypes = [c_vp, c_vp, c_vp]
        L.orc_sample_candidate.argtypes = [P(Space), ctypes.c_uint64, c_i64, c_vp]
        L.orc_sample_free.restype = c_i64
        L.orc_sample_free.argtypes = [P(Checker), P(Space), c_i64, ctypes.c_uint64, c_i64, ctypes.c_int, c_vp, P(c_i64)]
        L.orc_morton_key.restype = ctypes.c_uint64
        L.orc_morton_key.argtypes = [P(Space), c_vp]
        L.orc_det_log.restype = c_dbl
        L.orc_det_log.argtypes = [c_dbl]
        L.orc_det_exp.restype = c_dbl
        L.orc_det_exp.argtypes = [c_dbl]
        L.orc_det_sincos2pi.argtypes = [c_dbl, P(c_dbl), P(c_dbl)]
        L.orc_mc_run.argtypes = [P(McProblem), P(Checker), ctypes.c_uint64, c_i64, c_i64, c_vp, P(c_i64), c_vp, c_vp]
        L.orc_lqg_setup.argtypes = [ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.orc_lqg_setup.restype = ctypes.c_int
        L.orc_lqg_cost_terms.argtypes = [c_vp, c_vp, c_vp, c_dbl, c_vp]
        L.orc_lqg_steer.argtypes = [c_vp, c_vp, c_vp, c_dbl, P(c_dbl), P(c_dbl)]
        L.orc_lqg_state.argtypes = [c_vp, c_vp, c_vp, c_dbl, c_dbl, c_vp]
        L.orc_lqg_inball.argtypes = [c_vp, c_vp, c_i64, c_dbl, ctypes.c_int, c_i64, c_i64, c_vp, c_vp, c_vp]
        L.orc_lqg_is_free_motion.argtypes = [P(Checker), P(Space), c_vp, c_dbl, c_vp, c_vp, P(c_i64)]
        L.orc_lqg_is_free_motion.restype = ctypes.c_int
        L.orc_lqg_edges_free_csc.argtypes = [P(Checker), P(Space), c_vp, c_dbl, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp,
                                             P(c_i64)]
        L.orc_close_points.argtypes = [P(Checker), c_vp, c_vp, c_i64, ctypes.c_int, c_dbl, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.orc_close_points.restype = ctypes.c_int
        for f in ("orc_mod2pi", "orc_det_sin", "orc_det_cos", "orc_det_acos"):
            getattr(L, f).restype = c_dbl
            getattr(L, f).argtypes = [c_dbl]
        L.orc_det_atan2.restype = c_dbl
        L.orc_det_atan2.argtypes = [c_dbl, c_dbl]
        L.orc_car_steer.restype = c_dbl
        L.orc_car_steer.argtypes = [ctypes.c_int, c_dbl, c_dbl, c_vp, c_vp, P(ctypes.c_int), c_vp]
        L.orc_car_propagate.argtypes = [c_vp, c_vp, c_vp]
        L.orc_car_chopped.restype = c_dbl
        L.orc_car_chopped.argtypes = [ctypes.c_int, c_dbl, c_dbl, c_vp, c_vp]
        L.orc_car_inball.argtypes = [c_vp, c_i64, ctypes.c_int, c_dbl, c_dbl, c_dbl, ctypes.c_int, c_i64, c_i64, c_vp,
                                     c_vp, c_vp]
        L.orc_car_is_free_motion.argtypes = [P(Checker), P(Space), ctypes.c_int, c_dbl, c_dbl, c_vp, c_vp, P(c_i64)]
        L.orc_car_edges_free_csc.argtypes = [P(Checker), P(Space), ctypes.c_int, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_i64,
                                             c_i64, c_vp, P(c_i64)]
        L.orc_lq_steer.argtypes = [ctypes.c_int, c_vp, c_vp, c_vp, c_dbl, P(c_dbl), P(c_dbl)]
        L.orc_lq_cost_terms.argtypes = [ctypes.c_int, c_vp, c_vp, c_vp, c_dbl, c_vp]
        L.orc_lq_state.argtypes = [ctypes.c_int, c_vp, c_vp, c_dbl, c_dbl, c_vp]
        L.orc_lq_inball.argtypes = [c_vp, c_i64, ctypes.c_int, c_vp, c_dbl, ctypes.c_int, c_i64, c_i64, c_vp, c_vp, c_vp]
        L.orc_lq_is_free_motion.argtypes = [P(Checker), P(Space), ctypes.c_int, c_vp, c_dbl, c_vp, c_vp, P(c_i64)]
        L.orc_lq_edges_free_csc.argtypes = [P(Checker), P(Space), ctypes.c_int, c_vp, c_dbl, c_vp, c_vp, c_vp, c_i64,
                                            c_i64, c_vp, P(c_i64)]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---- 2-D obstacle tables built by the oracle's own constructors ---------------------------
def circle_record(c, r):
    out = np.zeros(7)
    if lib().orc_circle_build(float(c[0]), float(c[1]), float(r), _p(out)) != 0:
        raise ValueError("Radius must be positive")
    return out


def polygon_record(points):
    pts = _f64(points).reshape(-1, 2)
    out = np.zeros(4 + 6 * len(pts))
    rc = lib().orc_polygon_build(_p(pts), len(pts), _p(out))
    if rc == -1:
        raise ValueError("Polygons need at least 3 points! Try Line?")
    if rc == -2:
        raise ValueError("Polygon must be convex")
    return out


class Obstacles2D:
    """spec: ("compound", [specs...]) | ("circle", (cx, cy), r) | ("polygon", [(x, y), ...])"""

    def __init__(self, spec, fixed_point_test=False):
        gp, ga, kinds, gates, offs, data = [], [], [], [], [0], []

        def walk(s, parent):
            if s[0] == "compound":
                g = len(gp)
                gp.append(parent)
                ga.append(None)
                recs = [walk(c, g) for c in s[1]]
                if recs:  # SAT2D.jl:91-95
                    ga[g] = [min(r[0] for r in recs), max(r[1] for r in recs), min(r[2] for r in recs),
                             max(r[3] for r in recs)]
                else:  # SAT2D.jl:89
                    ga[g] = [0.0, 0.0, 0.0, 0.0]
                return ga[g]
            if s[0] == "circle":
                rec = circle_record(s[1], s[2])
                kinds.append(0)
                aabb = [rec[3], rec[4], rec[5], rec[6]]
            else:
                rec = polygon_record(s[1])
                kinds.append(1)
                aabb = [rec[0], rec[1], rec[2], rec[3]]
            gates.append(parent)
            data.extend(rec.tolist())
            offs.append(len(data))
            return aabb

        walk(spec, -1)
        self.gate_parent = np.asarray(gp, dtype=np.int32)
        self.gate_aabb = _f64(np.asarray(ga, dtype=np.float64).reshape(-1))
        self.shape_kind = np.asarray(kinds, dtype=np.int32)
        self.shape_gate = np.asarray(gates, dtype=np.int32)
        self.shape_off = np.asarray(offs, dtype=np.int32)
        self.data = _f64(data)
        self.c = Obs2D(len(gp), _p(self.gate_parent), _p(self.gate_aabb), len(kinds), _p(self.shape_kind),
                       _p(self.shape_gate), _p(self.shape_off), _p(self.data), 1 if fixed_point_test else 0)

    def points_free(self, Pts):
        Pts = _f64(Pts).reshape(-1, 2)
        out = np.zeros(len(Pts), dtype=np.uint8)
        lib().orc_points_free_2d(ctypes.byref(self.c), _p(Pts), len(Pts), _p(out))
        return out.astype(bool)

    def segments_free(self, V, W):
        V, W = _f64(V).reshape(-1, 2), _f64(W).reshape(-1, 2)
        out = np.zeros(len(V), dtype=np.uint8)
        lib().orc_segments_free_2d(ctypes.byref(self.c), _p(V), _p(W), len(V), _p(out))
        return out.astype(bool)

    def checker(self):
        return Checker(0, ctypes.pointer(self.c), None, None, 0, 2)


def spec_from_shape(shape):
    """Raw constructor inputs of a product-side shape tree (reads c/r, points, parts only)."""
    name = type(shape).__name__
    if name == "Compound2D":
        return ("compound", [spec_from_shape(p) for p in shape.parts])
    if name == "Circle":
        return ("circle", tuple(shape.c), shape.r)
    return ("polygon", [tuple(p) for p in shape.points])


class Boxes:
    def __init__(self, box_list):
        """box_list: list of d x 2 [lo hi] matrices (test/obstaclesets/ND.jl) or (lo, hi) pairs"""
        los, his = [], []
        for b in box_list:
            if isinstance(b, tuple):
                lo, hi = b
            else:
                b = np.asarray(b, dtype=np.float64)
                lo, hi = b[:, 0], b[:, 1]
            los.append(np.asarray(lo, dtype=np.float64))
            his.append(np.asarray(hi, dtype=np.float64))
        self.M = len(los)
        self.d = len(los[0]) if los else 1
        self.lo = _f64(np.stack(los)) if los else np.zeros((0, 1))
        self.hi = _f64(np.stack(his)) if his else np.zeros((0, 1))

    def points_free(self, Pts):
        Pts = _f64(Pts).reshape(-1, self.d)
        out = np.zeros(len(Pts), dtype=np.uint8)
        lib().orc_points_free_boxes(_p(self.lo), _p(self.hi), self.M, self.d, _p(Pts), len(Pts), _p(out))
        return out.astype(bool)

    def segments_free(self, V, W):
        V, W = _f64(V).reshape(-1, self.d), _f64(W).reshape(-1, self.d)
        out = np.zeros(len(V), dtype=np.uint8)
        lib().orc_segments_free_boxes(_p(self.lo), _p(self.hi), self.M, self.d, _p(V), _p(W), len(V), _p(out))
        return out.astype(bool)

    def checker(self):
        return Checker(1, None, _p(self.lo), _p(self.hi), self.M, self.d)


class StateSpace:
    """lo/hi bounds + s2w: None (Identity) | ("view", [1-based inds]) | ("matrix", C)"""

    def __init__(self, lo, hi, s2w=None):
        self.lo, self.hi = _f64(lo), _f64(hi)
        n = len(self.lo)
        self.inds = self.C = None
        kind, dw = 0, n
        if s2w is not None and s2w[0] == "view":
            kind, self.inds = 1, np.asarray([i - 1 for i in s2w[1]], dtype=np.int32)
            dw = len(self.inds)
        elif s2w is not None and s2w[0] == "matrix":
            C = np.asarray(s2w[1], dtype=np.float64)
            kind, dw = 2, C.shape[0]
            self.C = np.ascontiguousarray(C.ravel(order="F"))
        self.n, self.dw = n, dw
        self.c = Space(n, _p(self.lo), _p(self.hi), kind, dw, _p(self.inds), _p(self.C))


def states_free(obs, space, Pts):
    Pts = _f64(Pts).reshape(-1, space.n)
    out = np.zeros(len(Pts), dtype=np.uint8)
    cc = obs.checker()
    lib().orc_states_free(ctypes.byref(cc), ctypes.byref(space.c), _p(Pts), len(Pts), _p(out))
    return out.astype(bool)


def morton_key(space, x):
    x = _f64(x)
    return int(lib().orc_morton_key(ctypes.byref(space.c), _p(x)))


def sample_free(obs, space, N, seed, max_candidates=None, order=0):
    """first N free candidates of the Philox candidate stream (oracle/sample.c) -> (V[N x n], candidates used);
    order=1: stably sorted by Morton key"""
    V = np.zeros((N, space.n))
    cc = obs.checker()
    used = c_i64(0)
    mc = int(max_candidates) if max_candidates is not None else (1 << 62)
    got = lib().orc_sample_free(ctypes.byref(cc), ctypes.byref(space.c), N, seed, mc, int(order), _p(V), ctypes.byref(used))
    return V[:got], used.value


def sample_candidate(space, seed, c):
    x = np.zeros(space.n)
    lib().orc_sample_candidate(ctypes.byref(space.c), seed, c, _p(x))
    return x


def motions_free_straight(obs, space, V, W):
    """is_free_motion(v, w, CC, SS) per pair; returns (bool[n], CC.count increment)."""
    V, W = _f64(V).reshape(-1, space.n), _f64(W).reshape(-1, space.n)
    cc = obs.checker()
    cnt = c_i64(0)
    out = np.zeros(len(V), dtype=bool)
    L = lib()
    for i in range(len(V)):
        out[i] = bool(L.orc_is_free_motion_straight(ctypes.byref(cc), ctypes.byref(space.c), _p(V[i]), _p(W[i]),
                                                    ctypes.byref(cnt)))
    return out, cnt.value


def edges_free_csc(obs, space, V, colptr, rowval, c0=0):
    """validity per stored entry (row y -> column x); returns (uint8[nnz], count)"""
    V = _f64(V)
    colptr = np.ascontiguousarray(colptr, dtype=np.int64)
    rowval = np.ascontiguousarray(rowval, dtype=np.int64)
    ncols = len(colptr) - 1
    out = np.zeros(len(rowval), dtype=np.uint8)
    cc = obs.checker()
    cnt = c_i64(0)
    lib().orc_edges_free_csc(ctypes.byref(cc), ctypes.byref(space.c), _p(V), len(V), _p(colptr), _p(rowval), c0,
                             c0 + ncols, _p(out), ctypes.byref(cnt))
    return out, cnt.value


# ---- r-ball ------------------------------------------------------------------------------------
def rball_brute(V, r, pred=0, q0=0, q1=None):
    V = _f64(V)
    N, d = V.shape
    q1 = N if q1 is None else q1
    colptr = np.zeros(q1 - q0 + 1, dtype=np.int64)
    lib().orc_rball_brute(_p(V), N, d, float(r), pred, q0, q1, _p(colptr), None, None)
    nnz = int(colptr[-1] - 1)
    rowval = np.zeros(nnz, dtype=np.int64)
    nzval = np.zeros(nnz, dtype=np.float64)
    lib().orc_rball_brute(_p(V), N, d, float(r), pred, q0, q1, _p(colptr), _p(rowval), _p(nzval))
    return colptr, rowval, nzval


class KDTree:
    def __init__(self, V, leafsize=10):
        self.V = _f64(V)
        self.h = lib().orc_kdtree_build(_p(self.V), self.V.shape[0], self.V.shape[1], leafsize)

    def __del__(self):
        try:
            lib().orc_kdtree_free(self.h)
        except Exception:
            pass

    def rball(self, r, q0=0, q1=None):
        N = self.V.shape[0]
        q1 = N if q1 is None else q1
        colptr = np.zeros(q1 - q0 + 1, dtype=np.int64)
        lib().orc_rball_kdtree(self.h, float(r), q0, q1, _p(colptr), None, None)
        nnz = int(colptr[-1] - 1)
        rowval = np.zeros(nnz, dtype=np.int64)
        nzval = np.zeros(nnz, dtype=np.float64)
        lib().orc_rball_kdtree(self.h, float(r), q0, q1, _p(colptr), _p(rowval), _p(nzval))
        return colptr, rowval, nzval


# ---- linear-quadratic (double integrator) ----------------------------------------------------------
class DoubleIntegratorLQ:
    """d-dimensional double integrator with control penalty R (d x d SPD, default identity)."""

    def __init__(self, d, R=None):
        self.d = d
        self.R = _f64(np.eye(d) if R is None else R)

    def steer(self, x0, x1, r):
        x0, x1 = _f64(x0), _f64(x1)
        c, t = c_dbl(0), c_dbl(0)
        lib().orc_lq_steer(self.d, _p(self.R), _p(x0), _p(x1), float(r), ctypes.byref(c), ctypes.byref(t))
        return c.value, t.value

    def cost_terms(self, x0, x1, t):
        x0, x1 = _f64(x0), _f64(x1)
        out = np.zeros(3)
        lib().orc_lq_cost_terms(self.d, _p(self.R), _p(x0), _p(x1), float(t), _p(out))
        return out

    def state(self, x0, x1, t, s):
        x0, x1 = _f64(x0), _f64(x1)
        out = np.zeros(2 * self.d)
        lib().orc_lq_state(self.d, _p(x0), _p(x1), float(t), float(s), _p(out))
        return out

    def inball(self, V, r, forwards, q0=0, q1=None):
        V = _f64(V)
        N = V.shape[0]
        q1 = N if q1 is None else q1
        colptr = np.zeros(q1 - q0 + 1, dtype=np.int64)
        lib().orc_lq_inball(_p(V), N, self.d, _p(self.R), float(r), int(forwards), q0, q1, _p(colptr), None, None)
        nnz = int(colptr[-1] - 1)
        rowval, nzval = np.zeros(nnz, dtype=np.int64), np.zeros(nnz)
        lib().orc_lq_inball(_p(V), N, self.d, _p(self.R), float(r), int(forwards), q0, q1, _p(colptr), _p(rowval),
                            _p(nzval))
        return colptr, rowval, nzval

    def is_free_motion(self, obs, space, r, v, w):
        v, w = _f64(v), _f64(w)
        cc = obs.checker()
        cnt = c_i64(0)
        ok = lib().orc_lq_is_free_motion(ctypes.byref(cc), ctypes.byref(space.c), self.d, _p(self.R), float(r), _p(v),
                                         _p(w), ctypes.byref(cnt))
        return bool(ok), cnt.value

    def edges_free_csc(self, obs, space, r, V, colptr, rowval, c0=0):
        V = _f64(V)
        colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        rowval = np.ascontiguousarray(rowval, dtype=np.int64)
        out = np.zeros(len(rowval), dtype=np.uint8)
        cc = obs.checker()
        cnt = c_i64(0)
        lib().orc_lq_edges_free_csc(ctypes.byref(cc), ctypes.byref(space.c), self.d, _p(self.R), float(r), _p(V),
                                    _p(colptr), _p(rowval), c0, c0 + len(colptr) - 1, _p(out), ctypes.byref(cnt))
        return out, cnt.value


class LqgTables(ctypes.Structure):
    _MAXN = 6
    _fields_ = [("n", c_i32), ("np", c_i32), ("A", c_dbl * 36), ("c", c_dbl * 6), ("BRB", c_dbl * 36),
                ("Ak", (c_dbl * 36) * 6), ("dk", (c_dbl * 6) * 6), ("Gp", (c_dbl * 36) * 12)]


class LinearQuadraticGeneral:
    """xdot = A x + B u + c with nilpotent A (oracle/lq_general.c); same methods as DoubleIntegratorLQ."""

    def __init__(self, A, B, c, R):
        A, B, c, R = _f64(A), _f64(B), _f64(c), _f64(R)
        self.n, self.m = A.shape[0], B.shape[1]
        self.S = LqgTables()
        rc = lib().orc_lqg_setup(self.n, self.m, _p(A), _p(B), _p(c), _p(R), ctypes.byref(self.S))
        if rc != 0:
            raise ValueError("orc_lqg_setup failed (%d): A not nilpotent / R not SPD / size out of range" % rc)

    def steer(self, x0, x1, r):
        x0, x1 = _f64(x0), _f64(x1)
        c, t = c_dbl(0), c_dbl(0)
        lib().orc_lqg_steer(ctypes.byref(self.S), _p(x0), _p(x1), float(r), ctypes.byref(c), ctypes.byref(t))
        return c.value, t.value

    def cost_terms(self, x0, x1, t):
        x0, x1 = _f64(x0), _f64(x1)
        out = np.zeros(3)
        lib().orc_lqg_cost_terms(ctypes.byref(self.S), _p(x0), _p(x1), float(t), _p(out))
        return out

    def state(self, x0, x1, t, s):
        x0, x1 = _f64(x0), _f64(x1)
        out = np.zeros(self.n)
        lib().orc_lqg_state(ctypes.byref(self.S), _p(x0), _p(x1), float(t), float(s), _p(out))
        return out

    def inball(self, V, r, forwards, q0=0, q1=None):
        V = _f64(V)
        N = V.shape[0]
        q1 = N if q1 is None else q1
        colptr = np.zeros(q1 - q0 + 1, dtype=np.int64)
        lib().orc_lqg_inball(ctypes.byref(self.S), _p(V), N, float(r), int(forwards), q0, q1, _p(colptr), None, None)
        nnz = int(colptr[-1] - 1)
        rowval, nzval = np.zeros(nnz, dtype=np.int64), np.zeros(nnz)
        lib().orc_lqg_inball(ctypes.byref(self.S), _p(V), N, float(r), int(forwards), q0, q1, _p(colptr), _p(rowval),
                             _p(nzval))
        return colptr, rowval, nzval

    def is_free_motion(self, obs, space, r, v, w):
        v, w = _f64(v), _f64(w)
        cc = obs.checker()
        cnt = c_i64(0)
        ok = lib().orc_lqg_is_free_motion(ctypes.byref(cc), ctypes.byref(space.c), ctypes.byref(self.S), float(r), _p(v),
                                          _p(w), ctypes.byref(cnt))
        return bool(ok), cnt.value

    def edges_free_csc(self, obs, space, r, V, colptr, rowval, c0=0):
        V = _f64(V)
        colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        rowval = np.ascontiguousarray(rowval, dtype=np.int64)
        out = np.zeros(len(rowval), dtype=np.uint8)
        cc = obs.checker()
        cnt = c_i64(0)
        lib().orc_lqg_edges_free_csc(ctypes.byref(cc), ctypes.byref(space.c), ctypes.byref(self.S), float(r), _p(V),
                                     _p(colptr), _p(rowval), c0, c0 + len(colptr) - 1, _p(out), ctypes.byref(cnt))
        return out, cnt.value


class SimpleCar:
    """ReedsSheppExact (kind 0) / DubinsExact (kind 1) with turning radius r and speed s (oracle/cars.c)"""

    def __init__(self, kind, r=1.0, s=1.0):
        self.kind = {"reedsshepp": 0, "dubins": 1}.get(kind, kind)
        self.r, self.s = float(r), float(s)

    def steer(self, v, w):
        """-> (cost, segments[l, 3]) = (path length, steering_control as (duration, speed, curvature) rows)"""
        v, w = _f64(v), _f64(w)
        n = ctypes.c_int(0)
        segs = np.zeros(15)
        c = lib().orc_car_steer(self.kind, self.r, self.s, _p(v), _p(w), ctypes.byref(n), _p(segs))
        return c, segs.reshape(5, 3)[:n.value].copy()

    def propagate(self, v, u):
        v, u = _f64(v), _f64(u)
        out = np.zeros(3)
        lib().orc_car_propagate(_p(v), _p(u), _p(out))
        return out

    def chopped(self, v, w, chopval):
        v, w = _f64(v), _f64(w)
        return lib().orc_car_chopped(self.kind, self.r, float(chopval), _p(v), _p(w))

    def inball(self, V, r, forwards=True, chopval=None, q0=0, q1=None):
        V = _f64(V)
        N = V.shape[0]
        q1 = N if q1 is None else q1
        chop = float(r if chopval is None else chopval)
        colptr = np.zeros(q1 - q0 + 1, dtype=np.int64)
        lib().orc_car_inball(_p(V), N, self.kind, self.r, float(r), chop, int(forwards), q0, q1, _p(colptr), None, None)
        nnz = int(colptr[-1] - 1)
        rowval, nzval = np.zeros(nnz, dtype=np.int64), np.zeros(nnz)
        lib().orc_car_inball(_p(V), N, self.kind, self.r, float(r), chop, int(forwards), q0, q1, _p(colptr), _p(rowval),
                             _p(nzval))
        return colptr, rowval, nzval

    def is_free_motion(self, obs, space, v, w):
        v, w = _f64(v), _f64(w)
        cc = obs.checker()
        cnt = c_i64(0)
        ok = lib().orc_car_is_free_motion(ctypes.byref(cc), ctypes.byref(space.c), self.kind, self.r, self.s, _p(v), _p(w),
                                          ctypes.byref(cnt))
        return bool(ok), cnt.value

    def edges_free_csc(self, obs, space, V, colptr, rowval, c0=0):
        V = _f64(V)
        colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        rowval = np.ascontiguousarray(rowval, dtype=np.int64)
        out = np.zeros(len(rowval), dtype=np.uint8)
        cc = obs.checker()
        cnt = c_i64(0)
        lib().orc_car_edges_free_csc(ctypes.byref(cc), ctypes.byref(space.c), self.kind, self.r, self.s, _p(V), _p(colptr),
                                     _p(rowval), c0, c0 + len(colptr) - 1, _p(out), ctypes.byref(cnt))
        return out, cnt.value


# ---- k-nearest connections: specification of csrc/knn.cu (the reference defines nothing; parity unpinned) ---------
def knn_from_values(vals, k, skip):
    """indices (ascending) of the k smallest entries of vals, ties towards the smaller index, `skip` excluded"""
    idx = np.arange(len(vals))
    keep = idx != skip
    order = np.lexsort((idx[keep], vals[keep]))[:k]
    return np.sort(idx[keep][order])


def knn_brute(V, k):
    """k-NN table of a Euclidean sample set as (colptr, rowval, nzval), 1-based: distances in the reference's order
    (sum over coordinates left to right, one rounding per operation, then sqrt), k nearest by (distance, index)"""
    V = _f64(V)
    N, d = V.shape
    k = min(k, N - 1)
    colptr = 1 + k * np.arange(N + 1, dtype=np.int64)
    rowval = np.zeros(N * k, dtype=np.int64)
    nzval = np.zeros(N * k)
    for q in range(N):
        t = V[q, 0] - V[:, 0]
        s = t * t
        for i in range(1, d):
            t = V[q, i] - V[:, i]
            s = s + t * t
        dist = np.sqrt(s)
        sel = knn_from_values(dist, k, q)
        rowval[q * k:(q + 1) * k] = sel + 1
        nzval[q * k:(q + 1) * k] = dist[sel]
    return colptr, rowval, nzval


def knn_of_table(colptr, rowval, nzval, k):
    """the same selection applied to the columns of any table (what mpb200_table_knn does)"""
    cp = [1]
    rv, nz = [], []
    for w in range(len(colptr) - 1):
        a, b = colptr[w] - 1, colptr[w + 1] - 1
        rows, vals = rowval[a:b], nzval[a:b]
        if len(rows) > k:
            order = np.lexsort((np.arange(len(rows)), vals))[:k]
            order = np.sort(order)
            rows, vals = rows[order], vals[order]
        rv.append(rows); nz.append(vals); cp.append(cp[-1] + len(rows))
    return np.asarray(cp, dtype=np.int64), np.concatenate(rv).astype(np.int64), np.concatenate(nz)


def union_transpose(A, B, N):
    """out[:, v] = A[:, v] U { w : v in B[:, w] }, values from A else from B (what mpb200_table_union_transpose does)"""
    cols = [dict() for _ in range(N)]
    bcp, brv, bnz = B
    for w in range(N):
        for e in range(bcp[w] - 1, bcp[w + 1] - 1):
            cols[brv[e] - 1][w + 1] = bnz[e]
    acp, arv, anz = A
    for v in range(N):
        for e in range(acp[v] - 1, acp[v + 1] - 1):
            cols[v][int(arv[e])] = anz[e]
    cp = [1]
    rv, nz = [], []
    for v in range(N):
        rows = sorted(cols[v])
        rv.extend(rows); nz.extend(cols[v][r_] for r_ in rows); cp.append(cp[-1] + len(rows))
    return np.asarray(cp, dtype=np.int64), np.asarray(rv, dtype=np.int64), np.asarray(nz)


def close_points(obs, P, Ws, r2, want_all=False):
    """closeR(p, CC, W, r2) for every row of P with its own W (oracle/closest.c) ->
    (count[n], d2[n,S], shape[n,S], x[n,S,dw]) [+ (all_d2[n,S], all_x[n,S,dw])]"""
    P, Ws = _f64(P), _f64(Ws)
    n, dw = P.shape
    cc = obs.checker()
    S = cc.obs2d.contents.n_shapes if cc.kind == 0 else cc.M
    count = np.zeros(n, dtype=np.int32)
    d2 = np.full((n, max(S, 1)), np.inf)
    shape = np.full((n, max(S, 1)), -1, dtype=np.int32)
    x = np.zeros((n, max(S, 1), dw))
    all_d2 = np.zeros((n, max(S, 1))) if want_all else None
    all_x = np.zeros((n, max(S, 1), dw)) if want_all else None
    rc = lib().orc_close_points(ctypes.byref(cc), _p(P), _p(Ws), n, dw, float(r2), _p(count), _p(d2), _p(shape), _p(x),
                                _p(all_d2), _p(all_x))
    if rc != 0:
        raise ValueError("orc_close_points: unsupported workspace dimension")
    return (count, d2, shape, x) + ((all_d2, all_x) if want_all else ())


# ---- Monte-Carlo collision probability (spec: oracle/mc.c header; parity unpinned) -------------------
def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def det_log(x):
    return lib().orc_det_log(float(x))


def det_exp(x):
    return lib().orc_det_exp(float(x))


def det_sincos2pi(u):
    s, c = c_dbl(0), c_dbl(0)
    lib().orc_det_sincos2pi(float(u), ctypes.byref(s), ctypes.byref(c))
    return s.value, c.value


class McSpec:
    """Arrays of the estimator's problem statement (see oracle/mc.c)."""

    def __init__(self, F, G, Wz, wbar, alpha, mu, swept=False):
        self.F, self.G = _f64(F), _f64(G)                    # T x nz x nz, T x nz x q
        self.Wz, self.wbar = _f64(Wz), _f64(wbar)            # dw x nz, (T+1) x dw
        self.alpha = _f64(alpha)                             # K+1
        T, nz, q = self.F.shape[0], self.F.shape[1], self.G.shape[2]
        K = len(self.alpha) - 1
        self.mu = _f64(np.asarray(mu, dtype=np.float64).reshape(K, T * q)) if K else np.zeros((0, T * q))
        self.T, self.nz, self.q, self.dw, self.K, self.swept = T, nz, q, self.Wz.shape[0], K, bool(swept)
        self.c = McProblem(T, nz, q, self.dw, _p(self.F), _p(self.G), _p(self.Wz), _p(self.wbar), K, _p(self.alpha),
                           _p(self.mu), int(self.swept))


def mc_run(spec, obs, seed, first, n, per_rollout=False):
    sums = np.zeros(3)
    hits = c_i64(0)
    hit = np.zeros(n, dtype=np.uint8) if per_rollout else None
    w = np.zeros(n) if per_rollout else None
    cc = obs.checker()
    lib().orc_mc_run(ctypes.byref(spec.c), ctypes.byref(cc), seed, first, n, _p(sums), ctypes.byref(hits), _p(hit), _p(w))
    res = dict(S1=sums[0], S2=sums[1], S0=sums[2], n=n, hits=hits.value)
    if per_rollout:
        res["hit"], res["w"] = hit, w
    return res
