// predicates.cuh -- device collision predicates, evaluated exactly as the reference does.
//
// Every floating-point operation is an explicit round-to-nearest intrinsic
// (__dmul_rn/__dadd_rn/__dsub_rn/__ddiv_rn), so nvcc can never contract a*b+c into a
// DFMA: the reference (pure Julia, no muladd/@fastmath) rounds after every operation and
// collision booleans must match it bit for bit.
//
//   2-D SAT:  src/collisioncheckers/SAT2D.jl:69-76 (Line), :111-132 (point tests, incl. the
//             inverted point-in-polygon test :124-127), :158-180 (swept tests);
//             src/utilities/vec2Dutils.jl:5-36.
//   N-d box:  src/collisioncheckers/boxesND.jl:42-56.
//   wrappers: src/statespaces.jl:57-60,150-158.
//
// Packed 2-D obstacle table: one buffer of 8-byte words, staged into shared memory.
//   int view  I = (const int*)T:  I[0]=n_gates G  I[1]=n_shapes S  I[2]=flags  I[3]=index (in
//             doubles) of the gate AABBs;  I[4+g] = parent of gate g (-1 = root);
//             I[4+G+4s .. +3] = kind, gate, off (index into T, doubles), K of shape s
//             flags bit0: intended point-in-polygon test; bit1: gates are consistent (every gate
//             AABB contains the AABBs of everything below it -- true for tables built by the
//             reference constructors), which lets polygons skip the gate chain
//   doubles   T[I[3] + 4g .. +3] = xlo xhi ylo yhi of gate g
//             T[I[3] + 4G + 4s .. +3] = cull box of shape s (polygon: own AABB; circle: innermost
//             gate AABB or +-inf) -- a shape can only matter to edges whose AABB overlaps it
//   Circle  data at T[off]: cx cy r xlo xhi ylo yhi
//   Polygon data at T[off]: xlo xhi ylo yhi pts[2K] normals[2K] nextrema[2K]
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mpb {

constexpr int kMaxDim = 16;    // largest state / workspace dimension the kernels unroll for
constexpr int kMaxGates = 32;  // nesting depth x breadth of Compound2D nodes

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// vec2Dutils.jl:5,7,34-36
__device__ __forceinline__ double dot2(double a1, double a2, double b1, double b2) {
    return dadd(dmul(a1, b1), dmul(a2, b2));
}
__device__ __forceinline__ double cross2(double a1, double a2, double b1, double b2) {
    return dsub(dmul(a1, b2), dmul(a2, b1));
}
__device__ __forceinline__ bool overlapping(double i1, double i2, double j1, double j2) {
    return i1 <= j2 && j1 <= i2;
}
__device__ __forceinline__ bool ininterval(double x, double i1, double i2) { return i1 <= x && x <= i2; }

struct Line2 {  // SAT2D.jl:60-77
    double v1, v2, w1, w2, e1, e2, n1, n2, xl, xh, yl, yh, ndotv;
};
__device__ __forceinline__ Line2 make_line(double v1, double v2, double w1, double w2) {
    Line2 L;
    L.v1 = v1; L.v2 = v2; L.w1 = w1; L.w2 = w2;
    L.e1 = dsub(w1, v1); L.e2 = dsub(w2, v2);
    L.n1 = L.e2; L.n2 = -L.e1;  // perp(edge), vec2Dutils.jl:6
    if (v1 < w1) { L.xl = v1; L.xh = w1; } else { L.xl = w1; L.xh = v1; }  // minmaxV
    if (v2 < w2) { L.yl = v2; L.yh = w2; } else { L.yl = w2; L.yh = v2; }
    L.ndotv = dot2(v1, v2, L.n1, L.n2);
    return L;
}

// SAT2D.jl:122
__device__ __forceinline__ bool point_circle(const double *c, double px, double py) {
    double d1 = dsub(px, c[0]), d2 = dsub(py, c[1]);
    return dot2(d1, d2, d1, d2) <= dmul(c[2], c[2]);
}
// SAT2D.jl:124-127 (inverted on purpose; `fixed` selects the intended test)
__device__ __forceinline__ bool point_polygon(const double *P, int K, double px, double py, bool fixed) {
    if (!(ininterval(px, P[0], P[1]) && ininterval(py, P[2], P[3]))) return false;
    const double *nrm = P + 4 + 2 * K, *ext = P + 4 + 4 * K;
    for (int i = 0; i < K; ++i) {
        bool in = ininterval(dot2(px, py, nrm[2 * i], nrm[2 * i + 1]), ext[2 * i], ext[2 * i + 1]);
        if (fixed ? !in : in) return false;
    }
    return true;
}
// SAT2D.jl:165-171
__device__ __forceinline__ bool line_circle_ends_free(const Line2 &L, const double *c) {
    if (!(overlapping(L.xl, L.xh, c[3], c[4]) && overlapping(L.yl, L.yh, c[5], c[6]))) return false;
    double vc1 = dsub(c[0], L.v1), vc2 = dsub(c[1], L.v2);
    double d2 = dot2(L.e1, L.e2, L.e1, L.e2);
    double cr = cross2(L.e1, L.e2, vc1, vc2);
    if (dmul(d2, dmul(c[2], c[2])) < dmul(cr, cr)) return false;
    double t = dot2(vc1, vc2, L.e1, L.e2);
    return 0.0 <= t && t <= d2;
}
// SAT2D.jl:172-176 (+ :113-114, vec2Dutils.jl:18-27)
__device__ __forceinline__ bool line_polygon_ends_free(const Line2 &L, const double *P, int K) {
    if (!(overlapping(L.xl, L.xh, P[0], P[1]) && overlapping(L.yl, L.yh, P[2], P[3]))) return false;
    const double *pts = P + 4, *nrm = P + 4 + 2 * K, *ext = P + 4 + 4 * K;
    double mn = __longlong_as_double(0x7ff0000000000000LL), mx = -mn;
    for (int i = 0; i < K; ++i) {
        double d = dot2(pts[2 * i], pts[2 * i + 1], L.n1, L.n2);
        if (d < mn) mn = d;
        if (d > mx) mx = d;
    }
    if (!ininterval(L.ndotv, mn, mx)) return false;
    for (int i = 0; i < K; ++i) {
        double a = dot2(L.v1, L.v2, nrm[2 * i], nrm[2 * i + 1]);
        double b = dot2(L.w1, L.w2, nrm[2 * i], nrm[2 * i + 1]);
        double lo = (a < b) ? a : b, hi = (a < b) ? b : a;
        if (!overlapping(ext[2 * i], ext[2 * i + 1], lo, hi)) return false;
    }
    return true;
}

struct Obs2 {  // decoded header of the packed table
    const double *T;
    const int *I;
    int G, S;
    bool fixed;
    __device__ __forceinline__ explicit Obs2(const double *t, bool valid = true)
        : T(t), I(reinterpret_cast<const int *>(t)), G(0), S(0), fixed(false) {
        if (valid) { G = I[0]; S = I[1]; fixed = I[2] & 1; }
    }
    __device__ __forceinline__ const double *gate_aabb(int g) const { return T + I[3] + 4 * g; }
    __device__ __forceinline__ const double *cull_box(int s) const { return T + I[3] + 4 * G + 4 * s; }
    __device__ __forceinline__ bool consistent() const { return (I[2] & 2) != 0; }
    __device__ __forceinline__ int gate_parent(int g) const { return I[4 + g]; }
    __device__ __forceinline__ int kind(int s) const { return I[4 + G + 4 * s]; }
    __device__ __forceinline__ int gate(int s) const { return I[4 + G + 4 * s + 1]; }
    __device__ __forceinline__ const double *data(int s) const { return T + I[4 + G + 4 * s + 2]; }
    __device__ __forceinline__ int K(int s) const { return I[4 + G + 4 * s + 3]; }
};

// gate chain for a point (SAT2D.jl:129-132) / for a line's AABB (SAT2D.jl:158-161,119)
__device__ __forceinline__ uint32_t gates_point(const Obs2 &O, double px, double py) {
    uint32_t pass = 0;
    for (int g = 0; g < O.G; ++g) {
        const double *a = O.gate_aabb(g);
        int par = O.gate_parent(g);
        bool ok = (par < 0 || ((pass >> par) & 1u)) && ininterval(px, a[0], a[1]) && ininterval(py, a[2], a[3]);
        pass |= (uint32_t)ok << g;
    }
    return pass;
}
__device__ __forceinline__ uint32_t gates_line(const Obs2 &O, const Line2 &L) {
    uint32_t pass = 0;
    for (int g = 0; g < O.G; ++g) {
        const double *a = O.gate_aabb(g);
        int par = O.gate_parent(g);
        bool ok = (par < 0 || ((pass >> par) & 1u)) && overlapping(a[0], a[1], L.xl, L.xh) &&
                  overlapping(a[2], a[3], L.yl, L.yh);
        pass |= (uint32_t)ok << g;
    }
    return pass;
}
// one basic shape against a line: colliding(L, B) = ends_free || point(v) || point(w), SAT2D.jl:178
__device__ __forceinline__ bool line_shape(const Obs2 &O, int s, const Line2 &L) {
    const double *D = O.data(s);
    if (O.kind(s) == 0)
        return line_circle_ends_free(L, D) || point_circle(D, L.v1, L.v2) || point_circle(D, L.w1, L.w2);
    int K = O.K(s);
    return line_polygon_ends_free(L, D, K) || point_polygon(D, K, L.v1, L.v2, O.fixed) ||
           point_polygon(D, K, L.w1, L.w2, O.fixed);
}

// colliding(p, obstacles): SAT2D.jl:129-132
__device__ __forceinline__ bool point_colliding_2d(const double *T, double px, double py) {
    Obs2 O(T);
    uint32_t pass = gates_point(O, px, py);
    for (int s = 0; s < O.S; ++s) {
        int gate = O.gate(s);
        if (gate >= 0 && !((pass >> gate) & 1u)) continue;
        const double *D = O.data(s);
        if (O.kind(s) == 0) {
            if (point_circle(D, px, py)) return true;
        } else {
            if (point_polygon(D, O.K(s), px, py, O.fixed)) return true;
        }
    }
    return false;
}

// colliding(Line(v,w), obstacles): SAT2D.jl:158-161,178-180
__device__ __forceinline__ bool line_colliding_2d(const double *T, double v1, double v2, double w1, double w2) {
    Obs2 O(T);
    Line2 L = make_line(v1, v2, w1, w2);
    uint32_t pass = gates_line(O, L);
    for (int s = 0; s < O.S; ++s) {
        int gate = O.gate(s);
        if (gate >= 0 && !((pass >> gate) & 1u)) continue;
        if (line_shape(O, s, L)) return true;
    }
    return false;
}

// ---- N-d boxes: table = lo[M*d] then hi[M*d], box-major ---------------------------
// boxesND.jl:42-43
template <int D>
__device__ __forceinline__ bool box_point_free(const double *T, int M, const double *v) {
    const double *lo = T, *hi = T + (size_t)M * D;
    for (int k = 0; k < M; ++k) {
        bool outside = false;
#pragma unroll
        for (int i = 0; i < D; ++i) outside = outside || !(lo[k * D + i] <= v[i] && v[i] <= hi[k * D + i]);
        if (!outside) return false;
    }
    return true;
}
// boxesND.jl:44-51 for ONE box (quirk Q2 kept: one face per axis, no lambda range test, IEEE
// division by zero): broadphase-free || narrow-free
template <int D>
__device__ __forceinline__ bool box_one_segment_free(const double *l, const double *h, const double *v,
                                                     const double *bmin, const double *bmax, const double *dv) {
    bool broad = false;
#pragma unroll
    for (int i = 0; i < D; ++i) broad = broad || (h[i] < bmin[i] || l[i] > bmax[i]);
    if (broad) return true;
    bool hit = false;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double corner = (v[i] < l[i]) ? l[i] : h[i];
        double lam = ddiv(dsub(corner, v[i]), dv[i]);
        bool all = true;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            if (j == i) continue;
            double x = dadd(v[j], dmul(dv[j], lam));
            all = all && (l[j] <= x && x <= h[j]);
        }
        hit = hit || all;
    }
    return !hit;
}
// boxesND.jl:52-56
template <int D>
__device__ __forceinline__ bool box_segment_free(const double *T, int M, const double *v, const double *w) {
    const double *lo = T, *hi = T + (size_t)M * D;
    double bmin[D], bmax[D], dv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        bmin[i] = (w[i] < v[i]) ? w[i] : v[i];
        bmax[i] = (v[i] < w[i]) ? w[i] : v[i];
        dv[i] = dsub(w[i], v[i]);
    }
    for (int k = 0; k < M; ++k)
        if (!box_one_segment_free<D>(lo + k * D, hi + k * D, v, bmin, bmax, dv)) return false;
    return true;
}

// ---- warp-cooperative variants: the 32 lanes hold edges that share one endpoint (a column of
// the neighbour table), so their bounding boxes nearly coincide.  The warp first reduces the
// union box of its edges and culls obstacles against it with one obstacle per lane; an obstacle
// is culled only when EVERY lane's own AABB gate would have rejected it, so per-edge results
// are exactly the reference's.  Circles are never culled by their own AABB: their point test
// (SAT2D.jl:122) has no AABB gate in the reference.
// Conservative warp-wide min / max of doubles: each value is rounded OUTWARD to float, mapped to
// an order-preserving int and reduced with one REDUX instruction.  The result bounds the exact
// extremum from outside, which is all the culling needs (a superset box never culls an obstacle
// that some edge's own AABB test would accept).
__device__ __forceinline__ int float_order_key(float f) {
    int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float order_key_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ __forceinline__ double warp_min(double x) {
    return (double)order_key_float(__reduce_min_sync(0xffffffffu, float_order_key(__double2float_rd(x))));
}
__device__ __forceinline__ double warp_max(double x) {
    return (double)order_key_float(__reduce_max_sync(0xffffffffu, float_order_key(__double2float_ru(x))));
}
// `my_cull` = cull box of shape `lane` (xlo xhi ylo yhi; empty box for lane >= S), preloaded by the
// caller when S <= 32 so the common no-obstacle-nearby column costs a handful of instructions.
__device__ __forceinline__ bool warp_line_colliding_2d(const Obs2 &O, double cull_xl, double cull_xh, double cull_yl,
                                                       double cull_yh, double v1, double v2, double w1, double w2,
                                                       bool run) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int lane = threadIdx.x & 31;
    Line2 L = make_line(v1, v2, w1, w2);
    const double cxl = warp_min(run ? L.xl : inf), cxh = warp_max(run ? L.xh : -inf);
    const double cyl = warp_min(run ? L.yl : inf), cyh = warp_max(run ? L.yh : -inf);
    if (!(cxl <= cxh)) return false;  // no lane runs
    const bool lazy = O.consistent();
    bool have_pass = false, hit = false;
    uint32_t pass = 0;
    for (int s0 = 0; s0 < O.S; s0 += 32) {
        bool cand;
        if (O.S <= 32) {
            cand = overlapping(cxl, cxh, cull_xl, cull_xh) && overlapping(cyl, cyh, cull_yl, cull_yh);
        } else {
            const int s = s0 + lane;
            cand = s < O.S;
            if (cand) {
                const double *cb = O.cull_box(s);
                cand = overlapping(cxl, cxh, cb[0], cb[1]) && overlapping(cyl, cyh, cb[2], cb[3]);
            }
        }
        unsigned mask = __ballot_sync(0xffffffffu, cand);
        while (mask) {
            const int sj = s0 + __ffs(mask) - 1;
            mask &= mask - 1;
            bool gate_ok = true;
            if (!lazy || O.kind(sj) == 0) {  // circles have no AABB gate of their own (SAT2D.jl:122)
                if (!have_pass) { pass = run ? gates_line(O, L) : 0u; have_pass = true; }
                const int gate = O.gate(sj);
                gate_ok = gate < 0 || ((pass >> gate) & 1u);
            }
            if (run && !hit && gate_ok) hit = line_shape(O, sj, L);
        }
    }
    return hit;
}
template <int D>
__device__ __forceinline__ bool warp_box_segment_free(const double *T, int M, const double *v, const double *w,
                                                      bool run) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int lane = threadIdx.x & 31;
    const double *lo = T, *hi = T + (size_t)M * D;
    double bmin[D], bmax[D], dv[D], cmin[D], cmax[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        bmin[i] = (w[i] < v[i]) ? w[i] : v[i];
        bmax[i] = (v[i] < w[i]) ? w[i] : v[i];
        dv[i] = dsub(w[i], v[i]);
        cmin[i] = warp_min(run ? bmin[i] : inf);
        cmax[i] = warp_max(run ? bmax[i] : -inf);
    }
    if (!(cmin[0] <= cmax[0])) return true;  // no lane runs
    bool free_ = true;
    for (int k0 = 0; k0 < M; k0 += 32) {
        const int k = k0 + lane;
        bool cand = k < M;
        if (cand) {
            bool sep = false;  // broadphase against the union box: separated from every edge of the warp
#pragma unroll
            for (int i = 0; i < D; ++i) sep = sep || (hi[k * D + i] < cmin[i] || lo[k * D + i] > cmax[i]);
            cand = !sep;
        }
        unsigned mask = __ballot_sync(0xffffffffu, cand);
        while (mask) {
            const int kj = k0 + __ffs(mask) - 1;
            mask &= mask - 1;
            if (run && free_) free_ = box_one_segment_free<D>(lo + kj * D, hi + kj * D, v, bmin, bmax, dv);
        }
    }
    return free_;
}

// ---- (CC, SS) wrappers: statespaces.jl:57-60,150-158 --------------------------------
struct SpaceDev {  // passed by value as a kernel argument
    int n, s2w_kind, dw;
    double lo[kMaxDim], hi[kMaxDim];
    int inds[kMaxDim];
    double C[kMaxDim * 4];  // dw x n column-major, dw*n <= 64
};
template <int N>
__device__ __forceinline__ bool in_state_space(const SpaceDev &S, const double *v) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) ok = ok && (S.lo[i] <= v[i] && v[i] <= S.hi[i]);
    return ok;
}
template <int N, int DW>
__device__ __forceinline__ void state2workspace(const SpaceDev &S, const double *v, double *p) {
    if (S.s2w_kind == 0) {
#pragma unroll
        for (int i = 0; i < DW; ++i) p[i] = v[i < N ? i : 0];
    } else if (S.s2w_kind == 1) {
#pragma unroll
        for (int i = 0; i < DW; ++i) {
            double x = v[0];
#pragma unroll
            for (int j = 1; j < N; ++j) x = (S.inds[i] == j) ? v[j] : x;
            p[i] = x;
        }
    } else {
#pragma unroll
        for (int i = 0; i < DW; ++i) {
            double acc = dmul(S.C[i], v[0]);
#pragma unroll
            for (int j = 1; j < N; ++j) acc = dadd(acc, dmul(S.C[i + j * DW], v[j]));
            p[i] = acc;
        }
    }
}

}  // namespace mpb
