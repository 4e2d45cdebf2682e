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
// Packed 2-D obstacle table (doubles; small integers stored exactly as doubles):
//   T[0]=n_gates T[1]=n_shapes T[2]=flags T[3]=unused
//   gates  at T[4 + 5g]            : parent xlo xhi ylo yhi
//   shapes at T[4 + 5G + 4s]       : kind gate off K        (off = index into T)
//   Circle  data: cx cy r xlo xhi ylo yhi
//   Polygon data: xlo xhi ylo yhi pts[2K] normals[2K] nextrema[2K]
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

// colliding(p, obstacles): SAT2D.jl:129-132
__device__ __forceinline__ bool point_colliding_2d(const double *T, double px, double py) {
    const int G = (int)T[0], S = (int)T[1];
    const bool fixed = ((int)T[2]) & 1;
    uint32_t pass = 0;
    for (int g = 0; g < G; ++g) {
        const double *a = T + 4 + 5 * g;
        int par = (int)a[0];
        bool ok = (par < 0 || ((pass >> par) & 1u)) && ininterval(px, a[1], a[2]) && ininterval(py, a[3], a[4]);
        pass |= (uint32_t)ok << g;
    }
    const double *dir = T + 4 + 5 * G;
    for (int s = 0; s < S; ++s) {
        int gate = (int)dir[4 * s + 1];
        if (gate >= 0 && !((pass >> gate) & 1u)) continue;
        const double *D = T + (int)dir[4 * s + 2];
        if ((int)dir[4 * s] == 0) {
            if (point_circle(D, px, py)) return true;
        } else {
            if (point_polygon(D, (int)dir[4 * s + 3], px, py, fixed)) return true;
        }
    }
    return false;
}

// colliding(Line(v,w), obstacles): SAT2D.jl:158-161,178-180
__device__ __forceinline__ bool line_colliding_2d(const double *T, double v1, double v2, double w1, double w2) {
    const int G = (int)T[0], S = (int)T[1];
    const bool fixed = ((int)T[2]) & 1;
    Line2 L = make_line(v1, v2, w1, w2);
    uint32_t pass = 0;
    for (int g = 0; g < G; ++g) {
        const double *a = T + 4 + 5 * g;
        int par = (int)a[0];
        bool ok = (par < 0 || ((pass >> par) & 1u)) && overlapping(a[1], a[2], L.xl, L.xh) &&
                  overlapping(a[3], a[4], L.yl, L.yh);
        pass |= (uint32_t)ok << g;
    }
    const double *dir = T + 4 + 5 * G;
    for (int s = 0; s < S; ++s) {
        int gate = (int)dir[4 * s + 1];
        if (gate >= 0 && !((pass >> gate) & 1u)) continue;
        const double *D = T + (int)dir[4 * s + 2];
        if ((int)dir[4 * s] == 0) {
            if (line_circle_ends_free(L, D) || point_circle(D, v1, v2) || point_circle(D, w1, w2)) return true;
        } else {
            int K = (int)dir[4 * s + 3];
            if (line_polygon_ends_free(L, D, K) || point_polygon(D, K, v1, v2, fixed) ||
                point_polygon(D, K, w1, w2, fixed))
                return true;
        }
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
// boxesND.jl:44-56 (quirk Q2 kept: one face per axis, no lambda range test, IEEE division by zero)
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
    for (int k = 0; k < M; ++k) {
        const double *l = lo + k * D, *h = hi + k * D;
        bool broad = false;
#pragma unroll
        for (int i = 0; i < D; ++i) broad = broad || (h[i] < bmin[i] || l[i] > bmax[i]);
        if (broad) continue;
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
        if (hit) return false;
    }
    return true;
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
