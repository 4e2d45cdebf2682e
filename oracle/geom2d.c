/*
 * geom2d.c -- oracle restatement of the reference's 2-D separating-axis
 * collision predicates.  TEST INFRASTRUCTURE ONLY (see mp_oracle.h).
 *
 * Follows src/collisioncheckers/SAT2D.jl:12-185, src/utilities/vec2Dutils.jl:5-36,
 * src/utilities/utils.jl:3-39 (@any/@all) and src/collisioncheckers/robots2D.jl:12-14
 * literally, including the inverted point-in-polygon test (SAT2D.jl:124-127).
 * IEEE double, left-to-right evaluation, no FMA (-ffp-contract=off).
 */
#include "mp_oracle.h"
#include <math.h>
#include <stdlib.h>

/* vec2Dutils.jl:5 (StaticArrays dot on SVector{2}) */
static inline double dot2(double a1, double a2, double b1, double b2) { return a1 * b1 + a2 * b2; }
/* vec2Dutils.jl:7 */
static inline double cross2(double a1, double a2, double b1, double b2) { return a1 * b2 - a2 * b1; }
/* vec2Dutils.jl:34 */
static inline int overlapping(double i1, double i2, double j1, double j2) { return i1 <= j2 && j1 <= i2; }
/* vec2Dutils.jl:35 */
static inline int ininterval(double x, double i1, double i2) { return i1 <= x && x <= i2; }
/* vec2Dutils.jl:36 */
static inline void minmaxV(double x, double y, double *lo, double *hi)
{
    if (x < y) { *lo = x; *hi = y; } else { *lo = y; *hi = x; }
}
/* vec2Dutils.jl:18-27 */
static inline void project_nextrema(const double *pts, int K, double n1, double n2, double *dmin, double *dmax)
{
    double mn = INFINITY, mx = -INFINITY;
    for (int i = 0; i < K; ++i) {
        double d = dot2(pts[2 * i], pts[2 * i + 1], n1, n2);
        if (d < mn) mn = d;
        if (d > mx) mx = d;
    }
    *dmin = mn; *dmax = mx;
}

/* SAT2D.jl:12-25 : Circle(c, r) with its AABB */
int orc_circle_build(double cx, double cy, double r, double *out)
{
    if (r <= 0) return -1; /* SAT2D.jl:19 */
    out[0] = cx; out[1] = cy; out[2] = r;
    out[3] = cx - r; out[4] = cx + r;
    out[5] = cy - r; out[6] = cy + r;
    return 0;
}

/* SAT2D.jl:29-51 : Polygon(points).  out = xlo xhi ylo yhi pts normals nextrema.
 * normalize() is StaticArrays' inv(norm(v))*v (third-party, unpinned: SURVEY 8c). */
int orc_polygon_build(const double *pin, int K, double *out)
{
    if (K < 3) return -1; /* SAT2D.jl:39 */
    double *pts = out + 4, *nrm = out + 4 + 2 * K, *ext = out + 4 + 4 * K;
    double s = 0.0; /* SAT2D.jl:40 : orientation sum, i = 1..N in order */
    for (int i = 0; i < K; ++i) {
        int j = (i + 1 == K) ? 0 : i + 1;
        s += (pin[2 * j] - pin[2 * i]) * (pin[2 * j + 1] + pin[2 * i + 1]);
    }
    for (int i = 0; i < K; ++i) { /* reverse!(points) if clockwise, SAT2D.jl:41 */
        int src = (s > 0) ? (K - 1 - i) : i;
        pts[2 * i] = pin[2 * src]; pts[2 * i + 1] = pin[2 * src + 1];
    }
    double *ang = (double *)malloc(sizeof(double) * (size_t)(K + 1));
    for (int i = 0; i < K; ++i) {
        int j = (i + 1 == K) ? 0 : i + 1;
        double e1 = pts[2 * j] - pts[2 * i], e2 = pts[2 * j + 1] - pts[2 * i + 1]; /* :43 */
        double p1 = e2, p2 = -e1;                                                /* perp, vec2Dutils.jl:6 */
        double inv = 1.0 / sqrt(p1 * p1 + p2 * p2);                              /* :44 normalize */
        nrm[2 * i] = inv * p1; nrm[2 * i + 1] = inv * p2;
        ang[i] = atan2(nrm[2 * i + 1], nrm[2 * i]);
    }
    ang[K] = ang[0];
    int bad = 0; /* SAT2D.jl:45 convexity assertion */
    for (int i = 0; i < K; ++i) {
        double df = ang[i + 1] - ang[i];
        if (-M_PI <= df && df <= 0) bad = 1;
    }
    free(ang);
    if (bad) return -2;
    double xlo = pts[0], xhi = pts[0], ylo = pts[1], yhi = pts[1]; /* :46-47 extrema */
    for (int i = 1; i < K; ++i) {
        if (pts[2 * i] < xlo) xlo = pts[2 * i];
        if (pts[2 * i] > xhi) xhi = pts[2 * i];
        if (pts[2 * i + 1] < ylo) ylo = pts[2 * i + 1];
        if (pts[2 * i + 1] > yhi) yhi = pts[2 * i + 1];
    }
    out[0] = xlo; out[1] = xhi; out[2] = ylo; out[3] = yhi;
    for (int i = 0; i < K; ++i) /* :48 */
        project_nextrema(pts, K, nrm[2 * i], nrm[2 * i + 1], &ext[2 * i], &ext[2 * i + 1]);
    return 0;
}

/* ---- point collisions -------------------------------------------------- */
/* SAT2D.jl:122 */
static int point_circle(const double *c, double px, double py)
{
    double d1 = px - c[0], d2 = py - c[1];
    return dot2(d1, d2, d1, d2) <= c[2] * c[2];
}
/* SAT2D.jl:124-127 (quirk Q1: @all [!ininterval...]); flags bit0 selects the intended test */
static int point_polygon(const double *P, int K, double px, double py, int fixed)
{
    if (!(ininterval(px, P[0], P[1]) && ininterval(py, P[2], P[3]))) return 0;
    const double *nrm = P + 4 + 2 * K, *ext = P + 4 + 4 * K;
    for (int i = 0; i < K; ++i) {
        int in = ininterval(dot2(px, py, nrm[2 * i], nrm[2 * i + 1]), ext[2 * i], ext[2 * i + 1]);
        if (fixed ? !in : in) return 0;
    }
    return 1;
}

/* evaluate the compound gates for a point (SAT2D.jl:129-132) or a line AABB (:158-161) */
static uint64_t gates_point(const orc_obs2d *O, double px, double py)
{
    uint64_t pass = 0;
    for (int g = 0; g < O->n_gates; ++g) {
        const double *a = O->gate_aabb + 4 * g;
        int par = O->gate_parent[g];
        int ok = (par < 0 || ((pass >> par) & 1)) && ininterval(px, a[0], a[1]) && ininterval(py, a[2], a[3]);
        pass |= (uint64_t)ok << g;
    }
    return pass;
}
static uint64_t gates_aabb(const orc_obs2d *O, double xl, double xh, double yl, double yh)
{
    uint64_t pass = 0;
    for (int g = 0; g < O->n_gates; ++g) {
        const double *a = O->gate_aabb + 4 * g;
        int par = O->gate_parent[g];
        /* AABBseparated(C, L) = !(overlapping(C.xrange, L.xrange) && overlapping(C.yrange, L.yrange)), :119 */
        int ok = (par < 0 || ((pass >> par) & 1)) && overlapping(a[0], a[1], xl, xh) && overlapping(a[2], a[3], yl, yh);
        pass |= (uint64_t)ok << g;
    }
    return pass;
}
static inline int gate_ok(uint64_t pass, int g) { return g < 0 || ((pass >> g) & 1); }

int orc_point_colliding_2d(const orc_obs2d *O, double px, double py)
{
    uint64_t pass = gates_point(O, px, py);
    for (int s = 0; s < O->n_shapes; ++s) {
        if (!gate_ok(pass, O->shape_gate[s])) continue;
        const double *D = O->data + O->shape_off[s];
        if (O->shape_kind[s] == 0) {
            if (point_circle(D, px, py)) return 1;
        } else {
            int K = (O->shape_off[s + 1] - O->shape_off[s] - 4) / 6;
            if (point_polygon(D, K, px, py, O->flags & 1)) return 1;
        }
    }
    return 0;
}

/* ---- swept (segment) collisions ---------------------------------------- */
typedef struct { double v1, v2, w1, w2, e1, e2, n1, n2, xl, xh, yl, yh, ndotv; } line_t;
/* SAT2D.jl:69-76 */
static line_t make_line(double v1, double v2, double w1, double w2)
{
    line_t L;
    L.v1 = v1; L.v2 = v2; L.w1 = w1; L.w2 = w2;
    L.e1 = w1 - v1; L.e2 = w2 - v2;
    L.n1 = L.e2; L.n2 = -L.e1; /* perp(edge), not normalised */
    minmaxV(v1, w1, &L.xl, &L.xh);
    minmaxV(v2, w2, &L.yl, &L.yh);
    L.ndotv = dot2(v1, v2, L.n1, L.n2);
    return L;
}
/* SAT2D.jl:165-171 */
static int line_circle_ends_free(const line_t *L, const double *c)
{
    if (!(overlapping(L->xl, L->xh, c[3], c[4]) && overlapping(L->yl, L->yh, c[5], c[6]))) return 0;
    double vc1 = c[0] - L->v1, vc2 = c[1] - L->v2;
    double d2 = dot2(L->e1, L->e2, L->e1, L->e2);
    double cr = cross2(L->e1, L->e2, vc1, vc2);
    if (d2 * (c[2] * c[2]) < cr * cr) return 0;
    double t = dot2(vc1, vc2, L->e1, L->e2);
    return 0 <= t && t <= d2;
}
/* SAT2D.jl:172-176 with is_separating_axis :113-114 */
static int line_polygon_ends_free(const line_t *L, const double *P, int K)
{
    if (!(overlapping(L->xl, L->xh, P[0], P[1]) && overlapping(L->yl, L->yh, P[2], P[3]))) return 0;
    const double *pts = P + 4, *nrm = P + 4 + 2 * K, *ext = P + 4 + 4 * K;
    double mn, mx;
    project_nextrema(pts, K, L->n1, L->n2, &mn, &mx);
    if (!ininterval(L->ndotv, mn, mx)) return 0;
    for (int i = 0; i < K; ++i) {
        double a = dot2(L->v1, L->v2, nrm[2 * i], nrm[2 * i + 1]);
        double b = dot2(L->w1, L->w2, nrm[2 * i], nrm[2 * i + 1]);
        double lo, hi;
        minmaxV(a, b, &lo, &hi);
        if (!overlapping(ext[2 * i], ext[2 * i + 1], lo, hi)) return 0;
    }
    return 1;
}

/* colliding(Line(v,w), obstacles): SAT2D.jl:158-161,178-180 */
int orc_line_colliding_2d(const orc_obs2d *O, double v1, double v2, double w1, double w2)
{
    line_t L = make_line(v1, v2, w1, w2);
    uint64_t pass = gates_aabb(O, L.xl, L.xh, L.yl, L.yh);
    int fixed = O->flags & 1;
    for (int s = 0; s < O->n_shapes; ++s) {
        if (!gate_ok(pass, O->shape_gate[s])) continue;
        const double *D = O->data + O->shape_off[s];
        if (O->shape_kind[s] == 0) {
            if (line_circle_ends_free(&L, D) || point_circle(D, v1, v2) || point_circle(D, w1, w2)) return 1;
        } else {
            int K = (O->shape_off[s + 1] - O->shape_off[s] - 4) / 6;
            if (line_polygon_ends_free(&L, D, K) || point_polygon(D, K, v1, v2, fixed) ||
                point_polygon(D, K, w1, w2, fixed))
                return 1;
        }
    }
    return 0;
}

/* robots2D.jl:12 */
void orc_points_free_2d(const orc_obs2d *O, const double *P, int64_t n, uint8_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = !orc_point_colliding_2d(O, P[2 * i], P[2 * i + 1]);
}
/* robots2D.jl:13-14 */
void orc_segments_free_2d(const orc_obs2d *O, const double *V, const double *W, int64_t n, uint8_t *out)
{
    for (int64_t i = 0; i < n; ++i)
        out[i] = !orc_line_colliding_2d(O, V[2 * i], V[2 * i + 1], W[2 * i], W[2 * i + 1]);
}
