/*
 * cars.c -- CPU oracle for the chopped-metric car spaces (TEST INFRASTRUCTURE ONLY, see mp_oracle.h).
 *
 * Restates src/statespaces/simplecars.jl of the reference in source operation order:
 *   dubins / dubinsLSL! .. dubinsLRL!                      simplecars.jl:102-213
 *   reedsshepp / LpSpLp! .. LpRmSmLmRp! / Tau, Omega, M, R simplecars.jl:230-523
 *   propagate, collision_waypoints for one StepControl      simplecars.jl:55-82
 *   the chopped evaluation                                  primitivetypes.jl:95-100
 *   inball(V, ::ChoppedPreMetric, ::TreeDistanceDS, ...)    nearneighbors.jl:185-198, simplecars.jl:42-52
 *   collision_waypoints over a control sequence + push!(w)  statespaces.jl:134-142
 *   is_free_motion(v, w, CC, SS)                            statespaces.jl:153-158
 *
 * PARITY UNPINNED for the elementary functions: the reference calls openlibm's sin / cos / atan2 / acos, which
 * are not correctly rounded and cannot be run here.  Oracle and GPU share the SPECIFICATION below instead
 * (orc_det_sincos, orc_det_atan2, orc_det_acos: argument reduction + fixed polynomials in basic IEEE operations,
 * each within a few ulp of the true function -- tests/test_oracle_cars.py measures it against libm), so that the
 * two agree bit for bit; sqrt and mod are exact on both sides (mod2piF = Julia's mod(x, 2pi): fmod + sign fix).
 * The algebra around them is pinned by golden vectors (the reference's formulas re-run in 40-digit mpmath arithmetic,
 * tests/golden/gen_cars_golden.py -> cars.json), an independent geometric Dubins construction, closed-form known
 * answers and consistency (the returned control, propagated from v, must arrive at w; Reeds-Shepp <= Dubins;
 * symmetry): tests/test_oracle_cars.py.
 */
#include "mp_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double PI = 3.141592653589793;      /* Float64(pi) */
static const double TWO_PI = 6.283185307179586;  /* 2*Float64(pi) */

/* ---- elementary functions (specification shared with csrc/cars.cu) ------------------------------------ */
/* mod(x, 2pi) as Julia defines it for floats: r = rem(x, y) exactly; r == 0 -> +0; sign(r) != sign(y) -> r + y */
double orc_mod2pi(double x)
{
    double r = fmod(x, TWO_PI);
    if (r == 0.0) return 0.0;
    if (r < 0.0) return r + TWO_PI;
    return r;
}

/* sin and cos: Cody-Waite reduction by pi/2 in three parts (|x| < 2^20), Taylor polynomials on [-pi/4, pi/4];
 * sin(-x) == -sin(x) and cos(-x) == cos(x) bit for bit (the device code relies on it to share evaluations) */
void orc_det_sincos(double x, double *sn, double *cs)
{
    /* computed on |x| and reflected, so that sin is exactly odd and cos exactly even */
    double ax = fabs(x);
    double kf = floor(ax * 0.6366197723675814 + 0.5);
    double th = ((ax - kf * 1.5707963267341256) - kf * 6.077100506303966e-11) - kf * 2.0222662487111665e-21;
    double t2 = th * th;
    double ps = -1.0 / 355687428096000.0; /* -1/17! */
    ps = ps * t2 + 1.0 / 1307674368000.0;
    ps = ps * t2 - 1.0 / 6227020800.0;
    ps = ps * t2 + 1.0 / 39916800.0;
    ps = ps * t2 - 1.0 / 362880.0;
    ps = ps * t2 + 1.0 / 5040.0;
    ps = ps * t2 - 1.0 / 120.0;
    ps = ps * t2 + 1.0 / 6.0;
    double s = th - th * t2 * ps;
    double pc = 1.0 / 6402373705728000.0; /* 1/18! */
    pc = pc * t2 - 1.0 / 20922789888000.0;
    pc = pc * t2 + 1.0 / 87178291200.0;
    pc = pc * t2 - 1.0 / 479001600.0;
    pc = pc * t2 + 1.0 / 3628800.0;
    pc = pc * t2 - 1.0 / 40320.0;
    pc = pc * t2 + 1.0 / 720.0;
    pc = pc * t2 - 1.0 / 24.0;
    pc = pc * t2 + 0.5;
    double c = 1.0 - t2 * pc;
    int q = (int)((long long)kf & 3);
    if (q == 0) { *sn = s; *cs = c; }
    else if (q == 1) { *sn = c; *cs = -s; }
    else if (q == 2) { *sn = -s; *cs = -c; }
    else { *sn = -c; *cs = s; }
    if (x < 0.0) *sn = -*sn;
}
double orc_det_sin(double x) { double s, c; orc_det_sincos(x, &s, &c); return s; }
double orc_det_cos(double x) { double s, c; orc_det_sincos(x, &s, &c); return c; }

/* atan on [0, 1]: breakpoint c = round(4a)/4, t = (a - c)/(1 + a c) in [-1/8, 1/8], odd Taylor series to t^21 */
static double det_atan01(double a)
{
    static const double atan_c[5] = {0.0, 0.24497866312686414, 0.4636476090008061, 0.6435011087932844,
                                     0.7853981633974483};
    int ci = (int)(a * 4.0 + 0.5);
    double c = 0.25 * (double)ci;
    double t = (a - c) / (1.0 + a * c);
    double t2 = t * t;
    double p = 1.0 / 21.0;
    p = 1.0 / 19.0 - p * t2;
    p = 1.0 / 17.0 - p * t2;
    p = 1.0 / 15.0 - p * t2;
    p = 1.0 / 13.0 - p * t2;
    p = 1.0 / 11.0 - p * t2;
    p = 1.0 / 9.0 - p * t2;
    p = 1.0 / 7.0 - p * t2;
    p = 1.0 / 5.0 - p * t2;
    p = 1.0 / 3.0 - p * t2;
    p = 1.0 - p * t2;
    return atan_c[ci] + t * p;
}
double orc_det_atan2(double y, double x)
{
    double ax = fabs(x), ay = fabs(y), ang;
    if (ax == 0.0 && ay == 0.0) ang = 0.0;
    else if (ay <= ax) ang = det_atan01(ay / ax);
    else ang = 1.5707963267948966 - det_atan01(ax / ay);
    if (signbit(x)) ang = PI - ang;
    return signbit(y) ? -ang : ang;
}
/* acos through atan2; a radicand that rounding made negative is taken as 0 (the reference's acos would throw) */
double orc_det_acos(double x)
{
    double rad = (1.0 - x) * (1.0 + x);
    if (rad < 0.0) rad = 0.0;
    return orc_det_atan2(sqrt(rad), x);
}

/* ---- segments: (duration, signed speed, signed curvature) = StepControl(t, (u1, u2)) ------------------- */
typedef struct { double t, u1, u2; } seg;

/* carsegment2stepcontrol(t::Int, d) = StepControl(abs(d), (sign(d), t))   simplecars.jl:85 */
static seg mkseg(int turn, double d)
{
    seg s;
    s.t = fabs(d);
    s.u1 = d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0);
    s.u2 = (double)turn;
    return s;
}

/* ---- Dubins (simplecars.jl:102-213) ---------------------------------------------------------------------- */
static double dubinsLSL(double d, double a, double b, double c, seg *path)
{
    double ca, sa, cb, sb;
    orc_det_sincos(a, &sa, &ca);
    orc_det_sincos(b, &sb, &cb);
    double tmp = 2.0 + d * d - 2.0 * (ca * cb + sa * sb - d * (sa - sb));
    if (tmp < 0.0) return c;
    double th = orc_det_atan2(cb - ca, d + sa - sb);
    double t = orc_mod2pi(-a + th);
    double p = sqrt(tmp > 0.0 ? tmp : 0.0);
    double q = orc_mod2pi(b - th);
    double cnew = t + p + q;
    if (c <= cnew) return c;
    path[0] = mkseg(1, t); path[1] = mkseg(0, p); path[2] = mkseg(1, q);
    return cnew;
}
static double dubinsRSR(double d, double a, double b, double c, seg *path)
{
    double ca, sa, cb, sb;
    orc_det_sincos(a, &sa, &ca);
    orc_det_sincos(b, &sb, &cb);
    double tmp = 2.0 + d * d - 2.0 * (ca * cb + sa * sb - d * (sb - sa));
    if (tmp < 0.0) return c;
    double th = orc_det_atan2(ca - cb, d - sa + sb);
    double t = orc_mod2pi(a - th);
    double p = sqrt(tmp > 0.0 ? tmp : 0.0);
    double q = orc_mod2pi(-b + th);
    double cnew = t + p + q;
    if (c <= cnew) return c;
    path[0] = mkseg(-1, t); path[1] = mkseg(0, p); path[2] = mkseg(-1, q);
    return cnew;
}
static double dubinsRSL(double d, double a, double b, double c, seg *path)
{
    double ca, sa, cb, sb;
    orc_det_sincos(a, &sa, &ca);
    orc_det_sincos(b, &sb, &cb);
    double tmp = d * d - 2.0 + 2.0 * (ca * cb + sa * sb - d * (sa + sb));
    if (tmp < 0.0) return c;
    double p = sqrt(tmp > 0.0 ? tmp : 0.0);
    double th = orc_det_atan2(ca + cb, d - sa - sb) - orc_det_atan2(2.0, p);
    double t = orc_mod2pi(a - th);
    double q = orc_mod2pi(b - th);
    double cnew = t + p + q;
    if (c <= cnew) return c;
    path[0] = mkseg(-1, t); path[1] = mkseg(0, p); path[2] = mkseg(1, q);
    return cnew;
}
static double dubinsLSR(double d, double a, double b, double c, seg *path)
{
    double ca, sa, cb, sb;
    orc_det_sincos(a, &sa, &ca);
    orc_det_sincos(b, &sb, &cb);
    double tmp = -2.0 + d * d + 2.0 * (ca * cb + sa * sb + d * (sa + sb));
    if (tmp < 0.0) return c;
    double p = sqrt(tmp > 0.0 ? tmp : 0.0);
    double th = orc_det_atan2(-ca - cb, d + sa + sb) - orc_det_atan2(-2.0, p);
    double t = orc_mod2pi(-a + th);
    double q = orc_mod2pi(-b + th);
    double cnew = t + p + q;
    if (c <= cnew) return c;
    path[0] = mkseg(1, t); path[1] = mkseg(0, p); path[2] = mkseg(-1, q);
    return cnew;
}
static double dubinsRLR(double d, double a, double b, double c, seg *path)
{
    double ca, sa, cb, sb;
    orc_det_sincos(a, &sa, &ca);
    orc_det_sincos(b, &sb, &cb);
    double tmp = (6.0 - d * d + 2.0 * (ca * cb + sa * sb + d * (sa - sb))) / 8.0;
    if (fabs(tmp) >= 1.0) return c;
    double p = TWO_PI - orc_det_acos(tmp);
    double th = orc_det_atan2(ca - cb, d - sa + sb);
    double t = orc_mod2pi(a - th + p / 2.0);
    double q = orc_mod2pi(a - b - t + p);
    double cnew = t + p + q;
    if (c <= cnew) return c;
    path[0] = mkseg(-1, t); path[1] = mkseg(1, p); path[2] = mkseg(-1, q);
    return cnew;
}
static double dubinsLRL(double d, double a, double b, double c, seg *path)
{
    double ca, sa, cb, sb;
    orc_det_sincos(a, &sa, &ca);
    orc_det_sincos(b, &sb, &cb);
    double tmp = (6.0 - d * d + 2.0 * (ca * cb + sa * sb - d * (sa - sb))) / 8.0;
    if (fabs(tmp) >= 1.0) return c;
    double p = TWO_PI - orc_det_acos(tmp);
    double th = orc_det_atan2(-ca + cb, d + sa - sb);
    double t = orc_mod2pi(-a + th + p / 2.0);
    double q = orc_mod2pi(b - a - t + p);
    double cnew = t + p + q;
    if (c <= cnew) return c;
    path[0] = mkseg(1, t); path[1] = mkseg(-1, p); path[2] = mkseg(1, q);
    return cnew;
}

/* scalespeed!(scaleradius!(p, r), s)   simplecars.jl:86-99 */
static void scale_segments(seg *p, int l, double r, double s)
{
    for (int i = 0; i < l; ++i) {
        p[i].t = p[i].t * r; p[i].u2 = p[i].u2 / r;
        p[i].t = p[i].t / s; p[i].u1 = p[i].u1 * s;
    }
}

/* dubins(s1, s2, r, s) -> cost, 3 segments   simplecars.jl:196-213 */
static double dubins(const double *s1, const double *s2, double r, double s, seg *pmin)
{
    double vx = (s2[0] - s1[0]) / r, vy = (s2[1] - s1[1]) / r;
    double d = sqrt(vx * vx + vy * vy);
    double th = orc_det_atan2(vy, vx);
    double a = orc_mod2pi(s1[2] - th);
    double b = orc_mod2pi(s2[2] - th);
    double cmin = INFINITY;
    for (int i = 0; i < 3; ++i) pmin[i] = mkseg(0, 0.0);
    cmin = dubinsLSL(d, a, b, cmin, pmin);
    cmin = dubinsRSR(d, a, b, cmin, pmin);
    cmin = dubinsRSL(d, a, b, cmin, pmin);
    cmin = dubinsLSR(d, a, b, cmin, pmin);
    cmin = dubinsRLR(d, a, b, cmin, pmin);
    cmin = dubinsLRL(d, a, b, cmin, pmin);
    scale_segments(pmin, 3, r, s);
    return cmin * r;
}

/* ---- Reeds-Shepp (simplecars.jl:230-523) ----------------------------------------------------------------- */
static void Rpolar(double x, double y, double *r, double *th) { *r = sqrt(x * x + y * y); *th = orc_det_atan2(y, x); }
static double Mwrap(double t) { double m = orc_mod2pi(t); return m > PI ? m - TWO_PI : m; }
static double Tau(double u, double v, double E, double N)
{
    double delta = Mwrap(u - v);
    double A = orc_det_sin(u) - orc_det_sin(delta);
    double B = orc_det_cos(u) - orc_det_cos(delta) - 1.0;
    double r, th;
    Rpolar(E * A + N * B, N * A - E * B, &r, &th);
    double t = 2.0 * orc_det_cos(delta) - 2.0 * orc_det_cos(v) - 2.0 * orc_det_cos(u) + 3.0;
    return t < 0.0 ? Mwrap(th + PI) : Mwrap(th);
}
static double Omega(double u, double v, double E, double N, double t) { return Mwrap(Tau(u, v, E, N) - u + v - t); }

typedef struct { double c; int l; seg p[5]; } rs_best;

static int LpSpLp(double tx, double ty, double tt, rs_best *b)
{
    double r, th;
    Rpolar(tx - orc_det_sin(tt), ty - 1.0 + orc_det_cos(tt), &r, &th);
    double u = r, t = orc_mod2pi(th), v = orc_mod2pi(tt - t);
    double cnew = t + u + v;
    if (b->c <= cnew) return 0;
    b->p[0] = mkseg(1, t); b->p[1] = mkseg(0, u); b->p[2] = mkseg(1, v);
    b->c = cnew; b->l = 3;
    return 1;
}
static int LpSpRp(double tx, double ty, double tt, rs_best *b)
{
    double r, th, r1, th1;
    Rpolar(tx + orc_det_sin(tt), ty - 1.0 - orc_det_cos(tt), &r, &th);
    if (r * r < 4.0) return 0;
    double u = sqrt(r * r - 4.0);
    Rpolar(u, 2.0, &r1, &th1);
    double t = orc_mod2pi(th + th1), v = orc_mod2pi(t - tt);
    double cnew = t + u + v;
    if (b->c <= cnew) return 0;
    b->p[0] = mkseg(1, t); b->p[1] = mkseg(0, u); b->p[2] = mkseg(-1, v);
    b->c = cnew; b->l = 3;
    return 1;
}
static int LpRmLp(double tx, double ty, double tt, rs_best *b)
{
    double E = tx - orc_det_sin(tt), N = ty + orc_det_cos(tt) - 1.0;
    if (E * E + N * N > 16.0) return 0;
    double r, th;
    Rpolar(E, N, &r, &th);
    double u = orc_det_acos(1.0 - r * r / 8.0);
    double t = orc_mod2pi(th - u / 2.0 + PI);
    double v = orc_mod2pi(PI - u / 2.0 - th + tt);
    u = -u;
    double cnew = t - u + v;
    if (b->c <= cnew) return 0;
    b->p[0] = mkseg(1, t); b->p[1] = mkseg(-1, u); b->p[2] = mkseg(1, v);
    b->c = cnew; b->l = 3;
    return 1;
}
static int LpRmLm(double tx, double ty, double tt, rs_best *b)
{
    double E = tx - orc_det_sin(tt), N = ty + orc_det_cos(tt) - 1.0;
    if (E * E + N * N > 16.0) return 0;
    double r, th;
    Rpolar(E, N, &r, &th);
    double u = orc_det_acos(1.0 - r * r / 8.0);
    double t = orc_mod2pi(th - u / 2.0 + PI);
    double v = orc_mod2pi(PI - u / 2.0 - th + tt) - TWO_PI;
    u = -u;
    double cnew = t - u - v;
    if (b->c <= cnew) return 0;
    b->p[0] = mkseg(1, t); b->p[1] = mkseg(-1, u); b->p[2] = mkseg(1, v);
    b->c = cnew; b->l = 3;
    return 1;
}
static int LpRpuLmuRm(double tx, double ty, double tt, rs_best *b)
{
    double E = tx + orc_det_sin(tt), N = ty - orc_det_cos(tt) - 1.0;
    double p = (2.0 + sqrt(E * E + N * N)) / 4.0;
    if (p < 0.0 || p > 1.0) return 0;
    double u = orc_det_acos(p);
    double t = orc_mod2pi(Tau(u, -u, E, N));
    double v = orc_mod2pi(Omega(u, -u, E, N, tt)) - TWO_PI;
    double cnew = t + 2.0 * u - v;
    if (b->c <= cnew) return 0;
    b->p[0] = mkseg(1, t); b->p[1] = mkseg(-1, u); b->p[2] = mkseg(1, -u); b->p[3] = mkseg(-1, v);
    b->c = cnew; b->l = 4;
    return 1;
}
static int LpRmuLmuRp(double tx, double ty, double tt, rs_best *b)
{
    double E = tx + orc_det_sin(tt), N = ty - orc_det_cos(tt) - 1.0;
    double p = (20.0 - E * E - N * N) / 16.0;
    if (p < 0.0 || p > 1.0) return 0;
    double u = -orc_det_acos(p);
    double t = orc_mod2pi(Tau(u, u, E, N));
    double v = orc_mod2pi(Omega(u, u, E, N, tt));
    double cnew = t - 2.0 * u + v;
    if (b->c <= cnew) return 0;
    b->p[0] = mkseg(1, t); b->p[1] = mkseg(-1, u); b->p[2] = mkseg(1, u); b->p[3] = mkseg(-1, v);
    b->c = cnew; b->l = 4;
    return 1;
}
static int LpRmSmLm(double tx, double ty, double tt, rs_best *b)
{
    double E = tx - orc_det_sin(tt), N = ty + orc_det_cos(tt) - 1.0;
    double D, beta;
    Rpolar(E, N, &D, &beta);
    if (D < 2.0) return 0;
    double gamma = orc_det_acos(2.0 / D);
    double F = sqrt(D * D / 4.0 - 1.0);
    double t = orc_mod2pi(PI + beta - gamma);
    double u = 2.0 - 2.0 * F;
    if (u > 0.0) return 0;
    double v = orc_mod2pi(-3.0 * PI / 2.0 + gamma + tt - beta) - TWO_PI;
    double cnew = t + PI / 2.0 - u - v;
    if (b->c <= cnew) return 0;
    b->p[0] = mkseg(1, t); b->p[1] = mkseg(-1, -PI / 2.0); b->p[2] = mkseg(0, u); b->p[3] = mkseg(1, v);
    b->c = cnew; b->l = 4;
    return 1;
}
static int LpRmSmRm(double tx, double ty, double tt, rs_best *b)
{
    double E = tx + orc_det_sin(tt), N = ty - orc_det_cos(tt) - 1.0;
    double D, beta;
    Rpolar(E, N, &D, &beta);
    if (D < 2.0) return 0;
    double t = orc_mod2pi(beta + PI / 2.0);
    double u = 2.0 - D;
    if (u > 0.0) return 0;
    double v = orc_mod2pi(-PI - tt + beta) - TWO_PI;
    double cnew = t + PI / 2.0 - u - v;
    if (b->c <= cnew) return 0;
    b->p[0] = mkseg(1, t); b->p[1] = mkseg(-1, -PI / 2.0); b->p[2] = mkseg(0, u); b->p[3] = mkseg(-1, v);
    b->c = cnew; b->l = 4;
    return 1;
}
static int LpRmSmLmRp(double tx, double ty, double tt, rs_best *b)
{
    double E = tx + orc_det_sin(tt), N = ty - orc_det_cos(tt) - 1.0;
    double D, beta;
    Rpolar(E, N, &D, &beta);
    if (D < 2.0) return 0;
    double gamma = orc_det_acos(2.0 / D);
    double F = sqrt(D * D / 4.0 - 1.0);
    double t = orc_mod2pi(PI + beta - gamma);
    double u = 4.0 - 2.0 * F;
    if (u > 0.0) return 0;
    double v = orc_mod2pi(PI + beta - tt - gamma);
    double cnew = t + PI - u + v;
    if (b->c <= cnew) return 0;
    b->p[0] = mkseg(1, t); b->p[1] = mkseg(-1, -PI / 2.0); b->p[2] = mkseg(0, u); b->p[3] = mkseg(1, -PI / 2.0);
    b->p[4] = mkseg(-1, v);
    b->c = cnew; b->l = 5;
    return 1;
}

typedef int (*rs_fn)(double, double, double, rs_best *);
/* the sweep of simplecars.jl:283-339: family, then which of the 8 transformed targets (bit0 timeflip, bit1 reflect,
 * bit2 backwards) in the reference's order */
static const struct { rs_fn f; int nt; int tr[8]; } rs_plan[9] = {
    {LpSpLp, 4, {0, 1, 2, 3}},          {LpSpRp, 4, {0, 1, 2, 3}},     {LpRmLp, 2, {0, 2}},
    {LpRmLm, 8, {0, 1, 2, 3, 4, 5, 6, 7}}, {LpRpuLmuRm, 4, {0, 1, 2, 3}}, {LpRmuLmuRp, 4, {0, 1, 2, 3}},
    {LpRmSmLm, 8, {0, 1, 2, 3, 4, 5, 6, 7}}, {LpRmSmRm, 8, {0, 1, 2, 3, 4, 5, 6, 7}}, {LpRmSmLmRp, 4, {0, 1, 2, 3}},
};

/* reedsshepp(s1, s2, r, s) -> cost, *l segments   simplecars.jl:266-364 */
static double reedsshepp(const double *s1, const double *s2, double r, double s, seg *out, int *l_out)
{
    double dx = (s2[0] - s1[0]) / r, dy = (s2[1] - s1[1]) / r;
    double ct, st;
    orc_det_sincos(s1[2], &st, &ct);
    double T[8][3];
    T[0][0] = dx * ct + dy * st; T[0][1] = -dx * st + dy * ct; T[0][2] = orc_mod2pi(s2[2] - s1[2]);
    /* timeflip (x,y,th) -> (-x, y, -th); reflect -> (x, -y, -th); backwards -> (x cos + y sin, x sin - y cos, th) */
    T[1][0] = -T[0][0]; T[1][1] = T[0][1]; T[1][2] = -T[0][2];                 /* t   */
    T[2][0] = T[0][0]; T[2][1] = -T[0][1]; T[2][2] = -T[0][2];                 /* r   */
    T[3][0] = T[1][0]; T[3][1] = -T[1][1]; T[3][2] = -T[1][2];                 /* tr = reflect(timeflip) */
    {
        double sb, cb;
        orc_det_sincos(T[0][2], &sb, &cb);
        T[4][0] = T[0][0] * cb + T[0][1] * sb; T[4][1] = T[0][0] * sb - T[0][1] * cb; T[4][2] = T[0][2];  /* b */
    }
    T[5][0] = -T[4][0]; T[5][1] = T[4][1]; T[5][2] = -T[4][2];                 /* bt  */
    T[6][0] = T[4][0]; T[6][1] = -T[4][1]; T[6][2] = -T[4][2];                 /* br  */
    T[7][0] = T[5][0]; T[7][1] = -T[5][1]; T[7][2] = -T[5][2];                 /* btr */
    rs_best best;
    best.c = INFINITY; best.l = 0;
    int post = 0;
    for (int f = 0; f < 9; ++f)
        for (int k = 0; k < rs_plan[f].nt; ++k) {
            int tr = rs_plan[f].tr[k];
            if (rs_plan[f].f(T[tr][0], T[tr][1], T[tr][2], &best)) post = tr;
        }
    int l = best.l;
    scale_segments(best.p, l, r, s);
    if (post & 1) for (int i = 0; i < l; ++i) best.p[i].u1 = -best.p[i].u1;   /* timeflip! */
    if (post & 2) for (int i = 0; i < l; ++i) best.p[i].u2 = -best.p[i].u2;   /* reflect!  */
    for (int i = 0; i < l; ++i) out[i] = (post & 4) ? best.p[l - 1 - i] : best.p[i]; /* backwards! = reverse! */
    *l_out = l;
    return best.c * r;
}

/* kind: 0 = ReedsSheppExact, 1 = DubinsExact.  segs: 5 x (t, u1, u2) */
double orc_car_steer(int kind, double rturn, double speed, const double *v, const double *w, int *nseg, double *segs)
{
    seg p[5];
    int l = 3;
    double c = kind == 1 ? dubins(v, w, rturn, speed, p) : reedsshepp(v, w, rturn, speed, p, &l);
    if (nseg) *nseg = l;
    if (segs)
        for (int i = 0; i < 5; ++i) {
            segs[3 * i] = i < l ? p[i].t : 0.0; segs[3 * i + 1] = i < l ? p[i].u1 : 0.0; segs[3 * i + 2] = i < l ? p[i].u2 : 0.0;
        }
    return c;
}

/* propagate(M, v, u)   simplecars.jl:55-68 */
void orc_car_propagate(const double *v, const double *u3, double *out)
{
    double t = u3[0], s = u3[1], invr = u3[2];
    double dth = t * s * invr;
    if (fabs(dth) > 10.0 * 2.220446049250313e-16) {
        out[0] = v[0] + (orc_det_sin(v[2] + dth) - orc_det_sin(v[2])) / invr;
        out[1] = v[1] + (orc_det_cos(v[2]) - orc_det_cos(v[2] + dth)) / invr;
    } else {
        out[0] = v[0] + t * s * orc_det_cos(v[2]);
        out[1] = v[1] + t * s * orc_det_sin(v[2]);
    }
    out[2] = orc_mod2pi(v[2] + dth);
}

/* evaluate(::ChoppedPreMetric, v, w)   primitivetypes.jl:95-100 (lowerbound = Euclidean on (x, y), simplecars.jl:49) */
double orc_car_chopped(int kind, double rturn, double chopval, const double *v, const double *w)
{
    double dx = v[0] - w[0], dy = v[1] - w[1];
    double lb = sqrt(dx * dx + dy * dy);
    if (lb > chopval) return INFINITY;
    double d = orc_car_steer(kind, rturn, 1.0, v, w, NULL, NULL);
    return d <= chopval ? d : INFINITY;
}

/* inball(V, dist::ChoppedPreMetric, DS::TreeDistanceDS, v, r, forwards)   nearneighbors.jl:185-198: candidates =
 * inrange of the KD-tree over (x, y) (predicate  dx^2 + dy^2 <= r*r), then the chopped distance, kept when <= r.
 * Two-call protocol as orc_rball_brute (rowval == NULL: counts only). */
void orc_car_inball(const double *V, int64_t N, int kind, double rturn, double r, double chopval, int forwards,
                    int64_t q0, int64_t q1, int64_t *colptr, int64_t *rowval, double *nzval)
{
    int64_t at = 0;
    colptr[0] = 1;
    for (int64_t q = q0; q < q1; ++q) {
        const double *v = V + 3 * q;
        for (int64_t i = 0; i < N; ++i) {
            if (i == q) continue;
            const double *w = V + 3 * i;
            double t = v[0] - w[0];
            double s = t * t;
            t = v[1] - w[1];
            s = s + t * t;
            if (!(s <= r * r)) continue;
            double d = forwards ? orc_car_chopped(kind, rturn, chopval, v, w) : orc_car_chopped(kind, rturn, chopval, w, v);
            if (d <= r) {
                if (rowval) { rowval[at] = i + 1; nzval[at] = d; }
                ++at;
            }
        }
        colptr[q - q0 + 1] = at + 1;
    }
}

static int ws_segment_free(const orc_checker *CC, const orc_space *S, const double *a, const double *b)
{
    double p[2], q[2];
    orc_state2workspace(S, a, p);
    orc_state2workspace(S, b, q);
    if (CC->kind == 0) return !orc_line_colliding_2d(CC->obs2d, p[0], p[1], q[0], q[1]);
    return orc_box_segment_free(CC->box_lo, CC->box_hi, CC->M, CC->d, p, q);
}

/* is_free_motion(v, w, CC, SS)   statespaces.jl:153-158 over collision_waypoints(SS, v, w) (statespaces.jl:134-142:
 * per StepControl the arc points of simplecars.jl:71-82, then propagate; finally push!(w)).  *count += segment tests
 * actually run (the @all short-circuits). */
int orc_car_is_free_motion(const orc_checker *CC, const orc_space *S, int kind, double rturn, double speed,
                           const double *v, const double *w, int64_t *count)
{
    double segs[15];
    int l;
    orc_car_steer(kind, rturn, speed, v, w, &l, segs);
    const double thres = PI / 12.0;
    double cur[3] = {v[0], v[1], v[2]}, prev[3];
    int have_prev = 0;
    /* walk the waypoint list, testing (prev, next) as soon as next is known */
    for (int k = 0; k <= l; ++k) {
        /* waypoints of segment k: cur, then arc points; k == l: the final target w */
        double pts[32][3];
        int np = 0;
        if (k < l) {
            const double *u = segs + 3 * k;
            double s = u[1], invr = u[2];
            pts[np][0] = cur[0]; pts[np][1] = cur[1]; pts[np][2] = cur[2]; ++np;
            long long m = (long long)floor(u[0] * s * invr / thres);
            for (long long i = 1; i <= m && np < 32; ++i) {
                double ang = (double)i * thres;
                pts[np][0] = cur[0] + (orc_det_sin(cur[2] + ang) - orc_det_sin(cur[2])) / invr;
                pts[np][1] = cur[1] + (orc_det_cos(cur[2]) - orc_det_cos(cur[2] + ang)) / invr;
                pts[np][2] = orc_mod2pi(cur[2] + ang);
                ++np;
            }
        } else {
            pts[np][0] = w[0]; pts[np][1] = w[1]; pts[np][2] = w[2]; ++np;
        }
        for (int j = 0; j < np; ++j) {
            if (have_prev) {
                if (!orc_in_state_space(S, prev)) return 0;
                if (count) *count += 1;
                if (!ws_segment_free(CC, S, prev, pts[j])) return 0;
            }
            prev[0] = pts[j][0]; prev[1] = pts[j][1]; prev[2] = pts[j][2];
            have_prev = 1;
        }
        if (k < l) {
            double nxt[3];
            orc_car_propagate(cur, segs + 3 * k, nxt);
            cur[0] = nxt[0]; cur[1] = nxt[1]; cur[2] = nxt[2];
        }
    }
    return 1;
}

/* entry (row y, column x) of a table <-> is_free_motion(V[y], V[x], CC, SS)  (fmt.jl:75) */
void orc_car_edges_free_csc(const orc_checker *CC, const orc_space *S, int kind, double rturn, double speed,
                            const double *V, const int64_t *colptr, const int64_t *rowval, int64_t c0, int64_t c1,
                            uint8_t *out, int64_t *count)
{
    for (int64_t c = c0; c < c1; ++c)
        for (int64_t e = colptr[c - c0] - 1; e < colptr[c - c0 + 1] - 1; ++e)
            out[e] = (uint8_t)orc_car_is_free_motion(CC, S, kind, rturn, speed, V + 3 * (rowval[e] - 1), V + 3 * c, count);
}
