/*
 * lq.c -- oracle restatement of the linear-quadratic steering cost ("ControlNN") for
 * double-integrator systems.  TEST INFRASTRUCTURE ONLY (see mp_oracle.h).
 *
 * Follows src/statespaces/linearquadratic.jl:
 *   :46-53   DoubleIntegrator(d): A=[0 I;0 0], B=[0;I], c=0, R=rho*I, C=[I 0]
 *   :126-157 LinearQuadratic2BVP: cost(x,y,t) = t + (y-xbar(t))' Ginv(t) (y-xbar(t)), its t
 *            derivatives and the optimal state x(x,y,t,s)
 *   :175-190 topt_newton (safeguarded Newton, tol 1e-6), :191-195 steer
 *   :196-225 steer_pairwise + :68-77 helper_data_structures (all-pairs table; DSB = Dmat,
 *            DSF = Dmat'), nearneighbors.jl:165-177 (final filter: index != v && cost <= r)
 *   :85-88   collision_waypoints: 5 states at s = linspace(0, t*, 5); statespaces.jl:153-158
 *
 * The reference obtains cost/dcost/ddcost/x as SymPy-printed closures (third-party, unpinned
 * operation order: SURVEY 8c).  For the double integrator they reduce to
 *     cost(t) = t + alpha/t^3 - beta/t^2 + gamma/t
 *     alpha = 12 dp'R dp, beta = 12 dp'R(v0+v1), gamma = 4 (v0'Rv0 + v0'Rv1 + v1'Rv1)
 * and x(s) is the cubic Hermite curve; tests/golden/ holds vectors produced by re-running the
 * reference's SymPy construction (tests/golden/gen_lq_golden.py), which pin these formulas to
 * ~1e-12 relative.  Operation order below is THE specification shared with the CUDA kernel.
 */
#include "mp_oracle.h"
#include <math.h>
#include <stdlib.h>

#define MAXD 8

typedef struct { double alpha, beta, gamma; } abg_t;

/* R is d x d row-major (symmetric); x = (p[0..d), v[0..d)) */
static abg_t lq_abg(int d, const double *R, const double *x0, const double *x1)
{
    double dp[MAXD], sv[MAXD];
    const double *v0 = x0 + d, *v1 = x1 + d;
    for (int i = 0; i < d; ++i) { dp[i] = x1[i] - x0[i]; sv[i] = v0[i] + v1[i]; }
    double a = 0, b = 0, g = 0;
    for (int i = 0; i < d; ++i) {
        double Rdp = 0, Rv0 = 0, Rv1 = 0; /* (R dp)_i, (R v0)_i, (R v1)_i : j ascending */
        for (int j = 0; j < d; ++j) {
            Rdp = Rdp + R[i * d + j] * dp[j];
            Rv0 = Rv0 + R[i * d + j] * v0[j];
            Rv1 = Rv1 + R[i * d + j] * v1[j];
        }
        a = a + dp[i] * Rdp;
        b = b + sv[i] * Rdp;
        g = g + (v0[i] * Rv0 + v0[i] * Rv1 + v1[i] * Rv1);
    }
    abg_t r = { 12.0 * a, 12.0 * b, 4.0 * g };
    return r;
}
static inline double lq_cost(abg_t k, double t)
{
    double it = 1.0 / t, it2 = it * it, it3 = it2 * it;
    return t + ((k.alpha * it3 - k.beta * it2) + k.gamma * it);
}
static inline double lq_dcost(abg_t k, double t)
{
    double it = 1.0 / t, it2 = it * it, it3 = it2 * it;
    return 1.0 + ((2.0 * k.beta * it3 - 3.0 * k.alpha * (it3 * it)) - k.gamma * it2);
}
static inline double lq_ddcost(abg_t k, double t)
{
    double it = 1.0 / t, it2 = it * it, it3 = it2 * it;
    return (12.0 * k.alpha * (it3 * it2) - 6.0 * k.beta * (it2 * it2)) + 2.0 * k.gamma * it3;
}

#define LQ_MAX_NEWTON 200
/* linearquadratic.jl:175-190.  *capped is set if the safety cap on iterations was hit (never in tests). */
static double lq_topt_newton(abg_t k, double tm, int *capped)
{
    const double tol = 1e-6;
    double b = tm;
    if (lq_dcost(k, b) < 0) return tm;
    double a = tm / 100;
    while (lq_dcost(k, a) > 0) a /= 2;
    double t = tm / 2;
    double cdval = lq_dcost(k, t);
    int it = 0;
    while (fabs(cdval) > tol && fabs(a - b) > tol) {
        t = t - cdval / lq_ddcost(k, t);
        if (t < a || t > b) t = (a + b) / 2;
        cdval = lq_dcost(k, t);
        if (cdval > 0) b = t; else a = t;
        if (++it >= LQ_MAX_NEWTON) { if (capped) *capped = 1; break; }
    }
    return t;
}
/* linearquadratic.jl:191-195 */
void orc_lq_steer(int d, const double *R, const double *x0, const double *x1, double r, double *cost, double *topt)
{
    int same = 1;
    for (int i = 0; i < 2 * d; ++i) if (x0[i] != x1[i]) same = 0;
    if (same) { *cost = 0; *topt = 0; return; }
    abg_t k = lq_abg(d, R, x0, x1);
    double t = lq_topt_newton(k, r, NULL);
    *cost = lq_cost(k, t);
    *topt = t;
}
/* raw closures, for the golden-vector tests */
void orc_lq_cost_terms(int d, const double *R, const double *x0, const double *x1, double t, double *out3)
{
    abg_t k = lq_abg(d, R, x0, x1);
    out3[0] = lq_cost(k, t); out3[1] = lq_dcost(k, t); out3[2] = lq_ddcost(k, t);
}
/* x(x0,x1,t,s): cubic Hermite, Horner form; out = (p(s), v(s)) */
void orc_lq_state(int d, const double *x0, const double *x1, double t, double s, double *out)
{
    double it = 1.0 / t, it2 = it * it, it3 = it2 * it;
    const double *v0 = x0 + d, *v1 = x1 + d;
    for (int i = 0; i < d; ++i) {
        double dp = x1[i] - x0[i];
        double c2 = 3.0 * dp * it2 - (2.0 * v0[i] + v1[i]) * it;
        double c3 = (v0[i] + v1[i]) * it2 - 2.0 * dp * it3;
        out[i] = ((c3 * s + c2) * s + v0[i]) * s + x0[i];
        out[d + i] = (3.0 * c3 * s + 2.0 * c2) * s + v0[i];
    }
}

/* final neighbour tables: column q = { j != q : dcost(r) > 0 (prefilter, :213) and cost(t*) <= r },
 * forwards: cost(V[q] -> V[j]); backwards: cost(V[j] -> V[q]).  Two-call protocol like orc_rball_brute. */
void orc_lq_inball(const double *V, int64_t N, int d, const double *R, double r, int forwards, int64_t q0, int64_t q1,
                   int64_t *colptr, int64_t *rowval, double *nzval)
{
    int n = 2 * d;
    int count_only = (rowval == NULL);
    int64_t pos = 0;
    if (count_only) colptr[0] = 1;
    for (int64_t q = q0; q < q1; ++q) {
        for (int64_t j = 0; j < N; ++j) {
            if (j == q) continue;
            const double *x0 = forwards ? V + q * n : V + j * n;
            const double *x1 = forwards ? V + j * n : V + q * n;
            int same = 1;
            for (int i = 0; i < n; ++i) if (x0[i] != x1[i]) same = 0;
            abg_t k = lq_abg(d, R, x0, x1);
            if (!(lq_dcost(k, r) > 0)) continue; /* cands = cd .> 0, linearquadratic.jl:213 */
            double cost;
            if (same) {
                cost = 0; /* steer: x0 == x1 -> (0, 0), linearquadratic.jl:192 (duplicate states) */
            } else {
                double t = lq_topt_newton(k, r, NULL);
                cost = lq_cost(k, t);
            }
            if (cost <= r) {
                if (!count_only) { rowval[pos] = j + 1; nzval[pos] = cost; }
                ++pos;
            }
        }
        if (count_only) colptr[q - q0 + 1] = pos + 1;
    }
}

/* is_free_motion(v, w, CC, SS) for the LinearQuadratic metric: statespaces.jl:153-158 with
 * collision_waypoints of linearquadratic.jl:85-88.  count += 1 per segment test that runs. */
int orc_lq_is_free_motion(const orc_checker *CC, const orc_space *S, int d, const double *R, double r,
                          const double *v, const double *w, int64_t *count)
{
    double cost, t;
    orc_lq_steer(d, R, v, w, r, &cost, &t);
    double wps[5][2 * MAXD];
    for (int i = 0; i < 5; ++i) {
        double s = ((double)i * t) / 4.0; /* linspace(0, t, 5)[i+1] = ((len-i)*0 + (i-1)*t)/(len-1) */
        orc_lq_state(d, v, w, t, s, wps[i]);
    }
    for (int i = 0; i < 4; ++i)
        if (!orc_is_free_motion_straight(CC, S, wps[i], wps[i + 1], count)) return 0;
    return 1;
}
void orc_lq_edges_free_csc(const orc_checker *CC, const orc_space *S, int d, const double *R, double r,
                           const double *V, const int64_t *colptr, const int64_t *rowval, int64_t c0, int64_t c1,
                           uint8_t *out, int64_t *count)
{
    int n = 2 * d;
    for (int64_t x = c0; x < c1; ++x)
        for (int64_t e = colptr[x - c0] - 1; e < colptr[x - c0 + 1] - 1; ++e) {
            int64_t y = rowval[e] - 1;
            out[e] = (uint8_t)orc_lq_is_free_motion(CC, S, d, R, r, V + y * n, V + x * n, count);
        }
}
