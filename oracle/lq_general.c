/*
 * lq_general.c -- oracle restatement of the linear-quadratic steering cost for GENERAL linear-affine
 * systems  xdot = A x + B u + c  with nilpotent A, cost  int (1 + u'Ru).  TEST INFRASTRUCTURE ONLY.
 *
 * Follows src/statespaces/linearquadratic.jl:
 *   :94-98    expAt: e^{At} = sum_{k<n} A^k t^k / k!   (nilpotent A only)
 *   :126-157  LinearQuadratic2BVP:  G(t) = int_0^t e^{As} B R^-1 B' e^{A's} ds,  xbar(t) = e^{At} x + int_0^t e^{As} c ds,
 *             cost(x,y,t) = t + (y - xbar)' G^-1 (y - xbar), dcost = d/dt, ddcost = d2/dt2,
 *             x(x,y,t,s) = xbar(s) + G(s) e^{A'(t-s)} G(t)^-1 (y - xbar(t))
 *   :160-195  topt_newton, steer;   :196-225 steer_pairwise (prefilter dcost(r) > 0, keep cost <= r)
 *   :85-88    collision_waypoints
 *
 * The reference turns the symbolic expressions into closures through SymPy's printer (third-party,
 * unpinned operation order -- PARITY UNPINNED at the rounding level).  This file evaluates
 * the same quantities NUMERICALLY, and its operation order is THE specification the CUDA kernels
 * (csrc/lq_general.cu) reproduce bit for bit:
 *   tables (once per system):  Ak = A^k / k!,  dk = Ak c / (k+1),  BRB = B R^-1 B',
 *                              Gp = (1/p) sum_{i+j=p-1} Ak[i] BRB Ak[j]'      (G(t) = sum_p Gp t^p)
 *   per (x, y, t):  xbar = sum_k (Ak x) t^k + dk t^{k+1};  e = y - xbar;  G = sum_p Gp t^p;
 *                   lam = G^-1 e (Cholesky);  f = A y + c;  h = f + BRB lam;  mu = G^-1 h;
 *                   cost = t + e'lam;  dcost = 1 - 2 lam'f - lam'(BRB lam);  ddcost = 2 (mu + A'lam)'h
 * (dcost / ddcost: differentiate with dG/dt = A G + G A' + BRB and dxbar/dt = A xbar + c; the dcost form
 *  at t = r is exactly the reference's dense prefilter, linearquadratic.jl:205-211.)
 * Pinned by tests/golden/lq_general.json: the reference's own SymPy construction re-run with sympy 1.14 at
 * 50 digits (tests/golden/gen_lq_general_golden.py) for a drifting double integrator and a triple integrator.
 */
#include "mp_oracle.h"
#include <math.h>
#include <string.h>

#define GN ORC_LQG_MAXN

static void matmul(int n, const double *X, const double *Y, double *Z)
{ /* Z = X Y, row-major, k ascending */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < n; ++k) s = s + X[i * n + k] * Y[k * n + j];
            Z[i * n + j] = s;
        }
}

/* SPD solve by Cholesky (lower, row by row), forward then backward substitution; returns 0 if not SPD */
static int chol_factor(int n, const double *G, double *Lw)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = G[i * n + j];
            for (int k = 0; k < j; ++k) s = s - Lw[i * n + k] * Lw[j * n + k];
            if (i == j) {
                if (!(s > 0)) return 0;
                Lw[i * n + i] = sqrt(s);
            } else {
                Lw[i * n + j] = s / Lw[j * n + j];
            }
        }
    return 1;
}
static void chol_solve(int n, const double *Lw, const double *b, double *x)
{
    double y[GN];
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s = s - Lw[i * n + k] * y[k];
        y[i] = s / Lw[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s = s - Lw[k * n + i] * x[k];
        x[i] = s / Lw[i * n + i];
    }
}

/* tables of a system; returns 0 on success, <0 if A is not nilpotent / R not SPD / sizes out of range */
int orc_lqg_setup(int n, int m, const double *A, const double *B, const double *c, const double *R, orc_lqg *S)
{
    if (n < 1 || n > GN || m < 1 || m > GN) return -1;
    memset(S, 0, sizeof(*S));
    S->n = n;
    for (int i = 0; i < n * n; ++i) S->A[i] = A[i];
    for (int i = 0; i < n; ++i) S->c[i] = c[i];
    /* Ak[k] = A^k / k! */
    double P[GN * GN], Q[GN * GN];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) P[i * n + j] = (i == j) ? 1.0 : 0.0;
    double fact = 1.0;
    for (int k = 0; k < n; ++k) {
        if (k > 0) { matmul(n, P, A, Q); memcpy(P, Q, sizeof(double) * n * n); fact = fact * (double)k; }
        for (int i = 0; i < n * n; ++i) S->Ak[k][i] = P[i] / fact;
    }
    matmul(n, P, A, Q); /* A^n must vanish (linearquadratic.jl:96) */
    for (int i = 0; i < n * n; ++i) if (Q[i] != 0.0) return -2;
    /* dk[k] = Ak[k] c / (k+1) */
    for (int k = 0; k < n; ++k)
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int j = 0; j < n; ++j) s = s + S->Ak[k][i * n + j] * c[j];
            S->dk[k][i] = s / (double)(k + 1);
        }
    /* Rinv by Cholesky: Rinv column j = solve(R, e_j) */
    double Lr[GN * GN], Rinv[GN * GN], ej[GN], col[GN];
    if (!chol_factor(m, R, Lr)) return -3;
    for (int j = 0; j < m; ++j) {
        for (int i = 0; i < m; ++i) ej[i] = (i == j) ? 1.0 : 0.0;
        chol_solve(m, Lr, ej, col);
        for (int i = 0; i < m; ++i) Rinv[i * m + j] = col[i];
    }
    /* BRB = B Rinv B'  (B is n x m row-major) */
    double BR[GN * GN];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) {
            double s = 0;
            for (int k = 0; k < m; ++k) s = s + B[i * m + k] * Rinv[k * m + j];
            BR[i * m + j] = s;
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < m; ++k) s = s + BR[i * m + k] * B[j * m + k];
            S->BRB[i * n + j] = s;
        }
    /* Gp[p] = (1/p) sum_{i+j=p-1} Ak[i] BRB Ak[j]',  p = 1 .. 2n-1 */
    S->np = 2 * n - 1;
    for (int p = 1; p <= S->np; ++p) {
        double acc[GN * GN];
        for (int i = 0; i < n * n; ++i) acc[i] = 0;
        for (int i = 0; i < n; ++i) {
            int j = p - 1 - i;
            if (j < 0 || j >= n) continue;
            double T1[GN * GN];
            matmul(n, S->Ak[i], S->BRB, T1);
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) {
                    double s = 0;
                    for (int k = 0; k < n; ++k) s = s + T1[a * n + k] * S->Ak[j][b * n + k];
                    acc[a * n + b] = acc[a * n + b] + s;
                }
        }
        for (int i = 0; i < n * n; ++i) S->Gp[p - 1][i] = acc[i] / (double)p;
    }
    return 0;
}

/* xbar(t) = sum_k (Ak x) t^k + dk t^{k+1} */
static void lqg_xbar(const orc_lqg *S, const double *x, double t, double *xb)
{
    const int n = S->n;
    double tp = 1.0;
    for (int i = 0; i < n; ++i) xb[i] = 0;
    for (int k = 0; k < n; ++k) {
        const double tp1 = tp * t;
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int j = 0; j < n; ++j) s = s + S->Ak[k][i * n + j] * x[j];
            xb[i] = xb[i] + (s * tp + S->dk[k][i] * tp1);
        }
        tp = tp1;
    }
}
static void lqg_G(const orc_lqg *S, double t, double *G)
{
    const int n = S->n;
    for (int i = 0; i < n * n; ++i) G[i] = 0;
    double tp = 1.0;
    for (int p = 0; p < S->np; ++p) {
        tp = tp * t;
        for (int i = 0; i < n * n; ++i) G[i] = G[i] + S->Gp[p][i] * tp;
    }
}

/* out3 = (cost, dcost, ddcost) at time t; returns 0 if G(t) is not numerically SPD (then out3 = +inf, +1, 0) */
int orc_lqg_cost_terms(const orc_lqg *S, const double *x, const double *y, double t, double *out3)
{
    const int n = S->n;
    double xb[GN], e[GN], G[GN * GN], Lw[GN * GN], lam[GN], f[GN], h[GN], mu[GN], bl[GN];
    lqg_xbar(S, x, t, xb);
    for (int i = 0; i < n; ++i) e[i] = y[i] - xb[i];
    lqg_G(S, t, G);
    if (!chol_factor(n, G, Lw)) { out3[0] = INFINITY; out3[1] = 1.0; out3[2] = 0.0; return 0; }
    chol_solve(n, Lw, e, lam);
    for (int i = 0; i < n; ++i) {
        double s = 0, b = 0;
        for (int j = 0; j < n; ++j) { s = s + S->A[i * n + j] * y[j]; b = b + S->BRB[i * n + j] * lam[j]; }
        f[i] = s + S->c[i];
        bl[i] = b;
        h[i] = f[i] + b;
    }
    chol_solve(n, Lw, h, mu);
    double el = 0, lf = 0, lb = 0, dd = 0;
    for (int i = 0; i < n; ++i) {
        double atl = 0; /* (A' lam)_i */
        for (int j = 0; j < n; ++j) atl = atl + S->A[j * n + i] * lam[j];
        el = el + e[i] * lam[i];
        lf = lf + lam[i] * f[i];
        lb = lb + lam[i] * bl[i];
        dd = dd + (mu[i] + atl) * h[i];
    }
    out3[0] = t + el;
    out3[1] = (1.0 - 2.0 * lf) - lb;
    out3[2] = 2.0 * dd;
    return 1;
}
static double lqg_dcost(const orc_lqg *S, const double *x, const double *y, double t)
{
    double o[3];
    orc_lqg_cost_terms(S, x, y, t, o);
    return o[1];
}

#define LQG_MAX_NEWTON 200
#define LQG_MAX_HALVINGS 60
/* linearquadratic.jl:175-190 (the a /= 2 search is capped: with drift c the derivative can stay positive) */
static double lqg_topt_newton(const orc_lqg *S, const double *x, const double *y, double tm)
{
    const double tol = 1e-6;
    double b = tm;
    if (lqg_dcost(S, x, y, b) < 0) return tm;
    double a = tm / 100;
    for (int k = 0; k < LQG_MAX_HALVINGS && lqg_dcost(S, x, y, a) > 0; ++k) a /= 2;
    double t = tm / 2;
    double o[3];
    orc_lqg_cost_terms(S, x, y, t, o);
    double cdval = o[1];
    int it = 0;
    while (fabs(cdval) > tol && fabs(a - b) > tol) {
        t = t - cdval / o[2];
        if (!(t >= a && t <= b)) t = (a + b) / 2; /* also catches NaN */
        orc_lqg_cost_terms(S, x, y, t, o);
        cdval = o[1];
        if (cdval > 0) b = t; else a = t;
        if (++it >= LQG_MAX_NEWTON) break;
    }
    return t;
}
/* linearquadratic.jl:191-195 */
void orc_lqg_steer(const orc_lqg *S, const double *x0, const double *x1, double r, double *cost, double *topt)
{
    int same = 1;
    for (int i = 0; i < S->n; ++i) if (x0[i] != x1[i]) same = 0;
    if (same) { *cost = 0; *topt = 0; return; }
    double t = lqg_topt_newton(S, x0, x1, r);
    double o[3];
    orc_lqg_cost_terms(S, x0, x1, t, o);
    *cost = o[0];
    *topt = t;
}
/* x(x0, x1, t, s) = xbar(s) + G(s) e^{A'(t-s)} lam(t) */
void orc_lqg_state(const orc_lqg *S, const double *x0, const double *x1, double t, double s, double *out)
{
    const int n = S->n;
    double xb[GN], e[GN], G[GN * GN], Lw[GN * GN], lam[GN], w[GN], Gs[GN * GN];
    lqg_xbar(S, x0, t, xb);
    for (int i = 0; i < n; ++i) e[i] = x1[i] - xb[i];
    lqg_G(S, t, G);
    if (!chol_factor(n, G, Lw)) { for (int i = 0; i < n; ++i) out[i] = NAN; return; }
    chol_solve(n, Lw, e, lam);
    /* w = e^{A'(t-s)} lam = sum_k Ak[k]' lam (t-s)^k */
    const double ts = t - s;
    double tp = 1.0;
    for (int i = 0; i < n; ++i) w[i] = 0;
    for (int k = 0; k < n; ++k) {
        for (int i = 0; i < n; ++i) {
            double a = 0;
            for (int j = 0; j < n; ++j) a = a + S->Ak[k][j * n + i] * lam[j];
            w[i] = w[i] + a * tp;
        }
        tp = tp * ts;
    }
    lqg_xbar(S, x0, s, xb);
    lqg_G(S, s, Gs);
    for (int i = 0; i < n; ++i) {
        double a = 0;
        for (int j = 0; j < n; ++j) a = a + Gs[i * n + j] * w[j];
        out[i] = xb[i] + a;
    }
}

/* neighbour tables, as orc_lq_inball (two-call protocol) */
void orc_lqg_inball(const orc_lqg *S, const double *V, int64_t N, double r, int forwards, int64_t q0, int64_t q1,
                    int64_t *colptr, int64_t *rowval, double *nzval)
{
    const int n = S->n;
    int count_only = (rowval == NULL);
    int64_t pos = 0;
    if (count_only) colptr[0] = 1;
    for (int64_t q = q0; q < q1; ++q) {
        for (int64_t j = 0; j < N; ++j) {
            if (j == q) continue;
            const double *x0 = forwards ? V + q * n : V + j * n;
            const double *x1 = forwards ? V + j * n : V + q * n;
            if (!(lqg_dcost(S, x0, x1, r) > 0)) continue; /* cands = cd .> 0 */
            double cost, t;
            orc_lqg_steer(S, x0, x1, r, &cost, &t);
            if (cost <= r) {
                if (!count_only) { rowval[pos] = j + 1; nzval[pos] = cost; }
                ++pos;
            }
        }
        if (count_only) colptr[q - q0 + 1] = pos + 1;
    }
}

int orc_lqg_is_free_motion(const orc_checker *CC, const orc_space *Sp, const orc_lqg *S, double r, const double *v,
                           const double *w, int64_t *count)
{
    double cost, t;
    orc_lqg_steer(S, v, w, r, &cost, &t);
    double wps[5][GN];
    for (int i = 0; i < 5; ++i) {
        double s = ((double)i * t) / 4.0;
        if (t == 0.0) { for (int k = 0; k < S->n; ++k) wps[i][k] = v[k]; }  /* x0 == x1: the trajectory is the point */
        else orc_lqg_state(S, v, w, t, s, wps[i]);
    }
    for (int i = 0; i < 4; ++i)
        if (!orc_is_free_motion_straight(CC, Sp, wps[i], wps[i + 1], count)) return 0;
    return 1;
}
void orc_lqg_edges_free_csc(const orc_checker *CC, const orc_space *Sp, const orc_lqg *S, double r, const double *V,
                            const int64_t *colptr, const int64_t *rowval, int64_t c0, int64_t c1, uint8_t *out,
                            int64_t *count)
{
    const int n = S->n;
    for (int64_t x = c0; x < c1; ++x)
        for (int64_t e = colptr[x - c0] - 1; e < colptr[x - c0 + 1] - 1; ++e) {
            int64_t y = rowval[e] - 1;
            out[e] = (uint8_t)orc_lqg_is_free_motion(CC, Sp, S, r, V + y * n, V + x * n, count);
        }
}
