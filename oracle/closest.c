/*
 * closest.c -- oracle restatement of closest / closeR under a weight matrix W: the geometry the reference
 * ships for Monte-Carlo importance sampling.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows src/collisioncheckers/SAT2D.jl:
 *   :212-238  closest(p, C::Circle, W): eigen-decomposition of W, Newton on the multiplier lambda with the
 *             halving line search, closest point and squared W-distance
 *   :239-258  closest_polypts (closest point on a polygon boundary) and closest(p, P::Polygon, W):
 *             Cholesky transform  L = chol(W),  closest_polypts(L p, [L pt]),  back-transform,  (x-p)'W(x-p)
 *   :259-285  closest over Compound2D parts; closeR = the basic shapes closer than r2, sorted ascending
 * and src/collisioncheckers/boxesND.jl:
 *   :61-70    closest(p, BB::BoxBounds, W) = argmin (v-p)'W(v-p) over the box (the reference calls its BVLS port)
 *   :72-86    closest over the box list; closeR
 *
 * PARITY UNPINNED at the rounding level: eigfact / chol / inv come from LAPACK and StaticArrays, bvls.jl is an
 * active-set iteration whose path is not reproduced.  What IS pinned: every quantity is the unique minimiser of
 * a strictly convex problem, checked against an independent brute-force search (tests/test_oracle_closest.py).
 * The operation order below is THE specification the CUDA kernel (csrc/closest.cu) reproduces bit for bit:
 *   2x2 symmetric eigen: one Jacobi rotation;  chol(W) = upper L with W = L'L;  inv(L) v by back substitution;
 *   box: exact enumeration of the 3^d active sets (each coordinate free / at lo / at hi; free block solved by
 *   Cholesky), first minimum in code order;  Newton capped at 100 iterations, line search at 60 halvings.
 */
#include "mp_oracle.h"
#include <math.h>
#include <string.h>

/* symmetric 2x2 [[a,b],[b,c]] = s1 v1 v1' + s2 v2 v2' */
static void eig2(double a, double b, double c, double *s1, double *s2, double *v1, double *v2)
{
    if (b == 0.0) {
        *s1 = a; *s2 = c; v1[0] = 1; v1[1] = 0; v2[0] = 0; v2[1] = 1;
        return;
    }
    double tau = (c - a) / (2.0 * b);
    double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
    double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
    *s1 = a - t * b;
    *s2 = c + t * b;
    v1[0] = cs; v1[1] = -sn;
    v2[0] = sn; v2[1] = cs;
}

/* SAT2D.jl:212-238; rec = (cx, cy, r, ...) */
void orc_closest_circle(const double *p, const double *rec, const double *W, double *d2, double *x)
{
    double s1, s2, v1[2], v2[2];
    eig2(W[0], W[1], W[3], &s1, &s2, v1, v2);
    const double r = rec[2];
    double ct0 = p[0] - rec[0], ct1 = p[1] - rec[1];
    double p1 = v1[0] * ct0 + v1[1] * ct1;
    double p2 = v2[0] * ct0 + v2[1] * ct1;
    double lambda = 1.0;
    double q1 = (p1 * s1) / (lambda + s1), q2 = (p2 * s2) / (lambda + s2);
    double f = (q1 * q1 + q2 * q2) - r * r;
    for (int it = 0; it < 100 && fabs(f) > 1e-8; ++it) {
        double fp = (-2.0 / (lambda + s1)) * (q1 * q1) + (-2.0 / (lambda + s2)) * (q2 * q2);
        double alpha = 1.0, lnew = lambda, fnew = f, n1 = q1, n2 = q2;
        for (int k = 0; k < 60; ++k) { /* "crappy linesearch" */
            lnew = lambda - alpha * f / fp;
            n1 = (p1 * s1) / (lnew + s1);
            n2 = (p2 * s2) / (lnew + s2);
            fnew = (n1 * n1 + n2 * n2) - r * r;
            if (fabs(fnew) < fabs(f)) break;
            alpha /= 2;
        }
        f = fnew; lambda = lnew; q1 = n1; q2 = n2;
    }
    x[0] = (rec[0] + ((v1[0] * p1) * s1) / (lambda + s1)) + ((v2[0] * p2) * s2) / (lambda + s2);
    x[1] = (rec[1] + ((v1[1] * p1) * s1) / (lambda + s1)) + ((v2[1] * p2) * s2) / (lambda + s2);
    double e1 = p1 - q1, e2 = p2 - q2;
    *d2 = s1 * (e1 * e1) + s2 * (e2 * e2);
}

/* SAT2D.jl:254-258 with closest_polypts :240-252; rec = (xlo xhi ylo yhi, K points, ...) */
void orc_closest_polygon(const double *p, const double *rec, int K, const double *W, double *d2, double *x)
{
    const double *pts = rec + 4;
    double L11 = sqrt(W[0]), L12 = W[1] / L11, L22 = sqrt(W[3] - L12 * L12);
    double q0 = L11 * p[0] + L12 * p[1], q1 = L22 * p[1];
    double d2min = INFINITY, vm0 = 0, vm1 = 0;
    for (int i = 0; i < K; ++i) {
        int j = (i + 1 == K) ? 0 : i + 1;
        double a0 = L11 * pts[2 * i] + L12 * pts[2 * i + 1], a1 = L22 * pts[2 * i + 1];
        double b0 = L11 * pts[2 * j] + L12 * pts[2 * j + 1], b1 = L22 * pts[2 * j + 1];
        double e0 = b0 - a0, e1 = b1 - a1;
        double t = (e0 * (q0 - a0) + e1 * (q1 - a1)) / (e0 * e0 + e1 * e1);
        double v0, v1;
        if (t < 0) { v0 = a0; v1 = a1; }
        else if (t < 1) { v0 = a0 + t * e0; v1 = a1 + t * e1; }
        else { v0 = b0; v1 = b1; }
        double dd = (q0 - v0) * (q0 - v0) + (q1 - v1) * (q1 - v1);
        if (dd < d2min) { d2min = dd; vm0 = v0; vm1 = v1; }
    }
    x[1] = vm1 / L22;
    x[0] = (vm0 - L12 * x[1]) / L11;
    double g0 = x[0] - p[0], g1 = x[1] - p[1];
    *d2 = g0 * (W[0] * g0 + W[1] * g1) + g1 * (W[2] * g0 + W[3] * g1);
}

/* boxesND.jl:61-70: argmin (v-p)'W(v-p), lo <= v <= hi, W d x d row-major SPD, d <= ORC_CP_MAXD */
void orc_closest_box(const double *p, const double *lo, const double *hi, int d, const double *W, double *d2, double *x)
{
    int ncodes = 1;
    for (int i = 0; i < d; ++i) ncodes *= 3;
    double best = INFINITY;
    for (int i = 0; i < d; ++i) x[i] = p[i];
    for (int code = 0; code < ncodes; ++code) {
        int st[ORC_CP_MAXD], fr[ORC_CP_MAXD], nf = 0, cc = code;
        double v[ORC_CP_MAXD];
        for (int i = 0; i < d; ++i) {
            st[i] = cc % 3; cc /= 3;
            if (st[i] == 0) fr[nf++] = i;
            v[i] = (st[i] == 1) ? lo[i] : hi[i];
        }
        if (nf > 0) { /* W_ff (v_f - p_f) = - W_fc (v_c - p_c) */
            double A[ORC_CP_MAXD * ORC_CP_MAXD], b[ORC_CP_MAXD], Lw[ORC_CP_MAXD * ORC_CP_MAXD], y[ORC_CP_MAXD], z[ORC_CP_MAXD];
            for (int a = 0; a < nf; ++a) {
                double s = 0;
                for (int j = 0; j < d; ++j)
                    if (st[j] != 0) s = s + W[fr[a] * d + j] * (v[j] - p[j]);
                b[a] = -s;
                for (int c2 = 0; c2 < nf; ++c2) A[a * nf + c2] = W[fr[a] * d + fr[c2]];
            }
            int ok = 1;
            for (int i = 0; i < nf && ok; ++i)
                for (int j = 0; j <= i; ++j) {
                    double s = A[i * nf + j];
                    for (int k = 0; k < j; ++k) s = s - Lw[i * nf + k] * Lw[j * nf + k];
                    if (i == j) { if (!(s > 0)) { ok = 0; break; } Lw[i * nf + i] = sqrt(s); }
                    else Lw[i * nf + j] = s / Lw[j * nf + j];
                }
            if (!ok) continue;
            for (int i = 0; i < nf; ++i) {
                double s = b[i];
                for (int k = 0; k < i; ++k) s = s - Lw[i * nf + k] * y[k];
                y[i] = s / Lw[i * nf + i];
            }
            for (int i = nf - 1; i >= 0; --i) {
                double s = y[i];
                for (int k = i + 1; k < nf; ++k) s = s - Lw[k * nf + i] * z[k];
                z[i] = s / Lw[i * nf + i];
            }
            for (int a = 0; a < nf; ++a) v[fr[a]] = p[fr[a]] + z[a];
        }
        int feas = 1;
        for (int i = 0; i < d; ++i) if (!(lo[i] <= v[i] && v[i] <= hi[i])) feas = 0;
        if (!feas) continue;
        double q = 0;
        for (int i = 0; i < d; ++i) {
            double s = 0;
            for (int j = 0; j < d; ++j) s = s + W[i * d + j] * (v[j] - p[j]);
            q = q + (v[i] - p[i]) * s;
        }
        if (q < best) { best = q; for (int i = 0; i < d; ++i) x[i] = v[i]; }
    }
    *d2 = best;
}

/* closeR(p, CC, W, r2) for n points, each with its own W (dw x dw row-major): per point the basic shapes with
 * d2 < r2 in ascending d2 (stable: ties keep shape order).  Outputs have capacity S = number of basic shapes per
 * point: count[i], then d2[i*S + k], shape[i*S + k], x[(i*S + k)*dw + .] for k < count[i].
 * all_d2 / all_x (optional, n x S) receive closest() for EVERY basic shape, unsorted. */
int orc_close_points(const orc_checker *CC, const double *P, const double *Ws, int64_t n, int dw, double r2,
                     int32_t *count, double *d2_out, int32_t *shape_out, double *x_out, double *all_d2, double *all_x)
{
    const int S = (CC->kind == 0) ? CC->obs2d->n_shapes : CC->M;
    if (dw > ORC_CP_MAXD || (CC->kind == 0 && dw != 2) || (CC->kind == 1 && dw != CC->d)) return -1;
    for (int64_t i = 0; i < n; ++i) {
        const double *p = P + i * dw, *W = Ws + i * dw * dw;
        int cnt = 0;
        for (int s = 0; s < S; ++s) {
            double dd, xx[ORC_CP_MAXD];
            if (CC->kind == 0) {
                const orc_obs2d *O = CC->obs2d;
                const double *rec = O->data + O->shape_off[s];
                if (O->shape_kind[s] == 0) orc_closest_circle(p, rec, W, &dd, xx);
                else orc_closest_polygon(p, rec, (O->shape_off[s + 1] - O->shape_off[s] - 4) / 6, W, &dd, xx);
            } else {
                orc_closest_box(p, CC->box_lo + (size_t)s * dw, CC->box_hi + (size_t)s * dw, dw, W, &dd, xx);
            }
            if (all_d2) { all_d2[i * S + s] = dd; for (int k = 0; k < dw; ++k) all_x[(i * S + s) * dw + k] = xx[k]; }
            if (dd < r2) { /* insertion keeps ascending order, ties after earlier shapes */
                int pos = cnt;
                while (pos > 0 && d2_out[i * S + pos - 1] > dd) {
                    d2_out[i * S + pos] = d2_out[i * S + pos - 1];
                    shape_out[i * S + pos] = shape_out[i * S + pos - 1];
                    for (int k = 0; k < dw; ++k) x_out[(i * S + pos) * dw + k] = x_out[(i * S + pos - 1) * dw + k];
                    --pos;
                }
                d2_out[i * S + pos] = dd;
                shape_out[i * S + pos] = s;
                for (int k = 0; k < dw; ++k) x_out[(i * S + pos) * dw + k] = xx[k];
                ++cnt;
            }
        }
        count[i] = cnt;
    }
    return 0;
}
