// closest.cu -- K11: closest / closeR under a weight matrix, batched.
//
// Replaces closest(p, shape, W) / closeR(p, CC, W, r2) (SAT2D.jl:208-285: circle = Newton on the multiplier
// with the halving line search, polygon = Cholesky transform + closest_polypts; boxesND.jl:61-86: box =
// bounded least squares) for a whole batch of query points, each with its own weight matrix -- the proposal
// construction of the Monte-Carlo estimator asks this for every step of the nominal trajectory (W_t = the
// inverse marginal covariance).  One thread per (point, basic shape) evaluates closest(); one thread per point
// then keeps the shapes with d2 < r2 in ascending order (stable insertion, as the reference's sort!(by=first)).
// Operation order = oracle/closest.c, every operation an explicit _rn intrinsic: results are bit-identical.
#include "common.cuh"
#include "predicates.cuh"

namespace mpb {

constexpr int kCpMaxD = 4;

__device__ __forceinline__ void cp_eig2(double a, double b, double c, double *s1, double *s2, double *v1, double *v2) {
    if (b == 0.0) {
        *s1 = a; *s2 = c; v1[0] = 1.0; v1[1] = 0.0; v2[0] = 0.0; v2[1] = 1.0;
        return;
    }
    const double tau = ddiv(dsub(c, a), dmul(2.0, b));
    const double t = ddiv(tau >= 0 ? 1.0 : -1.0, dadd(fabs(tau), __dsqrt_rn(dadd(1.0, dmul(tau, tau)))));
    const double cs = ddiv(1.0, __dsqrt_rn(dadd(1.0, dmul(t, t)))), sn = dmul(t, cs);
    *s1 = dsub(a, dmul(t, b));
    *s2 = dadd(c, dmul(t, b));
    v1[0] = cs; v1[1] = -sn;
    v2[0] = sn; v2[1] = cs;
}

__device__ void cp_circle(const double *p, const double *rec, const double *W, double *d2, double *x) {
    double s1, s2, v1[2], v2[2];
    cp_eig2(W[0], W[1], W[3], &s1, &s2, v1, v2);
    const double r = rec[2];
    const double ct0 = dsub(p[0], rec[0]), ct1 = dsub(p[1], rec[1]);
    const double p1 = dadd(dmul(v1[0], ct0), dmul(v1[1], ct1));
    const double p2 = dadd(dmul(v2[0], ct0), dmul(v2[1], ct1));
    double lambda = 1.0;
    double q1 = ddiv(dmul(p1, s1), dadd(lambda, s1)), q2 = ddiv(dmul(p2, s2), dadd(lambda, s2));
    double f = dsub(dadd(dmul(q1, q1), dmul(q2, q2)), dmul(r, r));
    for (int it = 0; it < 100 && fabs(f) > 1e-8; ++it) {
        const double fp = dadd(dmul(ddiv(-2.0, dadd(lambda, s1)), dmul(q1, q1)), dmul(ddiv(-2.0, dadd(lambda, s2)), dmul(q2, q2)));
        double alpha = 1.0, lnew = lambda, fnew = f, n1 = q1, n2 = q2;
        for (int k = 0; k < 60; ++k) {
            lnew = dsub(lambda, ddiv(dmul(alpha, f), fp));
            n1 = ddiv(dmul(p1, s1), dadd(lnew, s1));
            n2 = ddiv(dmul(p2, s2), dadd(lnew, s2));
            fnew = dsub(dadd(dmul(n1, n1), dmul(n2, n2)), dmul(r, r));
            if (fabs(fnew) < fabs(f)) break;
            alpha = ddiv(alpha, 2.0);
        }
        f = fnew; lambda = lnew; q1 = n1; q2 = n2;
    }
    x[0] = dadd(dadd(rec[0], ddiv(dmul(dmul(v1[0], p1), s1), dadd(lambda, s1))), ddiv(dmul(dmul(v2[0], p2), s2), dadd(lambda, s2)));
    x[1] = dadd(dadd(rec[1], ddiv(dmul(dmul(v1[1], p1), s1), dadd(lambda, s1))), ddiv(dmul(dmul(v2[1], p2), s2), dadd(lambda, s2)));
    const double e1 = dsub(p1, q1), e2 = dsub(p2, q2);
    *d2 = dadd(dmul(s1, dmul(e1, e1)), dmul(s2, dmul(e2, e2)));
}

__device__ void cp_polygon(const double *p, const double *rec, int K, const double *W, double *d2, double *x) {
    const double *pts = rec + 4;
    const double L11 = __dsqrt_rn(W[0]), L12 = ddiv(W[1], L11), L22 = __dsqrt_rn(dsub(W[3], dmul(L12, L12)));
    const double q0 = dadd(dmul(L11, p[0]), dmul(L12, p[1])), q1 = dmul(L22, p[1]);
    double d2min = __longlong_as_double(0x7ff0000000000000LL), vm0 = 0.0, vm1 = 0.0;
    for (int i = 0; i < K; ++i) {
        const int j = (i + 1 == K) ? 0 : i + 1;
        const double a0 = dadd(dmul(L11, pts[2 * i]), dmul(L12, pts[2 * i + 1])), a1 = dmul(L22, pts[2 * i + 1]);
        const double b0 = dadd(dmul(L11, pts[2 * j]), dmul(L12, pts[2 * j + 1])), b1 = dmul(L22, pts[2 * j + 1]);
        const double e0 = dsub(b0, a0), e1 = dsub(b1, a1);
        const double t = ddiv(dadd(dmul(e0, dsub(q0, a0)), dmul(e1, dsub(q1, a1))), dadd(dmul(e0, e0), dmul(e1, e1)));
        double v0, v1;
        if (t < 0) { v0 = a0; v1 = a1; }
        else if (t < 1) { v0 = dadd(a0, dmul(t, e0)); v1 = dadd(a1, dmul(t, e1)); }
        else { v0 = b0; v1 = b1; }
        const double g0 = dsub(q0, v0), g1 = dsub(q1, v1);
        const double dd = dadd(dmul(g0, g0), dmul(g1, g1));
        if (dd < d2min) { d2min = dd; vm0 = v0; vm1 = v1; }
    }
    x[1] = ddiv(vm1, L22);
    x[0] = ddiv(dsub(vm0, dmul(L12, x[1])), L11);
    const double g0 = dsub(x[0], p[0]), g1 = dsub(x[1], p[1]);
    *d2 = dadd(dmul(g0, dadd(dmul(W[0], g0), dmul(W[1], g1))), dmul(g1, dadd(dmul(W[2], g0), dmul(W[3], g1))));
}

template <int D>
__device__ void cp_box(const double *p, const double *lo, const double *hi, const double *W, double *d2, double *x) {
    int ncodes = 1;
#pragma unroll
    for (int i = 0; i < D; ++i) ncodes *= 3;
    double best = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = p[i];
    for (int code = 0; code < ncodes; ++code) {
        int st[D], fr[D], nf = 0, cc = code;
        double v[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            st[i] = cc % 3; cc /= 3;
            if (st[i] == 0) fr[nf++] = i;
            v[i] = (st[i] == 1) ? lo[i] : hi[i];
        }
        if (nf > 0) {
            double A[D * D], b[D], Lw[D * D], y[D], z[D];
            for (int a = 0; a < nf; ++a) {
                double s = 0.0;
                for (int j = 0; j < D; ++j)
                    if (st[j] != 0) s = dadd(s, dmul(W[fr[a] * D + j], dsub(v[j], p[j])));
                b[a] = -s;
                for (int c2 = 0; c2 < nf; ++c2) A[a * nf + c2] = W[fr[a] * D + fr[c2]];
            }
            bool ok = true;
            for (int i = 0; i < nf && ok; ++i)
                for (int j = 0; j <= i; ++j) {
                    double s = A[i * nf + j];
                    for (int k = 0; k < j; ++k) s = dsub(s, dmul(Lw[i * nf + k], Lw[j * nf + k]));
                    if (i == j) { if (!(s > 0)) { ok = false; break; } Lw[i * nf + i] = __dsqrt_rn(s); }
                    else Lw[i * nf + j] = ddiv(s, Lw[j * nf + j]);
                }
            if (!ok) continue;
            for (int i = 0; i < nf; ++i) {
                double s = b[i];
                for (int k = 0; k < i; ++k) s = dsub(s, dmul(Lw[i * nf + k], y[k]));
                y[i] = ddiv(s, Lw[i * nf + i]);
            }
            for (int i = nf - 1; i >= 0; --i) {
                double s = y[i];
                for (int k = i + 1; k < nf; ++k) s = dsub(s, dmul(Lw[k * nf + i], z[k]));
                z[i] = ddiv(s, Lw[i * nf + i]);
            }
            for (int a = 0; a < nf; ++a) v[fr[a]] = dadd(p[fr[a]], z[a]);
        }
        bool feas = true;
        for (int i = 0; i < D; ++i) feas = feas && (lo[i] <= v[i] && v[i] <= hi[i]);
        if (!feas) continue;
        double q = 0.0;
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
            for (int j = 0; j < D; ++j) s = dadd(s, dmul(W[i * D + j], dsub(v[j], p[j])));
            q = dadd(q, dmul(dsub(v[i], p[i]), s));
        }
        if (q < best) { best = q; for (int i = 0; i < D; ++i) x[i] = v[i]; }
    }
    *d2 = best;
}

// thread per (point i, basic shape s): all_d2[i*S + s], all_x[(i*S + s)*DW ..]
template <int DW, int KIND>
__global__ void __launch_bounds__(128)
closest_kernel(const double *__restrict__ P, const double *__restrict__ Ws, int64_t n, int S, const double *__restrict__ T,
               int M, double *__restrict__ all_d2, double *__restrict__ all_x) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * S) return;
    const int64_t i = g / S;
    const int s = (int)(g - i * S);
    double p[DW], W[DW * DW], x[DW], d2;
#pragma unroll
    for (int k = 0; k < DW; ++k) p[k] = P[i * DW + k];
#pragma unroll
    for (int k = 0; k < DW * DW; ++k) W[k] = Ws[i * DW * DW + k];
    if (KIND == 0) {
        const Obs2 O(T);
        const double *rec = O.data(s);
        if (O.kind(s) == 0) cp_circle(p, rec, W, &d2, x);
        else cp_polygon(p, rec, O.K(s), W, &d2, x);
    } else {
        cp_box<DW>(p, T + (size_t)s * DW, T + (size_t)M * DW + (size_t)s * DW, W, &d2, x);
    }
    all_d2[g] = d2;
#pragma unroll
    for (int k = 0; k < DW; ++k) all_x[g * DW + k] = x[k];
}

// thread per point: closeR = the shapes with d2 < r2, ascending, ties in shape order
template <int DW>
__global__ void __launch_bounds__(128)
closeR_kernel(const double *__restrict__ all_d2, const double *__restrict__ all_x, int64_t n, int S, double r2,
              int *__restrict__ count, double *__restrict__ d2_out, int *__restrict__ shape_out, double *__restrict__ x_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cnt = 0;
    for (int s = 0; s < S; ++s) {
        const double dd = all_d2[i * S + s];
        if (!(dd < r2)) continue;
        int pos = cnt;
        while (pos > 0 && d2_out[i * S + pos - 1] > dd) {
            d2_out[i * S + pos] = d2_out[i * S + pos - 1];
            shape_out[i * S + pos] = shape_out[i * S + pos - 1];
            for (int k = 0; k < DW; ++k) x_out[(i * S + pos) * DW + k] = x_out[(i * S + pos - 1) * DW + k];
            --pos;
        }
        d2_out[i * S + pos] = dd;
        shape_out[i * S + pos] = s;
        for (int k = 0; k < DW; ++k) x_out[(i * S + pos) * DW + k] = all_x[(i * S + s) * DW + k];
        ++cnt;
    }
    count[i] = cnt;
}

int close_points_device(const mpb200_obstacles *o, const double *dP, const double *dW, int64_t n, int dw, double r2,
                        int *d_count, double *d_d2, int *d_shape, double *d_x, double *d_all_d2, double *d_all_x) {
    const int S = o->kind == 0 ? o->n_shapes : o->M;
    if (dw > kCpMaxD) return fail(MPB200_EARG, "closest points: workspace dimension <= %d", kCpMaxD);
    if (o->kind == 0 && dw != 2) return fail(MPB200_EARG, "2-D obstacles need a 2-D workspace");
    if (o->kind == 1 && o->d != dw) return fail(MPB200_EARG, "box dimension %d != workspace dimension %d", o->d, dw);
    if (n == 0 || S == 0) return 0;
    cudaStream_t st = ctx().stream;
    const unsigned g1 = (unsigned)ceil_div(n * S, 128), g2 = (unsigned)ceil_div(n, 128);
    const double *T = o->table.as<double>();
#define CALL(DW_, K_)                                                                                               \
    do {                                                                                                            \
        closest_kernel<DW_, K_><<<g1, 128, 0, st>>>(dP, dW, n, S, T, o->M, d_all_d2, d_all_x);                      \
        MPB_LAUNCHED();                                                                                             \
        closeR_kernel<DW_><<<g2, 128, 0, st>>>(d_all_d2, d_all_x, n, S, r2, d_count, d_d2, d_shape, d_x);           \
        MPB_LAUNCHED();                                                                                             \
    } while (0)
    if (o->kind == 0) CALL(2, 0);
    else if (dw == 1) CALL(1, 1);
    else if (dw == 2) CALL(2, 1);
    else if (dw == 3) CALL(3, 1);
    else CALL(4, 1);
#undef CALL
    return 0;
}

}  // namespace mpb
