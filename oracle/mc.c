/*
 * mc.c -- oracle for the Monte-Carlo (importance-sampled) trajectory collision-probability
 * estimator.  TEST INFRASTRUCTURE ONLY (see mp_oracle.h).
 *
 * PARITY UNPINNED: the estimator does not exist in /root/reference (only the paper links at
 * README.md:9-10 and the helper geometry closest/closeR in SAT2D.jl:208-285, boxesND.jl:61-86).
 * The specification below IS the contract; this file implements it
 * first and the CUDA kernel (csrc/mc.cu) must reproduce it bit for bit.  It is pinned only by
 * analytic known answers (tests/test_oracle_mc.py) and Random123's published Philox vectors.
 *
 * Specification
 *   closed loop   z_{t+1} = F_t z_t + G_t eps_t,  z_0 = 0,  eps_t in R^q,  t = 0..T-1
 *   workspace     w_t = wbar_t + Wz z_t                              (t = 1..T)
 *   event         hit = exists t: point w_t collides (swept = 0) or segment w_{t-1} -> w_t
 *                 collides (swept = 1, w_0 = wbar_0), using the reference's own predicates
 *   proposal      component k ~ alpha (k = 0: nominal, mean 0; k >= 1: mean mu_k over the
 *                 stacked noise), eps = xi + mu_k, xi ~ N(0, I)
 *   weight        w = 1 / (alpha_0 + sum_k alpha_k exp(mu_k.eps - |mu_k|^2 / 2))
 *   sums          S1 = sum w*hit, S2 = sum (w*hit)^2, S0 = sum w, n, hits
 *   randomness    Philox4x32-10, key = seed, counter = (rollout id lo, hi, block, stream):
 *                 stream 0 block 0 word pair 0 -> component choice; stream 1 block b -> Box-Muller
 *                 pair b (normals 2b, 2b+1 of the stacked noise); 53-bit uniforms in (0,1)
 *   arithmetic    IEEE double, one rounding per operation; log / sin / cos / exp are the
 *                 polynomial routines below (basic operations only), so that CPU and GPU agree
 *                 bit for bit.
 */
#include "mp_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- Philox4x32-10 (Salmon et al., SC'11; Random123) ---------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* 53-bit uniform in (0,1): ((hi >> 5) * 2^26 + (lo >> 6) + 0.5) * 2^-53 */
static inline double u53(uint32_t hi, uint32_t lo)
{
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

/* ---- deterministic elementary functions (basic operations only) ----------------------------- */
/* ln(x), x > 0 finite normal: x = m 2^e, m in [sqrt(1/2), sqrt(2)); ln m = 2 atanh((m-1)/(m+1)) */
double orc_det_log(double x)
{
    uint64_t b;
    memcpy(&b, &x, 8);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &b, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double z = (m - 1.0) / (m + 1.0), z2 = z * z;
    double s = 1.0 / 23.0;
    for (int k = 21; k >= 1; k -= 2) s = s * z2 + 1.0 / (double)k;
    return (double)e * 0.6931471805599453 + 2.0 * (z * s);
}
/* sin and cos of 2*pi*u, u in [0,1): octant reduction + Taylor series on [0, pi/4] */
void orc_det_sincos2pi(double u, double *sn, double *cs)
{
    double v = u * 8.0;
    int oct = (int)v;
    double f = v - (double)oct;
    if (oct & 1) f = 1.0 - f;
    double th = f * 0.7853981633974483, t2 = th * th;
    double ps = -1.0 / 355687428096000.0; /* -1/17! */
    ps = ps * t2 + 1.0 / 1307674368000.0; /* 1/15! */
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
    /* with th = (odd octant ? 1-f : f) pi/4:  oct 0:(s,c) 1:(c,s) 2:(c,-s) 3:(s,-c) 4:(-s,-c) 5:(-c,-s) 6:(-c,s) 7:(-s,c) */
    int o4 = oct & 3;
    double a = (o4 == 0 || o4 == 3) ? s : c;
    double b2 = (o4 == 0 || o4 == 3) ? c : s;
    *sn = (oct >= 4) ? -a : a;
    *cs = ((oct + 2) & 4) ? -b2 : b2;
}
/* exp(x): x = k ln2 + r, |r| <= ln2/2, Taylor degree 14, scaled by 2^k; clamps to [2^-1000, 2^1000] */
double orc_det_exp(double x)
{
    if (x > 690.0) x = 690.0;
    if (x < -690.0) x = -690.0;
    double kf = floor(x * 1.4426950408889634 + 0.5);
    double r = (x - kf * 0.6931471803691238) - kf * 1.9082149292705877e-10;
    double p = 1.0 / 87178291200.0; /* 1/14! */
    p = p * r + 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    int k = (int)kf;
    uint64_t b = (uint64_t)(k + 1023) << 52;
    double sc;
    memcpy(&sc, &b, 8);
    return p * sc;
}

/* ---- the estimator ---------------------------------------------------------------------------- */
typedef struct {
    int32_t T, nz, q, dw;
    const double *F;    /* T x nz x nz, row-major per step */
    const double *G;    /* T x nz x q */
    const double *Wz;   /* dw x nz row-major */
    const double *wbar; /* (T+1) x dw */
    int32_t K;
    const double *alpha; /* K+1 */
    const double *mu;    /* K x (T*q) */
    int32_t swept;
} orc_mc_problem;

#define MC_MAXZ 16
#define MC_MAXQ 8
#define MC_MAXK 64

static int ws_point_hit(const orc_checker *CC, const double *p)
{
    if (CC->kind == 0) return orc_point_colliding_2d(CC->obs2d, p[0], p[1]);
    return !orc_box_point_free(CC->box_lo, CC->box_hi, CC->M, CC->d, p);
}
static int ws_segment_hit(const orc_checker *CC, const double *p, const double *q)
{
    if (CC->kind == 0) return orc_line_colliding_2d(CC->obs2d, p[0], p[1], q[0], q[1]);
    return !orc_box_segment_free(CC->box_lo, CC->box_hi, CC->M, CC->d, p, q);
}

/* one rollout: returns hit, *w_out = importance weight */
/* hn2[k] = 0.5 * |mu_k|^2, summed in stacked order */
void orc_mc_half_norms(const orc_mc_problem *P, double *hn2)
{
    for (int k = 0; k < P->K; ++k) {
        const double *m = P->mu + (size_t)k * P->T * P->q;
        double n2 = 0.0;
        for (int g = 0; g < P->T * P->q; ++g) n2 = n2 + m[g] * m[g];
        hn2[k] = 0.5 * n2;
    }
}

int orc_mc_rollout(const orc_mc_problem *P, const orc_checker *CC, const double *hn2, uint64_t seed, int64_t id,
                   double *w_out)
{
    const int T = P->T, nz = P->nz, q = P->q, dw = P->dw, K = P->K;
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t ctr[4] = { (uint32_t)(uint64_t)id, (uint32_t)((uint64_t)id >> 32), 0, 0 }, rnd[4];
    /* component choice */
    orc_philox4x32_10(ctr, key, rnd);
    double u = u53(rnd[0], rnd[1]);
    int comp = 0;
    double acc = P->alpha[0];
    while (comp < K && u >= acc) { ++comp; acc = acc + P->alpha[comp]; }
    const double *mu_c = comp > 0 ? P->mu + (size_t)(comp - 1) * T * q : NULL;

    double z[MC_MAXZ], zn[MC_MAXZ], eps[MC_MAXQ], dots[MC_MAXK], wprev[MC_MAXZ], wcur[MC_MAXZ];
    for (int i = 0; i < nz; ++i) z[i] = 0.0;
    for (int k = 0; k < K; ++k) dots[k] = 0.0;
    for (int i = 0; i < dw; ++i) wprev[i] = P->wbar[i];
    int hit = 0;
    double pair[2] = { 0, 0 };
    for (int t = 0; t < T; ++t) {
        for (int j = 0; j < q; ++j) {
            int g = t * q + j; /* index in the stacked noise; Box-Muller pair g/2, element g%2 */
            if ((g & 1) == 0) {
                ctr[2] = (uint32_t)(g >> 1); ctr[3] = 1;
                orc_philox4x32_10(ctr, key, rnd);
                double u1 = u53(rnd[0], rnd[1]), u2 = u53(rnd[2], rnd[3]);
                double rr = sqrt(-2.0 * orc_det_log(u1)), sn, cs;
                orc_det_sincos2pi(u2, &sn, &cs);
                pair[0] = rr * cs; pair[1] = rr * sn;
            }
            double xi = pair[g & 1];
            eps[j] = mu_c ? xi + mu_c[g] : xi;
        }
        for (int k = 0; k < K; ++k) { /* mu_k . eps, accumulated in stacked order */
            const double *m = P->mu + ((size_t)k * T + t) * q;
            double d = dots[k];
            for (int j = 0; j < q; ++j) d = d + m[j] * eps[j];
            dots[k] = d;
        }
        if (!hit) {
            const double *F = P->F + (size_t)t * nz * nz, *G = P->G + (size_t)t * nz * q;
            for (int i = 0; i < nz; ++i) {
                double a = 0.0;
                for (int j = 0; j < nz; ++j) a = a + F[i * nz + j] * z[j];
                for (int j = 0; j < q; ++j) a = a + G[i * q + j] * eps[j];
                zn[i] = a;
            }
            for (int i = 0; i < nz; ++i) z[i] = zn[i];
            for (int i = 0; i < dw; ++i) {
                double a = P->wbar[(size_t)(t + 1) * dw + i];
                for (int j = 0; j < nz; ++j) a = a + P->Wz[i * nz + j] * z[j];
                wcur[i] = a;
            }
            hit = P->swept ? ws_segment_hit(CC, wprev, wcur) : ws_point_hit(CC, wcur);
            for (int i = 0; i < dw; ++i) wprev[i] = wcur[i];
        }
    }
    double den = P->alpha[0];
    for (int k = 0; k < K; ++k) den = den + P->alpha[k + 1] * orc_det_exp(dots[k] - hn2[k]);
    *w_out = 1.0 / den;
    return hit;
}

/* sums over rollouts [first, first+n) in id order; optional per-rollout outputs */
void orc_mc_run(const orc_mc_problem *P, const orc_checker *CC, uint64_t seed, int64_t first, int64_t n,
                double *sums /* S1 S2 S0 */, int64_t *hits, uint8_t *hit_out, double *w_out)
{
    double s1 = 0, s2 = 0, s0 = 0, hn2[MC_MAXK];
    int64_t h = 0;
    orc_mc_half_norms(P, hn2);
    for (int64_t i = 0; i < n; ++i) {
        double w;
        int hit = orc_mc_rollout(P, CC, hn2, seed, first + i, &w);
        if (hit_out) hit_out[i] = (uint8_t)hit;
        if (w_out) w_out[i] = w;
        s0 += w;
        if (hit) { s1 += w; s2 += w * w; ++h; }
    }
    sums[0] = s1; sums[1] = s2; sums[2] = s0;
    *hits = h;
}
