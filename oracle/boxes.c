/*
 * boxes.c -- oracle restatement of the N-d point-robot-among-boxes checker.
 * TEST INFRASTRUCTURE ONLY (see mp_oracle.h).
 *
 * Follows src/collisioncheckers/boxesND.jl:42-56 and blend (src/utilities/utils.jl:41-51)
 * literally, including quirk Q2: the narrow phase tests one face per axis, never
 * checks lambda in [0,1], and divides by v_to_w[i] even when it is zero (IEEE
 * +-Inf/NaN propagate, comparisons with NaN are false).
 * lo/hi are box-major: lo[k*d + i].
 */
#include "mp_oracle.h"

#define MAXD 32

/* boxesND.jl:42 : is_free_state(v, BB) = @any [!(lo_i <= v_i <= hi_i)] */
static int box_point_free1(const double *lo, const double *hi, int d, const double *v)
{
    for (int i = 0; i < d; ++i)
        if (!(lo[i] <= v[i] && v[i] <= hi[i])) return 1;
    return 0;
}
/* boxesND.jl:43 : @all over boxes */
int orc_box_point_free(const double *lo, const double *hi, int M, int d, const double *v)
{
    for (int k = 0; k < M; ++k)
        if (!box_point_free1(lo + (int64_t)k * d, hi + (int64_t)k * d, d, v)) return 0;
    return 1;
}
/* boxesND.jl:44-45 */
static int broadphase_free(const double *lo, const double *hi, int d, const double *l, const double *h)
{
    for (int i = 0; i < d; ++i)
        if (hi[i] < l[i] || lo[i] > h[i]) return 1;
    return 0;
}
/* boxesND.jl:46-51 */
static int narrow_free(const double *lo, const double *hi, int d, const double *v, const double *w)
{
    double v_to_w[MAXD], lambdas[MAXD];
    for (int i = 0; i < d; ++i) {
        v_to_w[i] = w[i] - v[i];
        double corner = (v[i] < lo[i]) ? lo[i] : hi[i]; /* blend(map(<, v, lo), lo, hi) */
        lambdas[i] = (corner - v[i]) / v_to_w[i];
    }
    for (int i = 0; i < d; ++i) {
        int all = 1;
        for (int j = 0; j < d; ++j) {
            if (i == j) continue;
            double x = v[j] + v_to_w[j] * lambdas[i];
            if (!(lo[j] <= x && x <= hi[j])) { all = 0; break; }
        }
        if (all) return 0;
    }
    return 1;
}
/* boxesND.jl:52-56 */
int orc_box_segment_free(const double *lo, const double *hi, int M, int d, const double *v, const double *w)
{
    double bb_min[MAXD], bb_max[MAXD];
    for (int i = 0; i < d; ++i) { /* map(min, v, w): Julia min(x,y) = ifelse(y < x, y, x) for non-NaN */
        bb_min[i] = (w[i] < v[i]) ? w[i] : v[i];
        bb_max[i] = (v[i] < w[i]) ? w[i] : v[i];
    }
    for (int k = 0; k < M; ++k) {
        const double *l = lo + (int64_t)k * d, *h = hi + (int64_t)k * d;
        if (!(broadphase_free(l, h, d, bb_min, bb_max) || narrow_free(l, h, d, v, w))) return 0;
    }
    return 1;
}

void orc_points_free_boxes(const double *lo, const double *hi, int M, int d, const double *P, int64_t n, uint8_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)orc_box_point_free(lo, hi, M, d, P + i * d);
}
void orc_segments_free_boxes(const double *lo, const double *hi, int M, int d, const double *V, const double *W,
                             int64_t n, uint8_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)orc_box_segment_free(lo, hi, M, d, V + i * d, W + i * d);
}
