/*
 * rball.c -- oracle restatement of the Euclidean r-ball query.
 * TEST INFRASTRUCTURE ONLY (see mp_oracle.h).
 *
 * Truth (brute): src/nearneighbors.jl:138-150 with colwise(Euclidean, v, V)
 * (src/statespaces/geometric.jl:4-6): s_j = sum_{i=1..d} (V[v]_i - V[j]_i)^2 in
 * index order, no FMA; member iff j != v and (pred 0) s_j <= r*r -- the leaf test of
 * NearestNeighbors.jl's inrange, which is what the TreeDistanceDS path
 * (nearneighbors.jl:179-183) evaluates -- or (pred 1) sqrt(s_j) <= r, the brute
 * fallback's own comparison (:144).  Stored value sqrt(s_j); indices ascending.
 *
 * KD-tree variant: a timing-faithful restatement of the TreeDistanceDS path
 * (KDTree(V; reorder=false) build, geometric.jl:14; per query inrange -> sort ->
 * drop self by binary search -> gather V[inds] -> colwise distances -> freshly
 * allocated SparseVector, nearneighbors.jl:179-183).  NearestNeighbors.jl itself is
 * a third-party dependency absent from /root/reference (REQUIRE:1-9, unpinned,
 * late-2016 era); its published algorithm is an implicit balanced kd-tree with
 * leafsize 10, split on the widest dimension at the median, inrange pruned by the
 * incrementally maintained squared distance to the node's hyper-rectangle.
 * Membership here uses the same exact leaf test as the brute truth, and pruning
 * is made strictly conservative, so both variants return identical sets.
 */
#include "mp_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline double sqdist(const double *a, const double *b, int d)
{
    double t = a[0] - b[0];
    double s = t * t;
    for (int i = 1; i < d; ++i) { t = a[i] - b[i]; s = s + t * t; }
    return s;
}

void orc_rball_brute(const double *V, int64_t N, int d, double r, int pred, int64_t q0, int64_t q1,
                     int64_t *colptr, int64_t *rowval, double *nzval)
{
    double r2 = r * r;
    int count_only = (rowval == NULL);
    int64_t pos = 0;
    if (count_only) colptr[0] = 1;
    for (int64_t v = q0; v < q1; ++v) {
        const double *a = V + v * d;
        for (int64_t j = 0; j < N; ++j) {
            if (j == v) continue;
            double s = sqdist(a, V + j * d, d);
            int in = pred ? (sqrt(s) <= r) : (s <= r2);
            if (in) {
                if (!count_only) { rowval[pos] = j + 1; nzval[pos] = sqrt(s); }
                ++pos;
            }
        }
        if (count_only) colptr[v - q0 + 1] = pos + 1;
    }
}

/* ---- kd-tree ------------------------------------------------------------ */
struct orc_kdtree {
    const double *V;
    int64_t N;
    int d, leafsize;
    int64_t *idx;     /* permutation (reorder=false: data stays in place, accessed through idx) */
    int64_t n_nodes;
    int32_t *split_dim;
    double *split_val;
    int64_t *lo, *hi; /* idx range per node */
    int64_t *left, *right;
    double *bbox;     /* root bounding box: d mins then d maxs */
};

static int64_t build_rec(orc_kdtree *T, int64_t lo, int64_t hi, double *bmin, double *bmax, int64_t *next)
{
    int64_t me = (*next)++;
    T->lo[me] = lo; T->hi[me] = hi; T->left[me] = -1; T->right[me] = -1;
    if (hi - lo <= T->leafsize) { T->split_dim[me] = -1; return me; }
    int d = T->d, sd = 0;
    double best = -1;
    for (int i = 0; i < d; ++i) if (bmax[i] - bmin[i] > best) { best = bmax[i] - bmin[i]; sd = i; }
    /* median by nth_element (quickselect) on idx[lo..hi) */
    int64_t mid = lo + (hi - lo) / 2, a = lo, b = hi - 1;
    while (a < b) {
        double pv = T->V[T->idx[(a + b) / 2] * d + sd];
        int64_t i = a, j = b;
        while (i <= j) {
            while (T->V[T->idx[i] * d + sd] < pv) ++i;
            while (T->V[T->idx[j] * d + sd] > pv) --j;
            if (i <= j) { int64_t t = T->idx[i]; T->idx[i] = T->idx[j]; T->idx[j] = t; ++i; --j; }
        }
        if (mid <= j) b = j; else if (mid >= i) a = i; else break;
    }
    double sv = T->V[T->idx[mid] * d + sd];
    T->split_dim[me] = sd; T->split_val[me] = sv;
    double save = bmax[sd];
    bmax[sd] = sv;
    T->left[me] = build_rec(T, lo, mid, bmin, bmax, next);
    bmax[sd] = save;
    save = bmin[sd];
    bmin[sd] = sv;
    T->right[me] = build_rec(T, mid, hi, bmin, bmax, next);
    bmin[sd] = save;
    return me;
}

orc_kdtree *orc_kdtree_build(const double *V, int64_t N, int d, int leafsize)
{
    orc_kdtree *T = (orc_kdtree *)calloc(1, sizeof(*T));
    T->V = V; T->N = N; T->d = d; T->leafsize = leafsize > 0 ? leafsize : 10;
    T->idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    for (int64_t i = 0; i < N; ++i) T->idx[i] = i;
    int64_t cap = 2 * (N / T->leafsize + 2) * 2 + 8;
    T->split_dim = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    T->split_val = (double *)malloc(sizeof(double) * (size_t)cap);
    T->lo = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    T->hi = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    T->left = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    T->right = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    T->bbox = (double *)malloc(sizeof(double) * 2 * (size_t)d);
    for (int i = 0; i < d; ++i) { T->bbox[i] = INFINITY; T->bbox[d + i] = -INFINITY; }
    for (int64_t j = 0; j < N; ++j)
        for (int i = 0; i < d; ++i) {
            double x = V[j * d + i];
            if (x < T->bbox[i]) T->bbox[i] = x;
            if (x > T->bbox[d + i]) T->bbox[d + i] = x;
        }
    double bmin[64], bmax[64];
    memcpy(bmin, T->bbox, sizeof(double) * (size_t)d);
    memcpy(bmax, T->bbox + d, sizeof(double) * (size_t)d);
    int64_t next = 0;
    if (N > 0) build_rec(T, 0, N, bmin, bmax, &next);
    T->n_nodes = next;
    return T;
}
void orc_kdtree_free(orc_kdtree *T)
{
    if (!T) return;
    free(T->idx); free(T->split_dim); free(T->split_val); free(T->lo); free(T->hi);
    free(T->left); free(T->right); free(T->bbox); free(T);
}

typedef struct { int64_t *a; int64_t n, cap; } ivec;
static void ipush(ivec *v, int64_t x)
{
    if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 32; v->a = (int64_t *)realloc(v->a, sizeof(int64_t) * (size_t)v->cap); }
    v->a[v->n++] = x;
}
static void inrange_rec(const orc_kdtree *T, int64_t node, const double *q, double r2, double r2_prune,
                        double min_d2, double *off, ivec *out)
{
    if (min_d2 > r2_prune) return;
    int sd = T->split_dim[node];
    if (sd < 0) {
        for (int64_t k = T->lo[node]; k < T->hi[node]; ++k) {
            int64_t j = T->idx[k];
            if (sqdist(q, T->V + j * T->d, T->d) <= r2) ipush(out, j);
        }
        return;
    }
    double diff = q[sd] - T->split_val[node];
    int64_t close = diff < 0 ? T->left[node] : T->right[node];
    int64_t far = diff < 0 ? T->right[node] : T->left[node];
    inrange_rec(T, close, q, r2, r2_prune, min_d2, off, out);
    double old = off[sd];
    double nd2 = min_d2 - old * old + diff * diff;
    if (fabs(diff) > fabs(old)) {
        off[sd] = diff;
        inrange_rec(T, far, q, r2, r2_prune, nd2, off, out);
        off[sd] = old;
    } else {
        inrange_rec(T, far, q, r2, r2_prune, min_d2, off, out);
    }
}
static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}
/* nearneighbors.jl:179-183 for one v */
int64_t orc_kdtree_inball(const orc_kdtree *T, int64_t v, double r, int64_t **inds_out, double **ds_out)
{
    int d = T->d;
    const double *q = T->V + v * d;
    double off[64];
    double min_d2 = 0;
    for (int i = 0; i < d; ++i) { /* distance from q to the root box (0 for members of the set) */
        double o = 0;
        if (q[i] < T->bbox[i]) o = T->bbox[i] - q[i];
        else if (q[i] > T->bbox[d + i]) o = q[i] - T->bbox[d + i];
        off[i] = o; min_d2 += o * o;
    }
    ivec hits = {0, 0, 0};
    double r2 = r * r;
    if (T->N > 0) inrange_rec(T, 0, q, r2, r2 * (1 + 1e-9) + 1e-300, min_d2, off, &hits);
    qsort(hits.a, (size_t)hits.n, sizeof(int64_t), cmp_i64);            /* inrange(..., true) sorts */
    int64_t lo = 0, hi = hits.n;                                         /* searchsortedfirst(inds, v) */
    while (lo < hi) { int64_t m = (lo + hi) / 2; if (hits.a[m] < v) lo = m + 1; else hi = m; }
    if (lo < hits.n) { memmove(hits.a + lo, hits.a + lo + 1, sizeof(int64_t) * (size_t)(hits.n - lo - 1)); hits.n--; } /* deleteat! */
    int64_t k = hits.n;
    double *G = (double *)malloc(sizeof(double) * (size_t)(k * d + 1)); /* V[inds] gather copy */
    for (int64_t e = 0; e < k; ++e) memcpy(G + e * d, T->V + hits.a[e] * d, sizeof(double) * (size_t)d);
    double *ds = (double *)malloc(sizeof(double) * (size_t)(k + 1));
    for (int64_t e = 0; e < k; ++e) ds[e] = sqrt(sqdist(q, G + e * d, d)); /* colwise(Euclidean, V[v], V[inds]) */
    free(G);
    if (!hits.a) hits.a = (int64_t *)malloc(sizeof(int64_t));
    *inds_out = hits.a; *ds_out = ds;
    return k;
}
void orc_rball_kdtree(const orc_kdtree *T, double r, int64_t q0, int64_t q1, int64_t *colptr, int64_t *rowval,
                      double *nzval)
{
    int count_only = (rowval == NULL);
    int64_t pos = 0;
    if (count_only) colptr[0] = 1;
    for (int64_t v = q0; v < q1; ++v) {
        int64_t *inds; double *ds;
        int64_t k = orc_kdtree_inball(T, v, r, &inds, &ds);
        if (!count_only)
            for (int64_t e = 0; e < k; ++e) { rowval[pos + e] = inds[e] + 1; nzval[pos + e] = ds[e]; }
        pos += k;
        if (count_only) colptr[v - q0 + 1] = pos + 1;
        free(inds); free(ds);
    }
}
