/*
 * mp_oracle.h -- CPU oracle for the MotionPlanning.jl hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * (schmrlng/MotionPlanning.jl) arithmetic, in source operation order, IEEE
 * double, no FMA contraction (built with -ffp-contract=off).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (motionplanning.jl_b200, libmpb200.so) never does.
 *
 * Parity pinning: the reference holds NO golden vectors for this path
 * (test/runtests.jl:5 is `@test 1 == 1`) and Julia is not installed, so the
 * reference cannot be executed here.  The oracle is pinned by
 *   (i)  hand-derived known-answer cases on the reference's own obstacle
 *        fixtures (test/obstaclesets/2D.jl, ND.jl)  -> tests/test_oracle_*.py
 *   (ii) for the linear-quadratic steering cost, golden vectors produced by
 *        re-running the reference's SymPy construction
 *        (src/statespaces/linearquadratic.jl:126-157) with Python sympy
 *        -> tests/golden/ (+ the generating script).
 * Where neither exists (the Monte-Carlo estimator, which is absent from the
 * reference) the header of the file says "parity unpinned".
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).
 */
#ifndef MP_ORACLE_H
#define MP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- 2-D obstacle table (same flat layout the C ABI takes) ------------- */
/* kind: 0 = Circle   data: cx cy r xlo xhi ylo yhi                 (7)
 *       1 = Polygon  data: xlo xhi ylo yhi pts[2K] normals[2K] nextrema[2K]
 * gates: the AABBs of the enclosing Compound2D nodes (SAT2D.jl:82-96); a
 * shape is only tested when every ancestor gate passes (SAT2D.jl:129-132,
 * 158-161).  gate_parent[g] < g or -1.                                    */
typedef struct {
    int32_t n_gates;
    const int32_t *gate_parent;
    const double *gate_aabb; /* xlo xhi ylo yhi per gate */
    int32_t n_shapes;
    const int32_t *shape_kind;
    const int32_t *shape_gate;
    const int32_t *shape_off; /* n_shapes+1 offsets into data */
    const double *data;
    int32_t flags; /* bit0: use the intended (non-inverted) point-in-polygon test */
} orc_obs2d;

/* state space: bounds + state->workspace map (statespaces.jl:29-34,45-60) */
typedef struct {
    int32_t n;          /* state dimension */
    const double *lo;   /* n */
    const double *hi;   /* n */
    int32_t s2w_kind;   /* 0 Identity, 1 VectorView(inds), 2 OutputMatrix(C) */
    int32_t dw;         /* workspace dimension */
    const int32_t *inds;/* dw, 0-based (kind 1) */
    const double *C;    /* dw x n column-major (kind 2) */
} orc_space;

/* geom2d.c */
int orc_circle_build(double cx, double cy, double r, double *out7);
int orc_polygon_build(const double *pts_xy, int K, double *out /*4+6K*/);
int orc_point_colliding_2d(const orc_obs2d *O, double px, double py);
int orc_line_colliding_2d(const orc_obs2d *O, double vx, double vy, double wx, double wy);
void orc_points_free_2d(const orc_obs2d *O, const double *P_aos, int64_t n, uint8_t *out);
void orc_segments_free_2d(const orc_obs2d *O, const double *V_aos, const double *W_aos, int64_t n, uint8_t *out);

/* boxes.c */
int orc_box_point_free(const double *lo, const double *hi, int M, int d, const double *v);
int orc_box_segment_free(const double *lo, const double *hi, int M, int d, const double *v, const double *w);
void orc_points_free_boxes(const double *lo, const double *hi, int M, int d, const double *P_aos, int64_t n, uint8_t *out);
void orc_segments_free_boxes(const double *lo, const double *hi, int M, int d, const double *V_aos, const double *W_aos, int64_t n, uint8_t *out);

/* space.c : the (CC,SS) wrappers, statespaces.jl:150-158 */
typedef struct {
    int32_t kind;            /* 0 = PointRobot2D, 1 = PointRobotNDBoxes */
    const orc_obs2d *obs2d;
    const double *box_lo, *box_hi;
    int32_t M, d;
} orc_checker;
int orc_in_state_space(const orc_space *S, const double *v);
void orc_state2workspace(const orc_space *S, const double *v, double *w);
int orc_is_free_state(const orc_checker *CC, const orc_space *S, const double *v);
/* straight (Euclidean) edge: waypoints (v,w), geometric.jl:20 */
int orc_is_free_motion_straight(const orc_checker *CC, const orc_space *S, const double *v, const double *w, int64_t *count);
void orc_states_free(const orc_checker *CC, const orc_space *S, const double *P_aos, int64_t n, uint8_t *out);
void orc_edges_free_csc(const orc_checker *CC, const orc_space *S, const double *V_aos, int64_t N,
                        const int64_t *colptr, const int64_t *rowval, int64_t c0, int64_t c1,
                        uint8_t *out, int64_t *count);

/* rball.c : Euclidean r-ball (nearneighbors.jl:138-150,179-183; geometric.jl:4-6) */
/* brute truth.  pred: 0 -> s <= r*r (tree semantics), 1 -> sqrt(s) <= r (brute fallback).
 * Two-call protocol: pass rowval=NULL to get counts (colptr, 1-based, n_q+1). */
void orc_rball_brute(const double *V_aos, int64_t N, int d, double r, int pred,
                     int64_t q0, int64_t q1, int64_t *colptr, int64_t *rowval, double *nzval);
/* KD-tree variant (timing-faithful restatement of the TreeDistanceDS path) */
typedef struct orc_kdtree orc_kdtree;
orc_kdtree *orc_kdtree_build(const double *V_aos, int64_t N, int d, int leafsize);
void orc_kdtree_free(orc_kdtree *T);
/* one query, reference-style: returns malloc'ed arrays the caller frees (mimics the
 * per-query SparseVector allocation, nearneighbors.jl:179-183) */
int64_t orc_kdtree_inball(const orc_kdtree *T, int64_t v, double r, int64_t **inds, double **ds);
/* batch over [q0,q1): count pass (rowval NULL) or fill pass */
void orc_rball_kdtree(const orc_kdtree *T, double r, int64_t q0, int64_t q1,
                      int64_t *colptr, int64_t *rowval, double *nzval);

/* lq.c : double-integrator linear-quadratic steering (linearquadratic.jl:46-53,68-88,126-225) */
void orc_lq_steer(int d, const double *R, const double *x0, const double *x1, double r, double *cost, double *topt);
void orc_lq_cost_terms(int d, const double *R, const double *x0, const double *x1, double t, double *out3);
void orc_lq_state(int d, const double *x0, const double *x1, double t, double s, double *out);
void orc_lq_inball(const double *V, int64_t N, int d, const double *R, double r, int forwards, int64_t q0, int64_t q1,
                   int64_t *colptr, int64_t *rowval, double *nzval);
int orc_lq_is_free_motion(const orc_checker *CC, const orc_space *S, int d, const double *R, double r,
                          const double *v, const double *w, int64_t *count);
void orc_lq_edges_free_csc(const orc_checker *CC, const orc_space *S, int d, const double *R, double r,
                           const double *V, const int64_t *colptr, const int64_t *rowval, int64_t c0, int64_t c1,
                           uint8_t *out, int64_t *count);

/* lq_general.c : general linear-affine systems xdot = A x + B u + c, nilpotent A (linearquadratic.jl:94-225).
 * All matrices row-major.  Tables: Ak[k] = A^k / k!, dk[k] = Ak[k] c / (k+1), BRB = B R^-1 B',
 * Gp[p-1] = coefficient of t^p in G(t). */
#define ORC_LQG_MAXN 6
typedef struct {
    int n, np;
    double A[ORC_LQG_MAXN * ORC_LQG_MAXN], c[ORC_LQG_MAXN], BRB[ORC_LQG_MAXN * ORC_LQG_MAXN];
    double Ak[ORC_LQG_MAXN][ORC_LQG_MAXN * ORC_LQG_MAXN], dk[ORC_LQG_MAXN][ORC_LQG_MAXN];
    double Gp[2 * ORC_LQG_MAXN][ORC_LQG_MAXN * ORC_LQG_MAXN];
} orc_lqg;
int orc_lqg_setup(int n, int m, const double *A, const double *B, const double *c, const double *R, orc_lqg *S);
int orc_lqg_cost_terms(const orc_lqg *S, const double *x, const double *y, double t, double *out3);
void orc_lqg_steer(const orc_lqg *S, const double *x0, const double *x1, double r, double *cost, double *topt);
void orc_lqg_state(const orc_lqg *S, const double *x0, const double *x1, double t, double s, double *out);
void orc_lqg_inball(const orc_lqg *S, const double *V, int64_t N, double r, int forwards, int64_t q0, int64_t q1,
                    int64_t *colptr, int64_t *rowval, double *nzval);
int orc_lqg_is_free_motion(const orc_checker *CC, const orc_space *Sp, const orc_lqg *S, double r, const double *v,
                           const double *w, int64_t *count);
void orc_lqg_edges_free_csc(const orc_checker *CC, const orc_space *Sp, const orc_lqg *S, double r, const double *V,
                            const int64_t *colptr, const int64_t *rowval, int64_t c0, int64_t c1, uint8_t *out,
                            int64_t *count);

/* cars.c : chopped-metric car spaces (simplecars.jl; primitivetypes.jl:79-100; nearneighbors.jl:185-198).
 * kind 0 = ReedsSheppExact, 1 = DubinsExact; states (x, y, theta); segs = 5 x (duration, speed, curvature). */
double orc_mod2pi(double x);
void orc_det_sincos(double x, double *sn, double *cs);
double orc_det_sin(double x);
double orc_det_cos(double x);
double orc_det_atan2(double y, double x);
double orc_det_acos(double x);
double orc_car_steer(int kind, double rturn, double speed, const double *v, const double *w, int *nseg, double *segs);
void orc_car_propagate(const double *v, const double *u3, double *out);
double orc_car_chopped(int kind, double rturn, double chopval, const double *v, const double *w);
void orc_car_inball(const double *V, int64_t N, int kind, double rturn, double r, double chopval, int forwards,
                    int64_t q0, int64_t q1, int64_t *colptr, int64_t *rowval, double *nzval);
int orc_car_is_free_motion(const orc_checker *CC, const orc_space *S, int kind, double rturn, double speed,
                           const double *v, const double *w, int64_t *count);
void orc_car_edges_free_csc(const orc_checker *CC, const orc_space *S, int kind, double rturn, double speed,
                            const double *V, const int64_t *colptr, const int64_t *rowval, int64_t c0, int64_t c1,
                            uint8_t *out, int64_t *count);

/* closest.c : closest / closeR under a weight matrix (SAT2D.jl:208-285, boxesND.jl:61-86); W row-major */
#define ORC_CP_MAXD 4
void orc_closest_circle(const double *p, const double *rec, const double *W, double *d2, double *x);
void orc_closest_polygon(const double *p, const double *rec, int K, const double *W, double *d2, double *x);
void orc_closest_box(const double *p, const double *lo, const double *hi, int d, const double *W, double *d2, double *x);
int orc_close_points(const orc_checker *CC, const double *P, const double *Ws, int64_t n, int dw, double r2,
                     int32_t *count, double *d2_out, int32_t *shape_out, double *x_out, double *all_d2, double *all_x);

/* ---- Philox4x32-10 (mc.c) and batched free-state sampling (sample.c; sampling.jl:23-37) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_sample_candidate(const orc_space *S, uint64_t seed, int64_t c, double *x);
uint64_t orc_morton_key(const orc_space *S, const double *x);
int64_t orc_sample_free(const orc_checker *CC, const orc_space *S, int64_t N, uint64_t seed, int64_t max_candidates,
                        int order, double *V_aos, int64_t *candidates);

#ifdef __cplusplus
}
#endif
#endif
