/*
 * space.c -- oracle restatement of the (CollisionChecker, StateSpace) validity
 * wrappers.  TEST INFRASTRUCTURE ONLY (see mp_oracle.h).
 *
 * Follows src/statespaces.jl:45-60 (state2workspace) and :150-158
 * (in_state_space, is_free_state, is_free_motion incl. quirk Q4: the last
 * waypoint is never bounds-checked), robots2D.jl:12-14, boxesND.jl:25-27.
 */
#include "mp_oracle.h"

#define MAXD 32

/* statespaces.jl:150 */
int orc_in_state_space(const orc_space *S, const double *v)
{
    for (int i = 0; i < S->n; ++i)
        if (!(S->lo[i] <= v[i] && v[i] <= S->hi[i])) return 0;
    return 1;
}
/* statespaces.jl:57-60 ; OutputMatrix: SMatrix*SVector, row i = sum_j C[i,j]*v[j], j ascending */
void orc_state2workspace(const orc_space *S, const double *v, double *w)
{
    if (S->s2w_kind == 0) {
        for (int i = 0; i < S->n; ++i) w[i] = v[i];
    } else if (S->s2w_kind == 1) {
        for (int i = 0; i < S->dw; ++i) w[i] = v[S->inds[i]];
    } else {
        for (int i = 0; i < S->dw; ++i) {
            double acc = S->C[i] * v[0];
            for (int j = 1; j < S->n; ++j) acc = acc + S->C[i + (int64_t)j * S->dw] * v[j];
            w[i] = acc;
        }
    }
}
static int ws_point_free(const orc_checker *CC, const double *p)
{
    if (CC->kind == 0) return !orc_point_colliding_2d(CC->obs2d, p[0], p[1]);     /* robots2D.jl:12 */
    return orc_box_point_free(CC->box_lo, CC->box_hi, CC->M, CC->d, p);          /* boxesND.jl:25 */
}
static int ws_segment_free(const orc_checker *CC, const double *p, const double *q, int64_t *count)
{
    if (count) *count += 1;                                                      /* robots2D.jl:13, boxesND.jl:26 */
    if (CC->kind == 0) return !orc_line_colliding_2d(CC->obs2d, p[0], p[1], q[0], q[1]);
    return orc_box_segment_free(CC->box_lo, CC->box_hi, CC->M, CC->d, p, q);
}
/* statespaces.jl:151-152 */
int orc_is_free_state(const orc_checker *CC, const orc_space *S, const double *v)
{
    double p[MAXD];
    if (!orc_in_state_space(S, v)) return 0;
    orc_state2workspace(S, v, p);
    return ws_point_free(CC, p);
}
/* statespaces.jl:153-158 with collision_waypoints = (v, w) (geometric.jl:20) */
int orc_is_free_motion_straight(const orc_checker *CC, const orc_space *S, const double *v, const double *w,
                                int64_t *count)
{
    double p[MAXD], q[MAXD];
    if (!orc_in_state_space(S, v)) return 0; /* only wps[1..end-1] are bounds-checked */
    orc_state2workspace(S, v, p);
    orc_state2workspace(S, w, q);
    return ws_segment_free(CC, p, q, count);
}
void orc_states_free(const orc_checker *CC, const orc_space *S, const double *P, int64_t n, uint8_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)orc_is_free_state(CC, S, P + i * S->n);
}
/* every stored entry (row y, column x) of a CSC neighbour table is the edge y -> x that
 * fmt.jl:75 would ask about: is_free_motion(V[y_min], V[x], CC, SS).  colptr/rowval 1-based,
 * colptr is relative to column c0 (length c1-c0+1, colptr[0] == 1). */
void orc_edges_free_csc(const orc_checker *CC, const orc_space *S, const double *V, int64_t N,
                        const int64_t *colptr, const int64_t *rowval, int64_t c0, int64_t c1,
                        uint8_t *out, int64_t *count)
{
    (void)N;
    int n = S->n;
    for (int64_t x = c0; x < c1; ++x)
        for (int64_t e = colptr[x - c0] - 1; e < colptr[x - c0 + 1] - 1; ++e) {
            int64_t y = rowval[e] - 1;
            out[e] = (uint8_t)orc_is_free_motion_straight(CC, S, V + y * n, V + x * n, count);
        }
}
