/*
 * mpb200.h -- C ABI of libmpb200.so, the B200 (sm_100a) back-end for the
 * data-parallel hot path of schmrlng/MotionPlanning.jl.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * A Julia maintainer binds these with `ccall((:sym, "libmpb200"), Cint, ...)`
 * (see INTEGRATION.md and julia/MotionPlanningB200.jl); the Python host mirror
 * in motionplanning.jl_b200/ binds them with ctypes.
 *
 * Conventions
 *  - every function returns 0 on success, a negative MPB200_E* code otherwise;
 *    mpb200_last_error() returns the thread-local message.  No CPU fallback:
 *    without a usable sm_100 GPU every compute entry point fails.
 *  - host buffers are caller-owned (Julia GC / numpy); the library copies inputs
 *    at *_create and writes outputs only between call and return.  Calls that hand
 *    data back wait for it; calls whose result pointers are all NULL only enqueue
 *    work on the launching stream (noted at each entry point).
 *  - samples cross as the reference stores them: Vector{SVector{d,Float64}} ==
 *    column-major d x N Float64 (primitivetypes.jl:21-23, statevec2mat).
 *  - neighbour tables cross as Julia SparseMatrixCSC{Float64,Int64} fields:
 *    colptr (ncols+1), rowval (nnz, ascending within a column), nzval (nnz),
 *    all indices 1-based Int64 (nearneighbors.jl:23-28, ImmutableNNC.D).
 *  - validity tables cross in Julia BitVector chunk layout: element k (0-based)
 *    is bit (k & 63) of UInt64 word (k >> 6); unused high bits are zero.
 *  - one process drives one GPU (mpb200_init(device)); multi-GPU sharding is by
 *    query (column) range, see mpb200_samples_set_query_range.
 */
#ifndef MPB200_H
#define MPB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPB200_OK 0
#define MPB200_EARG (-1)    /* bad argument */
#define MPB200_ECUDA (-2)   /* CUDA runtime error / no usable GPU */
#define MPB200_ESTATE (-3)  /* call out of order (e.g. fetch before build) */
#define MPB200_ENOMEM (-4)

typedef struct mpb200_samples mpb200_samples;     /* device copy of a sample set        */
typedef struct mpb200_table mpb200_table;         /* device-resident CSC neighbour table */
typedef struct mpb200_obstacles mpb200_obstacles; /* device copy of an obstacle set      */
typedef struct mpb200_lq mpb200_lq;               /* linear-quadratic 2BVP               */

/* ---- runtime ----------------------------------------------------------- */
/* Select the CUDA device for this process and create the library stream.
 * Fails (MPB200_ECUDA) when there is no GPU or it is not compute capability 10.x. */
int mpb200_init(int device);
void mpb200_shutdown(void);
const char *mpb200_last_error(void);
int mpb200_version(void);
/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the
 * library's own; pass NULL to return to the library stream. */
int mpb200_set_stream(void *cuda_stream);
int mpb200_synchronize(void);
/* Device arrays released by the library (destroyed tables / sample sets, outgrown buffers) are parked
 * in a free list and reused by later calls instead of going through cudaFree / cudaMalloc (milliseconds
 * each, device-synchronising).  This returns every parked block to the driver. */
int mpb200_release_cached(void);
/* Pinned host memory for result buffers (fast D2H); plain malloc'ed buffers work too. */
int mpb200_host_alloc(uint64_t bytes, void **out);
int mpb200_host_free(void *p);
/* Number of kernels this library has launched since mpb200_init (bench "gpu_launches"). */
int64_t mpb200_launch_count(void);
/* Device time of the last entry point, milliseconds, measured with CUDA events on the
 * launching stream; phase 0 = whole call, 1.. = entry-point specific.
 * The events are recorded inside the call and read back lazily here (this waits for the
 * recorded work), so entry points that return nothing from the device never block.
 * mpb200_last_ms_of keeps one record per kind of operation, so the three calls of a
 * planning step (table build, point validity, edge validity) can be issued back to back
 * and timed afterwards. */
#define MPB200_OP_TABLE 0  /* mpb200_inball_build / mpb200_lq_inball_build */
#define MPB200_OP_POINTS 1 /* mpb200_points_free */
#define MPB200_OP_EDGES 2  /* mpb200_edges_free / mpb200_lq_edges_free */
#define MPB200_OP_OTHER 3  /* mpb200_mc_collision_probability */
double mpb200_last_ms(int phase);
double mpb200_last_ms_of(int op, int phase);

/* ---- sample sets: replaces MetricNN/QuasiMetricNN.V (nearneighbors.jl:62-92) ---- */
int mpb200_samples_create(const double *V_aos, int64_t N, int d, mpb200_samples **out);
int mpb200_samples_destroy(mpb200_samples *s);
/* Restrict this process to query columns [q0, q1) (0-based, half-open) of the table:
 * the multi-GPU shard.  Default is [0, N). */
int mpb200_samples_set_query_range(mpb200_samples *s, int64_t q0, int64_t q1);

/* ---- Euclidean r-ball table --------------------------------------------------
 * Replaces helper_data_structures(V, ::Euclidean) (geometric.jl:14-15: KDTree build)
 * plus inball(V, dist, ::TreeDistanceDS, v, r) for every v in the query range
 * (nearneighbors.jl:179-183), i.e. it produces the ImmutableNNC.D matrix that
 * inball!(NN{..,ImmutableNNC}, v, r) = viewcol(D, v) serves (nearneighbors.jl:128).
 * Member iff j != v and sum_i (V[v]_i - V[j]_i)^2 <= r*r (index order, no FMA);
 * stored value sqrt of the same sum; rows ascending.
 * If *table is non-NULL on entry its device buffers are reused. */
int mpb200_inball_build(mpb200_samples *s, double r, mpb200_table **table, int64_t *nnz);
/* Copy a table to caller buffers; any pointer may be NULL to skip that array.
 * colptr has (q1-q0)+1 entries, relative to the shard (colptr[0] == 1). */
int mpb200_table_fetch(const mpb200_table *t, int64_t *colptr, int64_t *rowval, double *nzval);
int mpb200_table_nnz(const mpb200_table *t, int64_t *nnz, int64_t *ncols);
/* Device pointers of a table's arrays (colptr int64[ncols+1], rowval int64[nnz], nzval f64[nnz],
 * edge_bits uint64[ceil(nnz/64)] = last mpb200_edges_free / mpb200_lq_edges_free result or NULL),
 * valid until the next build on / destroy of the handle: lets the caller hand shards to NCCL
 * (all-gather across the GPUs of a box) without a host round trip. */
int mpb200_table_device_view(const mpb200_table *t, void **colptr, void **rowval, void **nzval, void **edge_bits);
int mpb200_table_destroy(mpb200_table *t);

/* ---- obstacle sets ------------------------------------------------------------ */
/* 2-D compound of circles and convex polygons, PRECOMPUTED on the host exactly as the
 * reference constructors do (SAT2D.jl:12-98): the device never re-derives normals.
 * shape_kind 0 = Circle  (data: cx cy r xlo xhi ylo yhi)
 *            1 = Polygon (data: xlo xhi ylo yhi, K points, K unit normals, K nextrema)
 * gates are the AABBs of enclosing Compound2D nodes (parent index or -1).
 * flags bit0: use the intended point-in-polygon test instead of the reference's
 * inverted one (SAT2D.jl:124-127). */
typedef struct {
    int32_t n_gates;
    const int32_t *gate_parent;
    const double *gate_aabb;
    int32_t n_shapes;
    const int32_t *shape_kind;
    const int32_t *shape_gate;
    const int32_t *shape_off;
    const double *data;
    int32_t flags;
} mpb200_obstacles2d_desc;
/* replaces PointRobot2D(obstacles) (robots2D.jl:5-10) */
int mpb200_obstacles2d_create(const mpb200_obstacles2d_desc *desc, mpb200_obstacles **out);
/* replaces PointRobotNDBoxes(boxes) (boxesND.jl:15-21); lo/hi box-major M x d */
int mpb200_boxes_create(const double *lo, const double *hi, int M, int d, mpb200_obstacles **out);
int mpb200_obstacles_destroy(mpb200_obstacles *o);

/* BoundedStateSpace bounds + State2Workspace (statespaces.jl:29-34,45-60) */
typedef struct {
    int32_t n;           /* state dimension */
    const double *lo;    /* n */
    const double *hi;    /* n */
    int32_t s2w_kind;    /* 0 Identity, 1 VectorView(inds), 2 OutputMatrix(C) */
    int32_t dw;          /* workspace dimension */
    const int32_t *inds; /* dw, 0-based (kind 1) */
    const double *C;     /* dw x n column-major (kind 2) */
} mpb200_space_desc;

/* ---- batched validity ---------------------------------------------------------- */
/* F[i] = is_free_state(V[i], CC, SS) for the samples of the query range [q0, q1) -- all N by
 * default (fmt.jl:31-36, sampling.jl:25; statespaces.jl:151-152; robots2D.jl:12;
 * boxesND.jl:42-43).  bits: ceil((q1-q0)/64) words, bit k = sample q0 + k.
 * bitchunks may be NULL: the call then only enqueues the work (no wait); the bits stay on the
 * device, ordered on the launching stream before every later call. */
int mpb200_points_free(const mpb200_samples *s, const mpb200_obstacles *o, const mpb200_space_desc *ss,
                       uint64_t *bitchunks);
/* For every stored entry (row y, column x) of a table, in storage order:
 * is_free_motion(V[y], V[x], CC, SS) with straight-line waypoints (fmt.jl:75;
 * statespaces.jl:153-158; geometric.jl:20; robots2D.jl:13-14; boxesND.jl:52-56).
 * bits: ceil(nnz/64) words.  *checks receives the number of segment checks that
 * CC.count would have been incremented by.  With bitchunks == NULL and checks == NULL the call
 * only enqueues the work (no wait); read the bits later with mpb200_table_fetch_edge_bits. */
int mpb200_edges_free(const mpb200_samples *s, const mpb200_table *t, const mpb200_obstacles *o,
                      const mpb200_space_desc *ss, uint64_t *bitchunks, int64_t *checks);
/* Convenience for config "batched r-ball neighbours + edge validity precompute": build the
 * Euclidean r-ball table and the validity of every stored edge with one call (= mpb200_inball_build
 * followed by mpb200_edges_free); the edge bits stay on the device with the table
 * (mpb200_table_fetch_edge_bits / mpb200_table_device_view). */
int mpb200_inball_build_checked(mpb200_samples *s, double r, const mpb200_obstacles *o, const mpb200_space_desc *ss,
                                mpb200_table **table, int64_t *nnz, int64_t *checks);
int mpb200_table_fetch_edge_bits(const mpb200_table *t, uint64_t *bitchunks);
/* State-level batches for callers that hold states, not indices (sampling.jl:25,
 * postprocessors.jl:11,21): n states / n straight segments, AoS d x n; out: 1 byte each. */
int mpb200_states_free(const double *v_aos, int64_t n, int d, const mpb200_obstacles *o,
                       const mpb200_space_desc *ss, uint8_t *out);
int mpb200_segments_free(const double *v_aos, const double *w_aos, int64_t n, int d, const mpb200_obstacles *o,
                         const mpb200_space_desc *ss, uint8_t *out);

/* Batched sample_free!(P, N) (sampling.jl:23-37): draws states uniformly in the state bounds
 * (sample_space, statespaces.jl) and keeps the first N for which is_free_state(v, CC, SS) holds -- generated,
 * tested and compacted on the device, so the sample set never crosses PCIe on its way in.  The reference draws
 * from Julia's global RNG, which cannot be reproduced; the candidate stream here is specified in
 * oracle/sample.c (Philox4x32-10 keyed by seed, counter = candidate number) and is independent of launch
 * geometry: the same (obstacles, space, N, seed) always yields the same samples, in candidate order.
 * order = MPB200_ORDER_MORTON numbers the accepted samples along a Z-order curve over the first min(n, 3)
 * coordinates (stable: ties keep candidate order) instead of in the order they were drawn: FMT* does not care how
 * i.i.d. samples are numbered, and with a spatially coherent numbering the neighbour-table and validity kernels
 * run about 12% faster (their column writes and gathers become local).
 * V_host (N x n, may be NULL) receives a host copy; *candidates the number of candidates consumed. */
#define MPB200_ORDER_CANDIDATE 0
#define MPB200_ORDER_MORTON 1
int mpb200_sample_free(const mpb200_obstacles *o, const mpb200_space_desc *ss, int64_t N, uint64_t seed, int32_t order,
                       mpb200_samples **out, double *V_host, int64_t *candidates);

/* ---- linear-quadratic steering cost ("ControlNN") --------------------------------
 * Replaces LinearQuadratic(A, B, c, R) / LinearQuadratic2BVP (linearquadratic.jl:6-39,126-157) for
 * xdot = A x + B u + c, cost int (1 + u'Ru).  A (n x n), B (n x m), R (m x m) column-major, c (n); n, m <= 6.
 * As in the reference, A must be nilpotent (expAt, :94-98); R symmetric positive definite; (A, B)
 * controllable.  Two evaluation paths, chosen here:
 *   - the DoubleIntegrator family (:46-53: n = 2m, A = [0 I; 0 0], B = [0; I], c = 0, m = 1..3) uses the
 *     closed form cost(t) = t + alpha/t^3 - beta/t^2 + gamma/t (csrc/lq.cu);
 *   - every other system (drift c != 0, integrator chains, general B) evaluates G(t), xbar(t) and the
 *     cost derivatives numerically from per-system tables (csrc/lq_general.cu; spec in oracle/lq_general.c,
 *     pinned to the reference's SymPy construction by tests/golden/lq_general.json).
 * Anything the reference rejects returns MPB200_EARG with the reference's message. */
int mpb200_lq_create(const double *A, const double *B, const double *c, const double *R, int n, int m,
                     mpb200_lq **out);
int mpb200_lq_destroy(mpb200_lq *lq);
/* Replaces helper_data_structures(V, ::LinearQuadratic) (linearquadratic.jl:68-77): the batched
 * steer_pairwise (:196-225: dcost(r) > 0 prefilter, safeguarded Newton :175-190, cost <= r)
 * plus the inball filter (nearneighbors.jl:165-177), for every column of the query range.
 * tableF column v: { j != v : cost(V[v] -> V[j]) <= r }  (DSF = Dmat', served by inballF!)
 * tableB column v: { j != v : cost(V[j] -> V[v]) <= r }  (DSB = Dmat , served by inballB!)
 * stored value = the cost; rows ascending.  Non-NULL *table handles are reused. */
int mpb200_lq_inball_build(mpb200_samples *s, const mpb200_lq *lq, double r, mpb200_table **tableF,
                           mpb200_table **tableB, int64_t *nnzF, int64_t *nnzB);
/* steer(L, v, w, r) = (cost, t*) for n explicit pairs (linearquadratic.jl:191-195); AoS 2m x n */
int mpb200_lq_steer(const mpb200_lq *lq, const double *v_aos, const double *w_aos, int64_t n, double r,
                    double *cost, double *topt);
/* is_free_motion(V[y], V[x], CC, SS) for every stored entry (row y, column x) of a table, with
 * collision_waypoints(::LinearQuadratic) = 5 states along the optimal trajectory
 * (linearquadratic.jl:85-88; statespaces.jl:153-158).  bits/checks as mpb200_edges_free. */
int mpb200_lq_edges_free(const mpb200_samples *s, const mpb200_table *t, const mpb200_lq *lq, double r,
                         const mpb200_obstacles *o, const mpb200_space_desc *ss, uint64_t *bitchunks,
                         int64_t *checks);
/* the same for n explicit state pairs (fmt.jl:75 called with states) */
int mpb200_lq_motions_free(const mpb200_lq *lq, double r, const double *v_aos, const double *w_aos, int64_t n,
                           const mpb200_obstacles *o, const mpb200_space_desc *ss, uint8_t *out, int64_t *checks);

/* ---- chopped-metric car spaces (Dubins / Reeds-Shepp) --------------------------------------------
 * Replaces, for SE2 states (x, y, theta; AoS 3 x N):
 *   ReedsSheppExact / DubinsExact and their evaluate = reedsshepp / dubins (simplecars.jl:5-27, 196-213, 266-364),
 *   ChoppedMetric / ChoppedQuasiMetric evaluation (primitivetypes.jl:79-100),
 *   helper_data_structures(V, ::Chopped...{ReedsSheppExact|DubinsExact}) = KD-tree over (x, y) (simplecars.jl:42-52)
 *   + inball(V, ::ChoppedPreMetric, ::TreeDistanceDS, v, r, forwards) (nearneighbors.jl:185-198),
 *   steering_control / propagate / collision_waypoints (simplecars.jl:55-82) behind is_free_motion
 *   (statespaces.jl:134-142, 153-158).
 * sin / cos / atan2 / acos are fixed polynomial routines (specified in oracle/cars.c; the reference calls openlibm:
 * parity unpinned there); everything else follows the reference's operation order. */
#define MPB200_CAR_REEDS_SHEPP 0
#define MPB200_CAR_DUBINS 1
/* tableF column v: { i != v : (x,y) within r, chopped d(V[v] -> V[i]) <= r }, stored value = the path length;
 * tableB column v: the same with d(V[i] -> V[v]) (Dubins; pass tableB = NULL for the symmetric Reeds-Shepp
 * metric, whose MetricNN serves inballF! and inballB! from the forward table).  chopval = the metric's chop value
 * (setup_steering sets it to r).  s must hold d = 3 states; the query range of s is honoured. */
int mpb200_car_inball_build(mpb200_samples *s, int32_t kind, double turning_radius, double r, double chopval,
                            mpb200_table **tableF, mpb200_table **tableB, int64_t *nnzF, int64_t *nnzB);
/* entries of the (x, y) candidate table behind the last mpb200_car_inball_build on s = pairs whose exact metric was
 * evaluated (per direction); measurement aid */
int mpb200_car_last_candidates(const mpb200_samples *s, int64_t *pairs);
/* (cost, steering_control) for n explicit pairs: segments = n x 5 x (duration, signed speed, signed curvature),
 * nseg[i] of them valid (3 for Dubins, 3..5 for Reeds-Shepp) */
int mpb200_car_steer(int32_t kind, double turning_radius, double speed, const double *v_aos, const double *w_aos,
                     int64_t n, double *cost, int32_t *nseg, double *segments);
/* is_free_motion(V[y], V[x], CC, SS) for every stored entry (row y, column x) of a table: arc waypoints every pi/12
 * of positive heading change (simplecars.jl:71-82), each but the last bounds-checked, consecutive ones swept. */
int mpb200_car_edges_free(const mpb200_samples *s, const mpb200_table *t, int32_t kind, double turning_radius,
                          double speed, const mpb200_obstacles *o, const mpb200_space_desc *ss, uint64_t *bitchunks,
                          int64_t *checks);
int mpb200_car_motions_free(int32_t kind, double turning_radius, double speed, const double *v_aos,
                            const double *w_aos, int64_t n, const mpb200_obstacles *o, const mpb200_space_desc *ss,
                            uint8_t *out, int64_t *checks);

/* ---- Monte-Carlo trajectory collision probability -----------------------------------
 * NOT in the reference (only paper links, README.md:9-10, and the helper geometry
 * closest/closeR, SAT2D.jl:208-285, boxesND.jl:61-86): specified in SURVEY.md section 11 /
 * DESIGN.md.  Closed loop z' = F_t z + G_t eps_t (z_0 = 0), workspace w_t = wbar_t + Wz z_t,
 * hit = some w_t (swept: some segment w_{t-1}->w_t) collides with the obstacle set, defensive
 * mixture proposal over the stacked noise, Philox4x32-10 keyed by (seed, rollout id).
 * All matrices row-major. */
typedef struct {
    int32_t T;           /* steps */
    int32_t nz;          /* closed-loop state dimension */
    int32_t q;           /* noise dimension per step */
    int32_t dw;          /* workspace dimension */
    const double *F;     /* T x nz x nz */
    const double *G;     /* T x nz x q */
    const double *Wz;    /* dw x nz */
    const double *wbar;  /* (T+1) x dw nominal workspace points (index 0 = start, never tested as a point) */
    int32_t K;           /* shifted mixture components (besides the nominal one) */
    const double *alpha; /* K+1 mixture weights, alpha[0] = nominal, sum 1 */
    const double *mu;    /* K x (T*q) mean shifts of the stacked noise */
    int32_t swept;       /* 0: point test per step, 1: segment test between consecutive steps */
} mpb200_mc_problem;
typedef struct {
    double s1;    /* sum of w * hit          */
    double s2;    /* sum of (w * hit)^2      */
    double s0;    /* sum of w (diagnostic: -> n) */
    int64_t n;    /* rollouts                */
    int64_t hits; /* unweighted hit count    */
} mpb200_mc_result;
/* Rollout ids [first, first + n): the shard of one GPU; sums of shards add up to the whole.
 * hit_out (n bytes) / w_out (n doubles) are optional per-rollout outputs. */
int mpb200_mc_collision_probability(const mpb200_mc_problem *p, const mpb200_obstacles *o, uint64_t seed,
                                    int64_t first, int64_t n, mpb200_mc_result *out, uint8_t *hit_out,
                                    double *w_out);

/* ---- k-nearest connections ---------------------------------------------------------------------
 * The reference exports knn / knnF / knnB / mutualknn* (nearneighbors.jl:9-11) and FMT*'s connections = :K branch
 * calls mutualknnF! / knnB! (fmt.jl:6,17-19,70,72), but defines none of them (`:K` throws there).  Specification
 * implemented here (parity unpinned; restated in oracle/oracle.py):
 *   knn(v, k): the k samples j != v with the smallest stored value, ties towards the smaller index, returned like
 *   every neighbourhood (ascending indices + values);  mutualknnF(v, k) = knnF(v, k) U { w : v in knnB(w, k) }.
 * Both are table operations on top of ANY neighbour table (Euclidean or steering cost):
 *   mpb200_table_knn keeps the k best entries of every column of t; columns of t with fewer than k entries are kept
 *   whole and counted in *short_cols -- rebuild t with a larger radius until it is 0 and the result is the k-NN table;
 *   mpb200_table_union_transpose forms out[:, v] = a[:, v] U { w : v in b[:, w] } (values from a, else from b;
 *   full-range tables over the same sample set).  Columns are limited to 8192 entries.
 * Non-NULL *out handles are reused.  Edge validity works on the results like on any table. */
int mpb200_table_knn(const mpb200_table *t, int k, mpb200_table **out, int64_t *nnz, int64_t *short_cols);
/* only the question "does every column of t hold at least k entries?": *short_cols = columns that do not */
int mpb200_table_short_columns(const mpb200_table *t, int k, int64_t *short_cols);
int mpb200_table_union_transpose(const mpb200_table *a, const mpb200_table *b, mpb200_table **out, int64_t *nnz);

/* ---- closest obstacle points under a weight matrix (Monte-Carlo proposal geometry) ----------
 * Replaces closest(p, shape, W) / closeR(p, CC, W, r2) (SAT2D.jl:208-285; boxesND.jl:61-86; robots2D.jl:25-26)
 * for n query points at once, each with its own SPD weight matrix W_i (dw x dw row-major; dw = 2 for 2-D
 * obstacle sets, = box dimension <= 4 for box lists).  S = number of basic shapes (circles + polygons, or boxes).
 * Per point: count[i] shapes lie closer than r2 (squared W-distance); d2 / shape / x list them in ascending
 * d2 (ties in shape order), capacity S per point: d2[i*S + k], shape[i*S + k], x[(i*S + k)*dw ..].
 * all_d2 / all_x (optional, n x S [x dw]) receive closest() for every basic shape, unsorted. */
int mpb200_close_points(const mpb200_obstacles *o, const double *p_aos, const double *W, int64_t n, int dw, double r2,
                        int32_t *count, double *d2, int32_t *shape, double *x, double *all_d2, double *all_x);

/* ---- multi-GPU exchange by direct peer stores (one process per GPU, one box) ----------
 * The reference has no distributed layer; this is the exchange step of the query-range-sharded
 * precompute: after a rank has built its table shard and the validity of its
 * edges, ONE call stores the shard's int32 column lengths and validity words into slot `rank` of
 * every peer's receive buffer over NVLink (CUDA IPC mapped memory, no NCCL, no staging copies) and
 * runs a device-side flag barrier, all stream-ordered and asynchronous.
 *   create  -> allocates this rank's receive buffer (2 alternating sets x world slots) and returns
 *              its 64-byte CUDA IPC handle; the caller exchanges the handles of all ranks by any
 *              means (torch.distributed all_gather, MPI, a file) and hands them to connect.
 *   push    -> enqueue pack + peer stores + barrier for the current contents of `t`
 *              (needs mpb200_edges_free / mpb200_lq_edges_free on it first).
 *   view    -> device pointer of the newest complete set: world slots of slot_bytes each;
 *              slot g = [int64 ncols, int64 nnz, int64 epoch, pad to 64 B | int32 counts at
 *              counts_off | uint64 validity words at words_off].  With status != NULL the call
 *              waits for the enqueued pushes and returns 0, or 1 + the rank that never arrived.
 * A set stays valid until this rank's second-next push. */
typedef struct mpb200_xchg mpb200_xchg;
#define MPB200_IPC_HANDLE_BYTES 64
#define MPB200_XCHG_MAX_WORLD 16
int mpb200_xchg_create(int rank, int world, int64_t max_ncols, int64_t word_cap, mpb200_xchg **out, void *ipc_handle);
int mpb200_xchg_connect(mpb200_xchg *x, const void *handles /* world x 64 bytes, rank order */);
/* Optional: tie a table handle to the exchange.  Every later mpb200_inball_build on it starts sending the column
 * lengths of the coming epoch right after its count scan, on a side stream underneath the fill and validity
 * kernels; the following mpb200_xchg_push then only sends the validity words.  x = NULL detaches. */
int mpb200_xchg_attach(mpb200_xchg *x, mpb200_table *t);
int mpb200_xchg_push(mpb200_xchg *x, const mpb200_table *t);
int mpb200_xchg_view(const mpb200_xchg *x, void **recv, int64_t *slot_bytes, int64_t *counts_off, int64_t *words_off,
                     int64_t *status);
int mpb200_xchg_destroy(mpb200_xchg *x);

/* ---- measured arithmetic-pipe peaks (roofline denominators for the FP64 / FP32 kernels) ----
 * Runs a register-resident instruction stream on every SM and returns operations per second
 * (an FMA counts as two), CUDA-event timed.  DADD_DMUL is the no-FMA rate the parity kernels
 * are bound by (one rounding per operation, as the reference computes). */
#define MPB200_PEAK_DADD_DMUL 0
#define MPB200_PEAK_DFMA 1
#define MPB200_PEAK_FFMA 2
int mpb200_pipe_peak(int kind, double *ops_per_s);
/* Duration (ms, CUDA events, best of 3 after an L2 eviction) of a kernel that does nothing but WRITE a table of
 * t's shape: every column as one contiguous Int64 + one Float64 burst at colptr[w], columns visited in the order the
 * fill kernel visits them, constant data.  This is what the reference's CSC output format costs on the device before
 * any neighbour work: the floor the fill kernel's time is compared with, next to the plain-copy HBM peak. */
int mpb200_table_write_floor(const mpb200_table *t, double *ms);

#ifdef __cplusplus
}
#endif
#endif
