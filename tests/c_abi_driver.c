/* c_abi_driver.c -- a plain-C caller of libmpb200 (includes ONLY include/mpb200.h, links nothing but the
 * library): proves the drop-in boundary from a non-Python host, the way a Julia `ccall` would use it.
 *
 *   c_abi_driver <in.bin> <out.bin>
 *
 * in.bin  (written by tests/test_gpu_c_abi.py from the C1 workload):
 *   int64 N, d, n_gates, n_shapes, data_len; double r; double lo[d], hi[d];
 *   double V[d*N] (column-major d x N, as Vector{SVector{d,Float64}} lies in memory);
 *   int32 gate_parent[n_gates]; double gate_aabb[4*n_gates];
 *   int32 shape_kind[n_shapes], shape_gate[n_shapes], shape_off[n_shapes+1]; double data[data_len]
 * out.bin: int64 nnz, checks; int64 colptr[N+1]; int64 rowval[nnz]; double nzval[nnz];
 *          uint64 point_bits[ceil(N/64)]; uint64 edge_bits[ceil(nnz/64)]
 * Exit code 0 on success; on any library error the message of mpb200_last_error() goes to stderr. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpb200.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int rc__ = (call);                                                                \
        if (rc__ != MPB200_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, mpb200_last_error());          \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

static void *rd(FILE *f, size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p || (bytes && fread(p, 1, bytes, f) != bytes)) { fprintf(stderr, "short read\n"); exit(3); }
    return p;
}

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    int64_t *hdr = rd(f, 5 * sizeof(int64_t));
    const int64_t N = hdr[0], d = hdr[1], n_gates = hdr[2], n_shapes = hdr[3], data_len = hdr[4];
    double *r = rd(f, sizeof(double));
    double *lo = rd(f, (size_t)d * 8), *hi = rd(f, (size_t)d * 8);
    double *V = rd(f, (size_t)(d * N) * 8);
    int32_t *gate_parent = rd(f, (size_t)n_gates * 4);
    double *gate_aabb = rd(f, (size_t)n_gates * 32);
    int32_t *shape_kind = rd(f, (size_t)n_shapes * 4), *shape_gate = rd(f, (size_t)n_shapes * 4);
    int32_t *shape_off = rd(f, (size_t)(n_shapes + 1) * 4);
    double *data = rd(f, (size_t)data_len * 8);
    fclose(f);

    CHECK(mpb200_init(0));
    mpb200_samples *s = NULL;
    CHECK(mpb200_samples_create(V, N, (int)d, &s));
    mpb200_table *t = NULL;
    int64_t nnz = 0;
    CHECK(mpb200_inball_build(s, *r, &t, &nnz));           /* helper_data_structures + every inball (geometric.jl:14, nearneighbors.jl:179-183) */
    int64_t *colptr = malloc((size_t)(N + 1) * 8), *rowval = malloc((size_t)(nnz + 1) * 8);
    double *nzval = malloc((size_t)(nnz + 1) * 8);
    CHECK(mpb200_table_fetch(t, colptr, rowval, nzval));   /* the fields of ImmutableNNC.D (nearneighbors.jl:23-28) */

    mpb200_obstacles2d_desc od = {(int32_t)n_gates, gate_parent, gate_aabb, (int32_t)n_shapes, shape_kind, shape_gate,
                                  shape_off, data, 0};
    mpb200_obstacles *o = NULL;
    CHECK(mpb200_obstacles2d_create(&od, &o));             /* PointRobot2D(obstacles) (robots2D.jl:5-10) */
    mpb200_space_desc ss = {(int32_t)d, lo, hi, 0, (int32_t)d, NULL, NULL};
    const size_t pw = (size_t)((N + 63) / 64), ew = (size_t)((nnz + 63) / 64);
    uint64_t *pbits = calloc(pw + 1, 8), *ebits = calloc(ew + 1, 8);
    int64_t checks = 0;
    CHECK(mpb200_points_free(s, o, &ss, pbits));           /* F[i] = is_free_state(V[i], CC, SS) (fmt.jl:31-36) */
    CHECK(mpb200_edges_free(s, t, o, &ss, ebits, &checks)); /* is_free_motion(V[y], V[x], CC, SS) per stored edge (fmt.jl:75) */

    /* error convention: a bad argument must come back as a code + message, never a crash */
    mpb200_samples *bad = NULL;
    if (mpb200_samples_create(V, N, 99, &bad) != MPB200_EARG || strlen(mpb200_last_error()) == 0) {
        fprintf(stderr, "bad argument was not rejected\n");
        return 4;
    }

    FILE *g = fopen(argv[2], "wb");
    if (!g) { perror(argv[2]); return 1; }
    fwrite(&nnz, 8, 1, g); fwrite(&checks, 8, 1, g);
    fwrite(colptr, 8, (size_t)(N + 1), g); fwrite(rowval, 8, (size_t)nnz, g); fwrite(nzval, 8, (size_t)nnz, g);
    fwrite(pbits, 8, pw, g); fwrite(ebits, 8, ew, g);
    fclose(g);
    CHECK(mpb200_table_destroy(t));
    CHECK(mpb200_obstacles_destroy(o));
    CHECK(mpb200_samples_destroy(s));
    mpb200_shutdown();
    printf("c_abi_driver ok: N=%lld nnz=%lld checks=%lld launches=n/a\n", (long long)N, (long long)nnz, (long long)checks);
    return 0;
}
