/* oracle/sample.c -- TEST INFRASTRUCTURE ONLY (see mp_oracle.h).
 *
 * Batched free-state sampling, SURVEY section 8(f) rank 1: the reference's sample_free!(P, N)
 * (sampling.jl:23-37) draws states uniformly in the state-space bounds (sample_space, statespaces.jl:
 * rand over [lo, hi]) and keeps those for which is_free_state(v, CC, SS) holds, until N are collected.
 * The reference draws from Julia's global MersenneTwister, which cannot be reproduced here -- PARITY UNPINNED
 * against the reference for the sample VALUES; what is specified (and shared bit for bit with the CUDA path)
 * is the candidate stream below and the acceptance rule, which is the reference's own predicate.
 *
 * Candidate c = 0, 1, 2, ... ; coordinate i of candidate c:
 *     rnd = Philox4x32-10(counter = (c & 0xffffffff, c >> 32, i / 2, 0x53414D50), key = (seed lo, seed hi))
 *     u   = u53(rnd[0], rnd[1]) for even i, u53(rnd[2], rnd[3]) for odd i          (53-bit, in (0, 1))
 *     x_i = lo_i + u * (hi_i - lo_i)                                               (one rounding per operation)
 * Output: the first N free candidates in candidate order, and the number of candidates consumed.
 * order = 1 additionally sorts the N samples by Morton key (stable: equal keys keep candidate order); key of a
 * state = bit interleave (coordinate 0 least significant) of q_i = min(2^20 - 1, trunc((x_i - lo_i) / (hi_i -
 * lo_i) * 2^20)) over the first min(n, 3) coordinates.  A spatially coherent numbering makes the neighbour
 * tables cheaper to write; FMT* itself does not care how i.i.d. samples are numbered.
 */
#include "mp_oracle.h"
#include <stdlib.h>
#include <string.h>

static inline double u53s(uint32_t hi, uint32_t lo)
{
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

void orc_sample_candidate(const orc_space *S, uint64_t seed, int64_t c, double *x)
{
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int b = 0; 2 * b < S->n; ++b) {
        const uint32_t ctr[4] = {(uint32_t)c, (uint32_t)((uint64_t)c >> 32), (uint32_t)b, 0x53414D50u};
        uint32_t rnd[4];
        orc_philox4x32_10(ctr, key, rnd);
        for (int h = 0; h < 2 && 2 * b + h < S->n; ++h) {
            const int i = 2 * b + h;
            const double u = u53s(rnd[2 * h], rnd[2 * h + 1]);
            const double w = S->hi[i] - S->lo[i];
            const double t = u * w;
            x[i] = S->lo[i] + t;
        }
    }
}

static uint64_t spread3(uint64_t x)
{
    x &= 0x1fffffULL;
    x = (x | (x << 32)) & 0x1f00000000ffffULL;
    x = (x | (x << 16)) & 0x1f0000ff0000ffULL;
    x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x << 2)) & 0x1249249249249249ULL;
    return x;
}
uint64_t orc_morton_key(const orc_space *S, const double *x)
{
    const int m = S->n < 3 ? S->n : 3;
    uint64_t key = 0;
    for (int i = 0; i < m; ++i) {
        const double t = x[i] - S->lo[i];
        const double w = S->hi[i] - S->lo[i];
        const double f = t / w;
        const double g = f * 1048576.0;
        uint64_t q = g > 0.0 ? (uint64_t)g : 0;
        if (q > 1048575ULL) q = 1048575ULL;
        key |= spread3(q) << i;
    }
    return key;
}
/* stable bottom-up merge sort of the sample rows by key */
static void sort_by_key(double *V, uint64_t *key, int64_t N, int n)
{
    double *V2 = (double *)malloc(sizeof(double) * (size_t)(N * n));
    uint64_t *k2 = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)N);
    double *a = V, *b = V2;
    uint64_t *ka = key, *kb = k2;
    for (int64_t w = 1; w < N; w *= 2) {
        for (int64_t lo = 0; lo < N; lo += 2 * w) {
            int64_t mid = lo + w < N ? lo + w : N, hi = lo + 2 * w < N ? lo + 2 * w : N;
            int64_t i = lo, j = mid, o = lo;
            while (i < mid || j < hi) {
                int take_left = j >= hi || (i < mid && ka[i] <= ka[j]);   /* <= keeps equal keys in order */
                int64_t src = take_left ? i++ : j++;
                kb[o] = ka[src];
                memcpy(b + o * n, a + src * n, sizeof(double) * (size_t)n);
                ++o;
            }
        }
        double *t = a; a = b; b = t;
        uint64_t *kt = ka; ka = kb; kb = kt;
    }
    if (a != V) memcpy(V, a, sizeof(double) * (size_t)(N * n));
    free(V2);
    free(k2);
}

/* returns the number of free samples written (N unless max_candidates ran out) */
int64_t orc_sample_free(const orc_checker *CC, const orc_space *S, int64_t N, uint64_t seed, int64_t max_candidates,
                        int order, double *V_aos, int64_t *candidates)
{
    int64_t got = 0, c = 0;
    double x[16];
    for (; got < N && c < max_candidates; ++c) {
        orc_sample_candidate(S, seed, c, x);
        if (orc_is_free_state(CC, S, x)) {
            for (int i = 0; i < S->n; ++i) V_aos[got * S->n + i] = x[i];
            ++got;
        }
    }
    if (candidates) *candidates = c;
    if (order == 1 && got > 1) {
        uint64_t *key = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)got);
        for (int64_t i = 0; i < got; ++i) key[i] = orc_morton_key(S, V_aos + i * S->n);
        sort_by_key(V_aos, key, got, S->n);
        free(key);
    }
    return got;
}
