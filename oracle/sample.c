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
 */
#include "mp_oracle.h"

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

/* returns the number of free samples written (N unless max_candidates ran out) */
int64_t orc_sample_free(const orc_checker *CC, const orc_space *S, int64_t N, uint64_t seed, int64_t max_candidates,
                        double *V_aos, int64_t *candidates)
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
    return got;
}
