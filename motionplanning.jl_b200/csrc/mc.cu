// mc.cu -- K10: importance-sampled Monte-Carlo trajectory collision probability.
//
// The estimator does not exist in the reference (only the paper links, README.md:9-10, and the
// helper geometry closest/closeR, SAT2D.jl:208-285, boxesND.jl:61-86); it is specified in
// SURVEY.md section 11 and restated in oracle/mc.c, whose arithmetic this kernel reproduces bit
// for bit: Philox4x32-10 keyed by (seed, rollout id), 53-bit uniforms, Box-Muller through the
// polynomial log / sincos below, closed-loop linear rollout z' = F_t z + G_t eps, per-step
// collision test with the reference's own predicates (predicates.cuh), mixture importance
// weight through the polynomial exp.  One thread per rollout (grid-stride over rollout ids);
// the per-component inner products live in shared memory; block partial sums are reduced in a
// fixed order, so a given launch geometry always returns the same bits.
#include "common.cuh"
#include "predicates.cuh"
#include "philox.cuh"

namespace mpb {

// ln(x): x = m 2^e, m in [sqrt(1/2), sqrt(2)); ln m = 2 atanh((m-1)/(m+1))   (oracle/mc.c: orc_det_log)
__device__ __forceinline__ double det_log(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m = __longlong_as_double((long long)b);
    if (m > 1.4142135623730951) { m = dmul(m, 0.5); e += 1; }
    const double z = ddiv(dsub(m, 1.0), dadd(m, 1.0)), z2 = dmul(z, z);
    double s = 1.0 / 23.0;
#pragma unroll
    for (int k = 21; k >= 1; k -= 2) s = dadd(dmul(s, z2), 1.0 / (double)k);
    return dadd(dmul((double)e, 0.6931471805599453), dmul(2.0, dmul(z, s)));
}
__device__ __forceinline__ void det_sincos2pi(double u, double *sn, double *cs) {
    const double v = dmul(u, 8.0);
    const int oct = (int)v;
    double f = dsub(v, (double)oct);
    if (oct & 1) f = dsub(1.0, f);
    const double th = dmul(f, 0.7853981633974483), t2 = dmul(th, th);
    double ps = -1.0 / 355687428096000.0;
    ps = dadd(dmul(ps, t2), 1.0 / 1307674368000.0);
    ps = dsub(dmul(ps, t2), 1.0 / 6227020800.0);
    ps = dadd(dmul(ps, t2), 1.0 / 39916800.0);
    ps = dsub(dmul(ps, t2), 1.0 / 362880.0);
    ps = dadd(dmul(ps, t2), 1.0 / 5040.0);
    ps = dsub(dmul(ps, t2), 1.0 / 120.0);
    ps = dadd(dmul(ps, t2), 1.0 / 6.0);
    const double s = dsub(th, dmul(dmul(th, t2), ps));
    double pc = 1.0 / 6402373705728000.0;
    pc = dsub(dmul(pc, t2), 1.0 / 20922789888000.0);
    pc = dadd(dmul(pc, t2), 1.0 / 87178291200.0);
    pc = dsub(dmul(pc, t2), 1.0 / 479001600.0);
    pc = dadd(dmul(pc, t2), 1.0 / 3628800.0);
    pc = dsub(dmul(pc, t2), 1.0 / 40320.0);
    pc = dadd(dmul(pc, t2), 1.0 / 720.0);
    pc = dsub(dmul(pc, t2), 1.0 / 24.0);
    pc = dadd(dmul(pc, t2), 0.5);
    const double c = dsub(1.0, dmul(t2, pc));
    const int o4 = oct & 3;
    const double a = (o4 == 0 || o4 == 3) ? s : c;
    const double b2 = (o4 == 0 || o4 == 3) ? c : s;
    *sn = (oct >= 4) ? -a : a;
    *cs = ((oct + 2) & 4) ? -b2 : b2;
}
__device__ __forceinline__ double det_exp(double x) {
    if (x > 690.0) x = 690.0;
    if (x < -690.0) x = -690.0;
    const double kf = floor(dadd(dmul(x, 1.4426950408889634), 0.5));
    const double r = dsub(dsub(x, dmul(kf, 0.6931471803691238)), dmul(kf, 1.9082149292705877e-10));
    double p = 1.0 / 87178291200.0;
    p = dadd(dmul(p, r), 1.0 / 6227020800.0);
    p = dadd(dmul(p, r), 1.0 / 479001600.0);
    p = dadd(dmul(p, r), 1.0 / 39916800.0);
    p = dadd(dmul(p, r), 1.0 / 3628800.0);
    p = dadd(dmul(p, r), 1.0 / 362880.0);
    p = dadd(dmul(p, r), 1.0 / 40320.0);
    p = dadd(dmul(p, r), 1.0 / 5040.0);
    p = dadd(dmul(p, r), 1.0 / 720.0);
    p = dadd(dmul(p, r), 1.0 / 120.0);
    p = dadd(dmul(p, r), 1.0 / 24.0);
    p = dadd(dmul(p, r), 1.0 / 6.0);
    p = dadd(dmul(p, r), 0.5);
    p = dadd(dmul(p, r), 1.0);
    p = dadd(dmul(p, r), 1.0);
    const int k = (int)kf;
    const double sc = __longlong_as_double((long long)(k + 1023) << 52);
    return dmul(p, sc);
}

struct McDev {
    int T, K, swept;
    const double *F, *G, *Wz, *wbar, *alpha, *mu, *hn2;
};

// hn2[k] = 0.5 |mu_k|^2 in stacked order (one thread per component; oracle: orc_mc_half_norms)
__global__ void mc_half_norms(const double *__restrict__ mu, int K, int TQ, double *__restrict__ hn2) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double n2 = 0.0;
    for (int g = 0; g < TQ; ++g) n2 = dadd(n2, dmul(mu[(size_t)k * TQ + g], mu[(size_t)k * TQ + g]));
    hn2[k] = dmul(0.5, n2);
}

constexpr int kMcThreads = 128;

template <int NZ, int Q, int DW, int KIND>
__global__ void __launch_bounds__(kMcThreads)
mc_rollout_kernel(McDev P, const double *__restrict__ g_table, int table_words, int M, bool use_smem,
                  unsigned long long seed, long long first, long long n, double *__restrict__ partials,
                  uint8_t *__restrict__ hit_out, double *__restrict__ w_out) {
    extern __shared__ double smem[];
    double *dots = smem;                                   // K x blockDim
    double *s_table = smem + (size_t)P.K * kMcThreads;     // obstacle table
    const double *T = g_table;
    if (use_smem) {
        for (int i = threadIdx.x; i < table_words; i += blockDim.x) s_table[i] = g_table[i];
        T = s_table;
    }
    __syncthreads();
    const Philox rng = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const int K = P.K, TT = P.T;
    double s1 = 0.0, s2 = 0.0, s0 = 0.0, nh = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long id = (unsigned long long)(first + i);
        const uint32_t id_lo = (uint32_t)id, id_hi = (uint32_t)(id >> 32);
        uint32_t rnd[4];
        rng(id_lo, id_hi, 0u, 0u, rnd);
        const double u = u53(rnd[0], rnd[1]);
        int comp = 0;
        double acc = P.alpha[0];
        while (comp < K && u >= acc) { ++comp; acc = dadd(acc, P.alpha[comp]); }
        const double *mu_c = comp > 0 ? P.mu + (size_t)(comp - 1) * TT * Q : nullptr;
        double z[NZ], wprev[DW], pair0 = 0.0, pair1 = 0.0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) z[a] = 0.0;
#pragma unroll
        for (int a = 0; a < DW; ++a) wprev[a] = P.wbar[a];
        for (int k = 0; k < K; ++k) dots[k * kMcThreads + threadIdx.x] = 0.0;
        bool hit = false;
        for (int t = 0; t < TT; ++t) {
            double eps[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                const int g = t * Q + j;
                if ((g & 1) == 0) {
                    rng(id_lo, id_hi, (uint32_t)(g >> 1), 1u, rnd);
                    const double u1 = u53(rnd[0], rnd[1]), u2 = u53(rnd[2], rnd[3]);
                    const double rr = sqrt(dmul(-2.0, det_log(u1)));
                    double sn, cs;
                    det_sincos2pi(u2, &sn, &cs);
                    pair0 = dmul(rr, cs); pair1 = dmul(rr, sn);
                }
                const double xi = (g & 1) ? pair1 : pair0;
                eps[j] = mu_c ? dadd(xi, mu_c[g]) : xi;
            }
            for (int k = 0; k < K; ++k) {
                const double *m = P.mu + ((size_t)k * TT + t) * Q;
                double d = dots[k * kMcThreads + threadIdx.x];
#pragma unroll
                for (int j = 0; j < Q; ++j) d = dadd(d, dmul(m[j], eps[j]));
                dots[k * kMcThreads + threadIdx.x] = d;
            }
            if (!hit) {
                const double *F = P.F + (size_t)t * NZ * NZ, *G = P.G + (size_t)t * NZ * Q;
                double zn[NZ], wcur[DW];
#pragma unroll
                for (int a = 0; a < NZ; ++a) {
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b < NZ; ++b) s = dadd(s, dmul(F[a * NZ + b], z[b]));
#pragma unroll
                    for (int j = 0; j < Q; ++j) s = dadd(s, dmul(G[a * Q + j], eps[j]));
                    zn[a] = s;
                }
#pragma unroll
                for (int a = 0; a < NZ; ++a) z[a] = zn[a];
#pragma unroll
                for (int a = 0; a < DW; ++a) {
                    double s = P.wbar[(size_t)(t + 1) * DW + a];
#pragma unroll
                    for (int b = 0; b < NZ; ++b) s = dadd(s, dmul(P.Wz[a * NZ + b], z[b]));
                    wcur[a] = s;
                }
                if (KIND == 0) {
                    hit = P.swept ? line_colliding_2d(T, wprev[0], wprev[DW > 1 ? 1 : 0], wcur[0], wcur[DW > 1 ? 1 : 0])
                                  : point_colliding_2d(T, wcur[0], wcur[DW > 1 ? 1 : 0]);
                } else {
                    hit = P.swept ? !box_segment_free<DW>(T, M, wprev, wcur) : !box_point_free<DW>(T, M, wcur);
                }
#pragma unroll
                for (int a = 0; a < DW; ++a) wprev[a] = wcur[a];
            }
        }
        double den = P.alpha[0];
        for (int k = 0; k < K; ++k)
            den = dadd(den, dmul(P.alpha[k + 1], det_exp(dsub(dots[k * kMcThreads + threadIdx.x], P.hn2[k]))));
        const double w = ddiv(1.0, den);
        if (hit_out) hit_out[i] = hit ? 1 : 0;
        if (w_out) w_out[i] = w;
        s0 += w;
        if (hit) { s1 += w; s2 += w * w; nh += 1.0; }
    }
    // fixed-order block reduction: xor-shuffle tree inside each warp, then warps in index order
    __shared__ double red[4][kMcThreads / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        nh += __shfl_xor_sync(0xffffffffu, nh, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2;
        red[2][threadIdx.x >> 5] = s0; red[3][threadIdx.x >> 5] = nh;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < kMcThreads / 32; ++w) t += red[threadIdx.x][w];
        partials[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

// single thread per quantity sums the block partials in block order
__global__ void mc_final_reduce(const double *__restrict__ partials, int nblocks, double *__restrict__ out4) {
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int b = 0; b < nblocks; ++b) t += partials[(size_t)b * 4 + threadIdx.x];
        out4[threadIdx.x] = t;
    }
}

// ---- host side -----------------------------------------------------------------------------------
#define MPB_MC_DISPATCH(NZ_, Q_, DW_, KIND_, CALL)                                          \
    do {                                                                                    \
        bool done__ = false;                                                                \
        if (KIND_ == 0 && DW_ == 2) {                                                       \
            if (NZ_ == 2 && Q_ == 2) { CALL(2, 2, 2, 0); done__ = true; }                   \
            else if (NZ_ == 4 && Q_ == 2) { CALL(4, 2, 2, 0); done__ = true; }              \
            else if (NZ_ == 4 && Q_ == 4) { CALL(4, 4, 2, 0); done__ = true; }              \
            else if (NZ_ == 8 && Q_ == 6) { CALL(8, 6, 2, 0); done__ = true; }              \
        } else if (KIND_ == 1 && DW_ == 2) {                                                \
            if (NZ_ == 2 && Q_ == 2) { CALL(2, 2, 2, 1); done__ = true; }                   \
            else if (NZ_ == 4 && Q_ == 2) { CALL(4, 2, 2, 1); done__ = true; }              \
            else if (NZ_ == 4 && Q_ == 4) { CALL(4, 4, 2, 1); done__ = true; }              \
            else if (NZ_ == 8 && Q_ == 6) { CALL(8, 6, 2, 1); done__ = true; }              \
        } else if (KIND_ == 1 && DW_ == 3) {                                                \
            if (NZ_ == 3 && Q_ == 3) { CALL(3, 3, 3, 1); done__ = true; }                   \
            else if (NZ_ == 6 && Q_ == 3) { CALL(6, 3, 3, 1); done__ = true; }              \
            else if (NZ_ == 6 && Q_ == 6) { CALL(6, 6, 3, 1); done__ = true; }              \
            else if (NZ_ == 12 && Q_ == 9) { CALL(12, 9, 3, 1); done__ = true; }            \
        }                                                                                   \
        if (!done__)                                                                        \
            return fail(MPB200_EARG, "unsupported MC shape (nz=%d, q=%d, dw=%d, checker=%d)", NZ_, Q_, DW_, KIND_); \
    } while (0)

int mc_run_device(const mpb200_mc_problem *p, const mpb200_obstacles *o, unsigned long long seed, long long first,
                  long long n, double *h_out4, uint8_t *h_hit, double *h_w) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const int T = p->T, nz = p->nz, q = p->q, dw = p->dw, K = p->K;
    if (o->kind == 1 && o->d != dw) return fail(MPB200_EARG, "box dimension %d != workspace dimension %d", o->d, dw);
    if (o->kind == 0 && dw != 2) return fail(MPB200_EARG, "2-D obstacles need a 2-D workspace");
    // pack the problem into one device buffer
    static DevBuf prob, part, outb, hitb, wb;
    const size_t nF = (size_t)T * nz * nz, nG = (size_t)T * nz * q, nW = (size_t)dw * nz, nB = (size_t)(T + 1) * dw,
                 nA = (size_t)K + 1, nM = (size_t)K * T * q, nH = (size_t)(K > 0 ? K : 1);
    const size_t total = nF + nG + nW + nB + nA + nM + nH;
    if (int rc = prob.reserve(sizeof(double) * total)) return rc;
    double *d = prob.as<double>();
    McDev P;
    P.T = T; P.K = K; P.swept = p->swept;
    size_t off = 0;
    auto put = [&](const double *src, size_t cnt, const double **dst) -> cudaError_t {
        *dst = d + off;
        cudaError_t e = cnt ? cudaMemcpyAsync(d + off, src, sizeof(double) * cnt, cudaMemcpyHostToDevice, st) : cudaSuccess;
        off += cnt;
        return e;
    };
    MPB_CUDA(put(p->F, nF, &P.F));
    MPB_CUDA(put(p->G, nG, &P.G));
    MPB_CUDA(put(p->Wz, nW, &P.Wz));
    MPB_CUDA(put(p->wbar, nB, &P.wbar));
    MPB_CUDA(put(p->alpha, nA, &P.alpha));
    MPB_CUDA(put(p->mu, nM, &P.mu));
    P.hn2 = d + off;
    if (K > 0) {
        mc_half_norms<<<ceil_div(K, 64), 64, 0, st>>>(P.mu, K, T * q, d + off);
        MPB_LAUNCHED();
    }
    const size_t table_bytes = sizeof(double) * (size_t)o->table_words;
    const size_t dots_bytes = sizeof(double) * (size_t)K * kMcThreads;
    const bool use_smem = table_bytes + dots_bytes <= 160 * 1024;
    const size_t smem = dots_bytes + (use_smem ? table_bytes : 0);
    if (dots_bytes > 160 * 1024) return fail(MPB200_EARG, "too many mixture components (K = %d)", K);
    int64_t blocks = ceil_div(n > 0 ? n : 1, kMcThreads);
    const int64_t cap = (int64_t)c.sm_count * 8;
    const int grid = (int)(blocks < cap ? blocks : cap);
    if (int rc = part.reserve(sizeof(double) * 4 * (size_t)grid)) return rc;
    if (int rc = outb.reserve(sizeof(double) * 4)) return rc;
    uint8_t *d_hit = nullptr;
    double *d_w = nullptr;
    if (h_hit) { if (int rc = hitb.reserve((size_t)n + 8)) return rc; d_hit = hitb.as<uint8_t>(); }
    if (h_w) { if (int rc = wb.reserve(sizeof(double) * (size_t)(n + 1))) return rc; d_w = wb.as<double>(); }
    phase_bank(MPB200_OP_OTHER);
    phase_mark(0);
#define CALL(NZ_, Q_, DW_, K_)                                                                                     \
    do {                                                                                                           \
        if (smem > 48 * 1024)                                                                                      \
            MPB_CUDA(cudaFuncSetAttribute(mc_rollout_kernel<NZ_, Q_, DW_, K_>,                                     \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
        mc_rollout_kernel<NZ_, Q_, DW_, K_><<<grid, kMcThreads, smem, st>>>(P, o->table.as<double>(), o->table_words, \
                                                                            o->M, use_smem, seed, first, n,        \
                                                                            part.as<double>(), d_hit, d_w);        \
    } while (0)
    MPB_MC_DISPATCH(nz, q, dw, o->kind, CALL);
#undef CALL
    MPB_LAUNCHED();
    mc_final_reduce<<<1, 32, 0, st>>>(part.as<double>(), grid, outb.as<double>());
    MPB_LAUNCHED();
    phase_mark(1);
    MPB_CUDA(cudaMemcpyAsync(h_out4, outb.p, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
    if (h_hit && n) MPB_CUDA(cudaMemcpyAsync(h_hit, d_hit, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (h_w && n) MPB_CUDA(cudaMemcpyAsync(h_w, d_w, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    MPB_CUDA(cudaStreamSynchronize(st));
    phases_collect(1);
    return 0;
}

}  // namespace mpb
