// peaks.cu -- measured arithmetic-pipe peaks of the device this process drives.
//
// MEASURED_PEAKS.json (driver-written) holds an HBM copy bandwidth and a bf16 tensor peak only; the
// FP64 kernels of this library (K5 ControlNN tables, K9 LQ edges, K10 Monte-Carlo rollouts) are bound
// by the FP64 CUDA-core pipe, and the parity rule (one rounding per operation) forbids FMA contraction
// in them, so their roofline denominator is the NO-FMA rate: separate DADD and DMUL instructions.
// mpb200_pipe_peak runs a register-resident dependency-free instruction stream per kind and returns
// operations per second (an FMA counts as 2), timed with CUDA events on the launching stream:
//   MPB200_PEAK_DADD_DMUL  alternating __dadd_rn / __dmul_rn   (what the parity kernels can reach)
//   MPB200_PEAK_DFMA       __fma_rn double                     (the advertised FP64 figure / 2 per FMA)
//   MPB200_PEAK_FFMA       __fmaf_rn                           (FP32 CUDA-core peak, for the FP32 prefilter)
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace mpb {

constexpr int kPeakChains = 8;      // independent accumulators per thread (covers the pipe latency)
constexpr int kPeakInner = 64;      // unrolled operations per chain per outer iteration

template <int KIND>
__global__ void __launch_bounds__(256) pipe_peak_kernel(int iters, double seed, double *__restrict__ sink) {
    double a[kPeakChains];
    float f[kPeakChains];
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) {
        a[c] = seed + 1e-3 * (threadIdx.x + c);
        f[c] = (float)a[c];
    }
    const double m = 1.0000001, s = 1e-9;
    const float mf = 1.0000001f, sf = 1e-9f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kPeakInner; ++k) {
#pragma unroll
            for (int c = 0; c < kPeakChains; ++c) {
                if (KIND == MPB200_PEAK_DADD_DMUL) a[c] = (k & 1) ? __dmul_rn(a[c], m) : __dadd_rn(a[c], s);
                if (KIND == MPB200_PEAK_DFMA) a[c] = __fma_rn(a[c], m, s);
                if (KIND == MPB200_PEAK_FFMA) f[c] = __fmaf_rn(f[c], mf, sf);
            }
        }
    }
    double t = 0;
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) t += a[c] + (double)f[c];
    if (t == 123.456) sink[0] = t;  // never true: keeps the chains alive
}

// ---- the write pattern of a CSC table, alone -----------------------------------------------------------------
// rball_fill has to emit every column as one contiguous Int64 burst + one Float64 burst (~220 B each at the FMT*
// radius) at colptr[w], and it visits the columns in GRID-CELL order while colptr runs in sample-index order: for
// samples numbered as drawn the bursts land at random offsets of a 440 MB range.  This kernel does ONLY that -- one
// warp per column, same visiting order, same streaming stores, constant data, no candidate loads, no distance, no
// sort -- so its duration is what the output format costs on this device before any neighbour work is done: the
// floor rball_fill can be compared with, next to the plain-copy HBM peak.
template <bool I32>
__global__ void __launch_bounds__(256)
table_write_pattern_kernel(const int64_t *__restrict__ colptr, const int *__restrict__ order, int64_t ncols,
                           long long *__restrict__ rowval, double *__restrict__ nzval) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // a warp takes 32 columns at a time: every lane fetches one column's extent (32 independent gathers in flight,
    // as in rball_fill, where thread t owns the record of query t), then the warp writes the columns one after the other
    for (int64_t t0 = gw * 32; t0 < ncols; t0 += nw * 32) {
        long long my_base = 0;
        int my_k = 0;
        if (t0 + lane < ncols) {
            const int64_t w = order ? order[t0 + lane] : t0 + lane;
            my_base = colptr[w] - 1;
            my_k = (int)(colptr[w + 1] - colptr[w]);
        }
        for (int c = 0; c < 32; ++c) {
            const long long base = __shfl_sync(0xffffffffu, my_base, c);
            const int k = __shfl_sync(0xffffffffu, my_k, c);
            for (int e = lane; e < k; e += 32) {
                if (I32) __stcs(reinterpret_cast<int *>(rowval) + base + e, e + 1);  // experiment: 4-byte row indices
                else __stcs(rowval + base + e, (long long)(e + 1));
                __stcs(nzval + base + e, 1.0);
            }
        }
    }
}

int table_write_floor_device(const mpb200_table *t, double *ms) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    static DevBuf rv, nz;
    if (int rc = rv.reserve(sizeof(int64_t) * (size_t)(t->nnz + 1))) return rc;
    if (int rc = nz.reserve(sizeof(double) * (size_t)(t->nnz + 1))) return rc;
    cudaEvent_t e0, e1;
    MPB_CUDA(cudaEventCreate(&e0));
    MPB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(t->ncols > 0 ? t->ncols : 1, 8 * 32), (int64_t)c.sm_count * 16);
    const int *order = t->has_order ? t->col_order.as<int>() : nullptr;
    // experiment (MPB200_FLOOR_BANDS=B): the same columns visited band by band (B index ranges), cell order inside a band
    static DevBuf banded;
    const char *eb = getenv("MPB200_FLOOR_BANDS");
    const int bands = eb ? atoi(eb) : 0;
    if (bands > 1 && order) {
        std::vector<int> h((size_t)t->ncols), o;
        MPB_CUDA(cudaMemcpyAsync(h.data(), order, sizeof(int) * (size_t)t->ncols, cudaMemcpyDeviceToHost, st));
        MPB_CUDA(cudaStreamSynchronize(st));
        const int64_t per = ceil_div(t->ncols, bands);
        o.reserve(h.size());
        for (int b = 0; b < bands; ++b)
            for (int w : h)
                if (w / per == b) o.push_back(w);
        if (int rc = banded.reserve(sizeof(int) * o.size())) return rc;
        MPB_CUDA(cudaMemcpyAsync(banded.p, o.data(), sizeof(int) * o.size(), cudaMemcpyHostToDevice, st));
        MPB_CUDA(cudaStreamSynchronize(st));
        order = banded.as<int>();
    }
    for (int rep = 0; rep < 4; ++rep) {
        MPB_CUDA(cudaMemsetAsync(rv.p, 0, 512 << 20 < rv.cap ? (size_t)(512 << 20) : rv.cap, st));  // evict the table from L2
        MPB_CUDA(cudaEventRecord(e0, st));
        if (getenv("MPB200_FLOOR_I32"))
            table_write_pattern_kernel<true><<<grid, 256, 0, st>>>(t->colptr.as<int64_t>(), order, t->ncols, rv.as<long long>(), nz.as<double>());
        else
            table_write_pattern_kernel<false><<<grid, 256, 0, st>>>(t->colptr.as<int64_t>(), order, t->ncols, rv.as<long long>(), nz.as<double>());
        MPB_LAUNCHED();
        MPB_CUDA(cudaEventRecord(e1, st));
        MPB_CUDA(cudaEventSynchronize(e1));
        float m = 0;
        MPB_CUDA(cudaEventElapsedTime(&m, e0, e1));
        if (rep > 0 && m < best) best = m;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    rv.release();
    nz.release();
    *ms = best;
    return 0;
}

int pipe_peak_device(int kind, double *ops_per_s) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    static DevBuf sink;
    if (int rc = sink.reserve(64)) return rc;
    const int blocks = c.sm_count * 8, iters = 400;
    cudaEvent_t e0, e1;
    MPB_CUDA(cudaEventCreate(&e0));
    MPB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up; best of the rest
        MPB_CUDA(cudaEventRecord(e0, st));
        if (kind == MPB200_PEAK_DADD_DMUL) pipe_peak_kernel<MPB200_PEAK_DADD_DMUL><<<blocks, 256, 0, st>>>(iters, 1.0, sink.as<double>());
        else if (kind == MPB200_PEAK_DFMA) pipe_peak_kernel<MPB200_PEAK_DFMA><<<blocks, 256, 0, st>>>(iters, 1.0, sink.as<double>());
        else pipe_peak_kernel<MPB200_PEAK_FFMA><<<blocks, 256, 0, st>>>(iters, 1.0, sink.as<double>());
        MPB_LAUNCHED();
        MPB_CUDA(cudaEventRecord(e1, st));
        MPB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double per_op = (kind == MPB200_PEAK_DADD_DMUL) ? 1.0 : 2.0;
    const double ops = (double)blocks * 256.0 * iters * kPeakInner * kPeakChains * per_op;
    *ops_per_s = ops / (best * 1e-3);
    return 0;
}

}  // namespace mpb
