// order.cu -- spatial (Morton) ordering of a device-resident sample set.
//
// The r-ball kernels visit queries in grid-cell order but write columns in INDEX order, so with samples in
// the order they were drawn every ~220-byte column burst lands at an unrelated address and the partially
// written sectors at column boundaries are completed by another warp much later.  With the samples stored in
// a spatially coherent order the same work runs 12% faster at C2 (scripts/order_experiment.py: 0.553 ->
// 0.487 ms per step).  FMT* does not care in which order i.i.d. samples are numbered (sampling.jl:23-37 appends
// them as they are drawn), so mpb200_sample_free can hand them out in Morton order.
//
// Key of a state (bit-identical to oracle/sample.c: orc_morton_key): the first min(n, 3) coordinates, each
// q_i = min(2^20 - 1, trunc((x_i - lo_i) / (hi_i - lo_i) * 2^20)), bit-interleaved with coordinate 0 in the
// least significant position.  Sorting is a stable LSD radix sort (CUB), so equal keys keep candidate order.
#include "common.cuh"
#include "predicates.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace mpb {

__device__ __forceinline__ unsigned long long spread3(unsigned long long x) {  // 21 bits -> every third bit
    x &= 0x1fffffULL;
    x = (x | (x << 32)) & 0x1f00000000ffffULL;
    x = (x | (x << 16)) & 0x1f0000ff0000ffULL;
    x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x << 2)) & 0x1249249249249249ULL;
    return x;
}

__global__ void __launch_bounds__(256) morton_keys(const double *__restrict__ V, int64_t N, int n, SpaceDev S,
                                                   unsigned long long *__restrict__ keys, int *__restrict__ idx) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const int m = n < 3 ? n : 3;
    unsigned long long key = 0;
    for (int i = 0; i < m; ++i) {
        const double g = dmul(ddiv(dsub(V[j * n + i], S.lo[i]), dsub(S.hi[i], S.lo[i])), 1048576.0);
        unsigned long long q = g > 0.0 ? (unsigned long long)g : 0ULL;  // NaN and negatives -> 0
        if (q > 1048575ULL) q = 1048575ULL;
        key |= spread3(q) << i;
    }
    keys[j] = key;
    idx[j] = (int)j;
}
__global__ void __launch_bounds__(256) gather_states(const double *__restrict__ V, const int *__restrict__ perm,
                                                     int64_t N, int n, double *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * n) return;
    const int64_t j = t / n;
    out[t] = V[(int64_t)perm[j] * n + (t - j * n)];
}

int make_space(const mpb200_space_desc *ss, int d_state, SpaceDev *out, int *dw);

// Reorders dV (N x n, AoS) in place by Morton key; `scratch` is grown as needed.
int morton_reorder_device(double *dV, int64_t N, const mpb200_space_desc *ss, DevBuf &scratch) {
    if (N <= 1) return 0;
    SpaceDev S;
    int dw = 0;
    if (int rc = make_space(ss, ss->n, &S, &dw)) return rc;
    const int n = ss->n;
    cudaStream_t st = ctx().stream;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, (int)N, 0, 60, st);
    const size_t kb = sizeof(unsigned long long) * (size_t)N, ib = ((sizeof(int) * (size_t)N + 15) / 16) * 16;
    const size_t vb = sizeof(double) * (size_t)(N * n);
    tmp_bytes = ((tmp_bytes + 255) / 256) * 256;
    if (int rc = scratch.reserve(2 * kb + 2 * ib + vb + tmp_bytes + 256)) return rc;
    char *base = scratch.as<char>();
    unsigned long long *k_in = reinterpret_cast<unsigned long long *>(base), *k_out = k_in + N;
    int *i_in = reinterpret_cast<int *>(base + 2 * kb), *i_out = reinterpret_cast<int *>(base + 2 * kb + ib);
    double *v_tmp = reinterpret_cast<double *>(base + 2 * kb + 2 * ib);
    void *tmp = base + 2 * kb + 2 * ib + vb;
    const unsigned nb = (unsigned)ceil_div(N, 256);
    morton_keys<<<nb, 256, 0, st>>>(dV, N, n, S, k_in, i_in);
    MPB_LAUNCHED();
    MPB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, i_in, i_out, (int)N, 0, 60, st));
    ctx().launches += 8;  // CUB's passes (histogram + one scatter per digit), not individually counted
    gather_states<<<(unsigned)ceil_div(N * n, 256), 256, 0, st>>>(dV, i_out, N, n, v_tmp);
    MPB_LAUNCHED();
    MPB_CUDA(cudaMemcpyAsync(dV, v_tmp, vb, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // namespace mpb
