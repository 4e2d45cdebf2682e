// scan.cuh -- three-launch exclusive prefix sum (block sums -> scan of sums -> apply).
// Used for cell_start (int32 -> int32) and colptr (int32 counts -> int64, 1-based).
#pragma once
#include "common.cuh"

namespace mpb {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *warp_sums /*>= 8 + 1*/, T *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        T s = (lane < kScanThreads / 32) ? warp_sums[lane] : T(0);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    T warp_off = wid ? warp_sums[wid - 1] : T(0);
    if (total) *total = warp_sums[kScanThreads / 32 - 1];
    T r = warp_off + x - v;
    __syncthreads();
    return r;
}

template <class Tin, class Tout>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const Tin *__restrict__ in, int64_t n,
                                                               Tout *__restrict__ tile_sums) {
    __shared__ Tout ws[kScanThreads / 32 + 1];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    Tout s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (i < n) s += (Tout)in[i];
    }
    Tout tot;
    block_exclusive_scan<Tout>(s, ws, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of the tile sums in place; total -> *total_out
template <class Tout>
__global__ void __launch_bounds__(kScanThreads) scan_sums_inplace(Tout *__restrict__ sums, int64_t nb,
                                                                  Tout *__restrict__ total_out) {
    __shared__ Tout ws[kScanThreads / 32 + 1];
    Tout carry = 0;
    for (int64_t base = 0; base < nb; base += kScanThreads) {
        int64_t i = base + threadIdx.x;
        Tout v = (i < nb) ? sums[i] : Tout(0);
        Tout tot;
        Tout ex = block_exclusive_scan<Tout>(v, ws, &tot);
        if (i < nb) sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// out[i] = base + exclusive_prefix(in)[i] for i < n; out[n] = base + total
template <class Tin, class Tout>
__global__ void __launch_bounds__(kScanThreads) scan_apply(const Tin *__restrict__ in, int64_t n,
                                                           const Tout *__restrict__ tile_offs,
                                                           Tout *__restrict__ out, Tout base_value) {
    __shared__ Tout ws[kScanThreads / 32 + 1];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    Tout v[kScanItems];
    Tout s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k;
        v[k] = (i < n) ? (Tout)in[i] : Tout(0);
        s += v[k];
    }
    Tout tot;
    Tout ex = block_exclusive_scan<Tout>(s, ws, &tot);
    Tout run = base_value + tile_offs[blockIdx.x] + ex;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
        if (i == n - 1) out[n] = run;
    }
}

// host driver: 3 launches on ctx().stream.  tmp must hold ceil(n/kScanTile) Tout.
template <class Tin, class Tout>
int exclusive_scan(const Tin *in, int64_t n, Tout *out, Tout base_value, DevBuf &tmp, Tout *d_total) {
    cudaStream_t st = ctx().stream;
    if (n <= 0) {
        // out[0] = base, total = 0
        Tout b = base_value, z = 0;
        MPB_CUDA(cudaMemcpyAsync(out, &b, sizeof(Tout), cudaMemcpyHostToDevice, st));
        if (d_total) MPB_CUDA(cudaMemcpyAsync(d_total, &z, sizeof(Tout), cudaMemcpyHostToDevice, st));
        MPB_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    int64_t nb = ceil_div(n, kScanTile);
    if (int rc = tmp.reserve(sizeof(Tout) * (size_t)nb)) return rc;
    Tout *sums = tmp.as<Tout>();
    scan_tile_sums<Tin, Tout><<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, sums);
    MPB_LAUNCHED();
    scan_sums_inplace<Tout><<<1, kScanThreads, 0, st>>>(sums, nb, d_total);
    MPB_LAUNCHED();
    scan_apply<Tin, Tout><<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, sums, out, base_value);
    MPB_LAUNCHED();
    return 0;
}

}  // namespace mpb
