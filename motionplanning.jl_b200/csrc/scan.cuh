// scan.cuh -- three-launch exclusive prefix sum (block sums -> scan of sums -> apply).
// Used for cell_start (int32 -> int32) and colptr (int32 counts -> int64, 1-based).
#pragma once
#include "common.cuh"
#include <cstdlib>

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

// ---- single-pass form: decoupled look-back (one launch instead of three) --------------------------------------
// state[0] = ticket counter, state[1 + tile] = flag << 62 | value  (flag 1: tile aggregate, 2: inclusive prefix).
// Tiles are taken in ticket order, so every tile a block waits for belongs to a block that is already running
// (forward progress without co-residency assumptions).  The wait is bounded: a broken launch traps instead of
// hanging the device.  Values must stay below 2^62 (counts of table entries: they do).
__device__ __forceinline__ unsigned long long scan_ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void scan_st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <class Tin, class Tout>
__global__ void __launch_bounds__(kScanThreads) scan_lookback(const Tin *__restrict__ in, int64_t n, Tout *__restrict__ out,
                                                              Tout base_value, unsigned long long *__restrict__ state,
                                                              Tout *__restrict__ d_total) {
    __shared__ Tout ws[kScanThreads / 32 + 1];
    __shared__ long long s_tile;
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(state, 1ULL);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t first = tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
    Tout v[kScanItems];
    Tout s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = first + k;
        v[k] = (i < n) ? (Tout)in[i] : Tout(0);
        s += v[k];
    }
    Tout tot;
    const Tout ex = block_exclusive_scan<Tout>(s, ws, &tot);
    unsigned long long *mine = state + 1 + tile;
    if (threadIdx.x < 32) {  // warp 0: publish, then look back 32 tiles at a time
        const int lane = threadIdx.x;
        unsigned long long prefix = 0;
        if (tile == 0) {
            if (lane == 0) scan_st_release(mine, (2ULL << 62) | (unsigned long long)tot);
        } else {
            if (lane == 0) scan_st_release(mine, (1ULL << 62) | (unsigned long long)tot);
            int64_t look = tile - 1;  // newest tile not yet accounted for
            while (true) {
                const int64_t t = look - lane;
                unsigned long long w = 2ULL << 62;  // tiles before the first: an empty prefix
                if (t >= 0) {
                    unsigned tries = 0;
                    while (((w = scan_ld_acquire(state + 1 + t)) >> 62) == 0)
                        if (++tries > (1u << 22)) __trap();  // seconds, not forever
                }
                const unsigned has_prefix = __ballot_sync(0xffffffffu, (w >> 62) == 2);
                const int stop = has_prefix ? (__ffs(has_prefix) - 1) : 31;  // nearest tile with a complete prefix
                unsigned long long part = (lane <= stop) ? (w & ~(3ULL << 62)) : 0ULL;
#pragma unroll
                for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (has_prefix) break;
                look -= 32;
            }
            if (lane == 0) scan_st_release(mine, (2ULL << 62) | (prefix + (unsigned long long)tot));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    Tout run = base_value + (Tout)s_prefix + ex;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = first + k;
        if (i < n) out[i] = run;
        run += v[k];
        if (i == n - 1) {
            out[n] = run;
            if (d_total) *d_total = run - base_value;
        }
    }
}

// host driver: one launch (decoupled look-back; MPB200_SCAN3=1 selects the three-launch form) on ctx().stream.
// tmp must hold ceil(n/kScanTile) + 2 Tout / 8-byte words.
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
    static const bool three_launch = getenv("MPB200_SCAN3") != nullptr;
    if (!three_launch) {
        if (int rc = tmp.reserve(sizeof(unsigned long long) * (size_t)(nb + 2))) return rc;
        unsigned long long *state = tmp.as<unsigned long long>();
        MPB_CUDA(cudaMemsetAsync(state, 0, sizeof(unsigned long long) * (size_t)(nb + 1), st));
        scan_lookback<Tin, Tout><<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, out, base_value, state, d_total);
        MPB_LAUNCHED();
        return 0;
    }
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
