// tc_rball.cu -- K3 on the 5th-generation tensor cores: the all-pairs prefilter of the d >= 4
// Euclidean r-ball as a TF32 tcgen05.mma with TMEM accumulators, K4 (exact FP64 recheck) unchanged.
//
// For centred points x~ = x - c the membership test  |x - y|^2 <= r^2  is
//      x~.y~ + (r^2 - |y~|^2)/2  >=  |x~|^2 / 2 .
// The left side is ONE dense contraction over K = 16: slots 0..d-1 hold the coordinates, two
// more slots hold a hi/lo TF32 split of (r^2 - |y~|^2)/2 on the sample side against 1.0 on the
// query side.  A CTA owns 128 query columns (operand A, 128 x 16, staged once) and sweeps the
// samples in tiles of 128 (operand B): two tcgen05.mma.kind::tf32 (M=128, N=128, K=8) per tile
// write a 128 x 128 FP32 accumulator into tensor memory; the four warps read their TMEM lanes
// back with tcgen05.ld (one query row per thread) and keep a running max per 32 columns, so the
// epilogue costs ~1 instruction per pair; only chunks whose max passes  |x~|^2/2 - delta  are
// rescanned, and only survivors reach the exact FP64 test that decides membership and the stored
// distance.  delta bounds the TF32 input rounding and FP32 accumulation error, so
// the prefilter has no false negatives.  Up to four CTAs share an SM (4 x 128 TMEM columns), which
// overlaps one CTA's MMA with the others' epilogues without an explicit pipeline.
//
// What bounded it (and what did not): earlier revisions took the accumulator read-back (tcgen05.ld at "64 B per clock
// per SM") for the bound because the measured times sat just under it.  A full-size ncu capture showed tensor memory
// busy 21% of the cycles and the issue slots / CTA barrier saturated by the RARE path instead: one lane's inline exact
// recheck kept its warp, and at the barrier its CTA, waiting.  Every mode therefore pushes the candidates that pass into
// a per-warp queue that is rechecked 32 at a time, one candidate per lane, and full-range builds (MODE 3) sweep each
// unordered pair once: C3 went from 0.91 s to 0.31 s, 95 B per clock per SM of accumulators read back.
#include "common.cuh"
#include "scan.cuh"
#include "tc_rball.cuh"

namespace mpb {

constexpr int kTcM = 128;      // query rows per CTA (= threads: one TMEM lane per thread)
#ifndef MPB_TC_N
#define MPB_TC_N 64
#endif
constexpr int kTcN = MPB_TC_N;  // samples per MMA tile (= TMEM columns per CTA): 64 -> eight CTAs share an SM's 512 columns
constexpr int kTcCtas = 512 / kTcN;
constexpr int kTcK = 16;       // contraction length: d coordinates + 2 threshold slots, zero padded

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Operand arrays in the canonical K-major / no-swizzle UMMA layout, per tile of 128 rows:
//   [chunk 0..3][row 0..127][4 x tf32]   (16-byte chunks; core matrix = 8 rows x 16 B contiguous)
// B version: slots d, d+1 = hi/lo of (r^2 - |y~|^2)/2.   nrm_half[j] = |x~_j|^2 / 2 (FP32).
template <int D>
__global__ void __launch_bounds__(256)
tc_prepare(const double *__restrict__ V, int64_t N, int64_t Npad, const double *__restrict__ center, double r2,
           float *__restrict__ opB, float *__restrict__ nrm_half) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Npad) return;
    float slot[kTcK];
#pragma unroll
    for (int k = 0; k < kTcK; ++k) slot[k] = 0.0f;
    float nh = __int_as_float(0x7f800000);  // padding rows: +inf norm -> never a candidate as a query
    if (j < N) {
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double xc = V[j * D + k] - center[k];
            n2 += xc * xc;
            slot[k] = to_tf32((float)xc);
        }
        const float v = (float)(0.5 * (r2 - n2));
        const float hi = to_tf32(v);
        slot[D] = hi;
        slot[D + 1] = to_tf32(v - hi);
        nh = (float)(0.5 * n2);
    } else {
        slot[D] = -3.0e38f;  // padding samples: the contraction is hugely negative -> never a candidate
    }
    nrm_half[j] = nh;
    const int64_t tile = j >> 7, row = j & 127;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        reinterpret_cast<float4 *>(opB)[(tile * 4 + c) * 128 + row] =
            make_float4(slot[4 * c], slot[4 * c + 1], slot[4 * c + 2], slot[4 * c + 3]);
}

// ---- raw tcgen05 / mbarrier PTX ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), leading (K-chunk) byte offset >> 4 in [16,30), stride (8-row group)
// byte offset >> 4 in [32,46), version 1 in [46,48), layout type 0 in [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ULL << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32
// (bits 7-9, 10-12 = 2), both K-major, N >> 3 in [17,23), M >> 4 in [24,29)
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded spin: a mis-programmed MMA must not hang the GPU -- trap after ~seconds instead
    for (unsigned tries = 0; tries < (1u << 28); ++tries) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int D>
__device__ __forceinline__ double tc_exact_sq(const double *__restrict__ a, const double *__restrict__ b) {
    double t = __dsub_rn(a[0], b[0]);
    double s = __dmul_rn(t, t);
#pragma unroll
    for (int i = 1; i < D; ++i) {
        t = __dsub_rn(a[i], b[i]);
        s = __dadd_rn(s, __dmul_rn(t, t));
    }
    return s;
}

// MODE 0: count only, MODE 2: count + append (index, exact squared distance) to the column's slab.
// MODE 3: the SYMMETRIC sweep (full-range builds): the relation and the stored distance are symmetric -- (a-b)^2 and
// (b-a)^2 are the same bits -- so a CTA only multiplies its 128 queries against sample tiles at or after its own and
// every accepted pair (q, j), j > q, is appended to BOTH columns, with an atomic slot counter per column.  Half the
// MMAs, half the accumulator read-back (the kernel's bound, see the header); the slabs are then unordered and
// slab_sort_to_csc sorts each column by index.  CTAs are launched longest-first (tile 0 sweeps everything).
template <int D, int MODE>
__global__ void __launch_bounds__(kTcM, kTcCtas)
tc_rball_kernel(const double *__restrict__ V, const float *__restrict__ opB, const float *__restrict__ nrm_half,
                int64_t N, int64_t Npad, int64_t q0, int64_t nq, double r2, float delta, int *__restrict__ counts,
                int cap, int *__restrict__ slab_j, double *__restrict__ slab_s, int *__restrict__ rcounts) {
    __shared__ __align__(128) float4 sA[4 * kTcM];   // 8 KB
    __shared__ __align__(128) float4 sB[4 * kTcN];   // 8 KB
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    // candidates that passed the tensor-core test wait in a per-warp queue until 32 of them can be checked
    // exactly by 32 lanes at once (one candidate per lane), instead of one lane's loop stalling its whole warp
    __shared__ uint32_t s_queue[kTcM / 32][64];
    __shared__ int s_own[kTcM];   // accepted entries per query row (MODE 2 / 3: appended to the front of its slab row)
    const int tid = threadIdx.x, warp = tid >> 5;
    const int lane = tid & 31;
    s_own[tid] = 0;
    int q_count = 0;                                  // queued candidates of this warp (warp-uniform)
    const int64_t w = (int64_t)blockIdx.x * kTcM + tid;
    const bool active = w < nq;
    const int64_t q = q0 + w;

    // operand A: this thread's query row, threshold slots replaced by 1.0 (exact in TF32)
    {
        const int64_t qq = active ? q : 0;
        const int64_t tile = qq >> 7, row = qq & 127;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float4 v = reinterpret_cast<const float4 *>(opB)[(tile * 4 + c) * 128 + row];
            float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = 4 * c + e;
                if (k == D || k == D + 1) f[e] = 1.0f;
                if (!active) f[e] = 0.0f;
            }
            sA[c * kTcM + tid] = make_float4(f[0], f[1], f[2], f[3]);
        }
    }
    // pass iff acc >= |x~|^2/2 - delta ; inactive rows never pass
    const float pass_at = active ? (nrm_half[q] - delta) : __int_as_float(0x7f800000);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(kTcN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        mbar_init(smem_u32(&s_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc_tf32(kTcM, kTcN);
    // chunk stride (LBO) = 128 rows x 16 B, 8-row group stride (SBO) = 128 B; the second MMA (K slots
    // 8..15) starts two chunks further
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const uint64_t a_desc0 = umma_desc(a0, kTcM * 16, 128), a_desc1 = umma_desc(a0 + 2 * kTcM * 16, kTcM * 16, 128);
    const uint64_t b_desc0 = umma_desc(b0, kTcN * 16, 128), b_desc1 = umma_desc(b0 + 2 * kTcN * 16, kTcN * 16, 128);
    const uint32_t bar = smem_u32(&s_bar);
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);

    int cnt = 0;
    uint32_t phase = 0;
    // MODE 3: exact FP64 test of the first n (<= 32) queued candidates, one per lane, and both appends.  A candidate is
    // (sample j << 5 | query row of this warp).  Own column: the hits of one query are ranked in queue order (which is
    // ascending j) with match_any / popc on top of the row's counter in shared memory, so the front of the slab row
    // stays ascending without atomics; partner column: back of its slab row, slot from a global atomic.
    auto drain = [&](int n) {
        const bool act = lane < n;
        const uint32_t c = act ? s_queue[warp][lane] : 0u;
        const int ql = act ? (int)(c & 31u) : 32 + lane;          // inactive lanes match nobody
        const int64_t j = (int64_t)(c >> 5);
        const int64_t wl = (int64_t)blockIdx.x * kTcM + warp * 32 + (ql & 31);   // row of this launch (slab row in MODE 2)
        const int64_t qg = q0 + wl;                                              // global sample index of the query
        double s64 = 0.0;
        bool hit = false;
        if (act) {
            s64 = tc_exact_sq<D>(V + qg * D, V + j * D);
            hit = s64 <= r2;
        }
        const unsigned hitmask = __ballot_sync(0xffffffffu, hit);
        const unsigned peers = __match_any_sync(0xffffffffu, ql);
        const unsigned lt = (1u << lane) - 1u;
        int base = 0;
        if (act) base = s_own[warp * 32 + ql];
        __syncwarp();
        if (hit && MODE != 0) {
            const int pos = base + __popc(peers & hitmask & lt);
            if (pos < cap) { slab_j[wl * cap + pos] = (int)j; slab_s[wl * cap + pos] = s64; }
            if (MODE == 3) {
                const int slot = atomicAdd(&rcounts[j], 1);
                if (slot < cap) { slab_j[j * cap + (cap - 1 - slot)] = (int)qg; slab_s[j * cap + (cap - 1 - slot)] = s64; }
            }
        }
        if (act && lane == __ffs(peers) - 1) s_own[warp * 32 + ql] = base + __popc(peers & hitmask);
        __syncwarp();
        // keep what is left of the queue at its front
        const int rem = q_count - n;
        const uint32_t keep = (lane < rem) ? s_queue[warp][n + lane] : 0u;
        __syncwarp();
        if (lane < rem) s_queue[warp][lane] = keep;
        __syncwarp();
        q_count = rem;
    };
    // operand image: [tile of 128 samples][chunk 0..3][row 0..127]; an MMA tile = rows (t0 & 127) .. + kTcN of it
    constexpr int kStage = 4 * kTcN / kTcM;  // float4 per thread per tile
    float4 nextB[kStage];
    auto load_tile = [&](int64_t t0) {
        const float4 *src = reinterpret_cast<const float4 *>(opB) + (t0 >> 7) * (4 * 128) + (t0 & 127);
#pragma unroll
        for (int u = 0; u < kStage; ++u) {
            const int i = tid + u * kTcM;
            nextB[u] = src[(i / kTcN) * 128 + (i % kTcN)];
        }
    };
    const int64_t t_first = (MODE == 3) ? (int64_t)blockIdx.x * kTcM : 0;
    if (t_first < Npad) load_tile(t_first);
    for (int64_t t0 = t_first; t0 < Npad; t0 += kTcN) {
        // stage operand B (loaded one tile ahead into registers, so its global-load latency is behind the epilogue)
        {
#pragma unroll
            for (int u = 0; u < kStage; ++u) sB[tid + u * kTcM] = nextB[u];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_tf32(tmem, a_desc0, b_desc0, idesc, 0u);
            umma_tf32(tmem, a_desc1, b_desc1, idesc, 1u);
            umma_commit(bar);
        }
        if (t0 + kTcN < Npad) load_tile(t0 + kTcN);
        mbar_wait(bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: row `tid` of the 128 x 128 accumulator, 32 columns at a time
#pragma unroll 1
        for (int c0 = 0; c0 < kTcN; c0 += 32) {
            uint32_t rr[32];
            tmem_ld32(my_tmem + (uint32_t)c0, rr);
            // running maxima of the columns i = k (mod 4), three operands per instruction (FMNMX3)
            float m0 = fmax3(__uint_as_float(rr[0]), __uint_as_float(rr[4]), __uint_as_float(rr[8]));
            float m1 = fmax3(__uint_as_float(rr[1]), __uint_as_float(rr[5]), __uint_as_float(rr[9]));
            float m2 = fmax3(__uint_as_float(rr[2]), __uint_as_float(rr[6]), __uint_as_float(rr[10]));
            float m3 = fmax3(__uint_as_float(rr[3]), __uint_as_float(rr[7]), __uint_as_float(rr[11]));
#pragma unroll
            for (int i = 12; i < 28; i += 8) {
                m0 = fmax3(m0, __uint_as_float(rr[i]), __uint_as_float(rr[i + 4]));
                m1 = fmax3(m1, __uint_as_float(rr[i + 1]), __uint_as_float(rr[i + 5]));
                m2 = fmax3(m2, __uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 6]));
                m3 = fmax3(m3, __uint_as_float(rr[i + 3]), __uint_as_float(rr[i + 7]));
            }
            m0 = fmaxf(m0, __uint_as_float(rr[28]));
            m1 = fmaxf(m1, __uint_as_float(rr[29]));
            m2 = fmaxf(m2, __uint_as_float(rr[30]));
            m3 = fmaxf(m3, __uint_as_float(rr[31]));
            const bool pass = fmaxf(fmax3(m0, m1, m2), m3) >= pass_at;
            {
                if (__any_sync(0xffffffffu, pass)) {   // warp-uniform
                    uint32_t mask = 0;
                    if (pass) {
                        // the four running maxima cover the columns i = k (mod 4): only the quarters whose maximum
                        // passed are compared value by value (a passing lane typically has one candidate)
                        if (m0 >= pass_at) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4) mask |= (__uint_as_float(rr[i]) >= pass_at) ? (1u << i) : 0u;
                        }
                        if (m1 >= pass_at) {
#pragma unroll
                            for (int i = 1; i < 32; i += 4) mask |= (__uint_as_float(rr[i]) >= pass_at) ? (1u << i) : 0u;
                        }
                        if (m2 >= pass_at) {
#pragma unroll
                            for (int i = 2; i < 32; i += 4) mask |= (__uint_as_float(rr[i]) >= pass_at) ? (1u << i) : 0u;
                        }
                        if (m3 >= pass_at) {
#pragma unroll
                            for (int i = 3; i < 32; i += 4) mask |= (__uint_as_float(rr[i]) >= pass_at) ? (1u << i) : 0u;
                        }
                        const int64_t j0 = t0 + c0;
                        const int64_t lo = q - j0;
                        if (MODE == 3) {   // only partners j with q < j < N (the pair's other half belongs to row j)
                            if (lo >= 31) mask = 0u;
                            else if (lo >= 0) mask &= ~((2u << (int)lo) - 1u);   // bits 0 .. lo are j <= q
                        } else if (lo >= 0 && lo < 32) {
                            mask &= ~(1u << (int)lo);                             // j == q
                        }
                        const int64_t hi = N - j0;              // bits >= hi are padding samples
                        if (hi <= 0) mask = 0u;
                        else if (hi < 32) mask &= (1u << (int)hi) - 1u;
                    }
                    for (;;) {
                        const bool has = mask != 0u;
                        const unsigned bal = __ballot_sync(0xffffffffu, has);
                        if (bal == 0u) break;
                        if (has) {
                            const int i = __ffs(mask) - 1;
                            mask &= mask - 1u;
                            s_queue[warp][q_count + __popc(bal & ((1u << lane) - 1u))] =
                                ((uint32_t)(t0 + c0 + i) << 5) | (uint32_t)lane;
                        }
                        q_count += __popc(bal);
                        __syncwarp();
                        if (q_count >= 32) drain(32);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // accumulator and sB are free for the next tile
    }
    if (q_count > 0) drain(q_count);
    __syncwarp();
    if (active) counts[w] = s_own[tid];  // MODE 3: own entries; the entries appended by partners are counted in rcounts
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTcN));
}

// ---- host side -------------------------------------------------------------------------------------
// delta: half of the worst-case error of the computed test quantity (see header + DESIGN.md 10)
static float tc_delta(double r, int D, double Rmax2) {
    const double u_tf32 = 4.8828125e-4;                       // 2^-11, cvt.rna
    const double e_dot = 2.0 * u_tf32 * (1.0 + u_tf32) * Rmax2;     // |x~.y~ - tf32(x~).tf32(y~)| <= 2u(1+u)|x~||y~|
    const double e_acc = 64.0 * 1.1920929e-7 * (Rmax2 + 0.5 * (r * r + Rmax2));  // FP32 accumulation of 16 products
    const double e_thr = 3.6e-7 * 0.5 * (r * r + Rmax2);     // hi/lo split + FP32 rounding of (r^2-|y~|^2)/2 and |x~|^2/2
    return (float)(1.5 * (e_dot + e_acc + 2.0 * e_thr) + 1e-30);
}


// centre of the samples' bounding box and the squared radius of the centred cloud
static double tc_center(const mpb200_samples *s, int D, double *center) {
    double Rmax2 = 0;
    for (int k = 0; k < D; ++k) {
        center[k] = 0.5 * (s->h_bbox[k] + s->h_bbox[D + k]);
        const double h = fmax(fabs(s->h_bbox[k] - center[k]), fabs(s->h_bbox[D + k] - center[k]));
        Rmax2 += h * h;
    }
    return Rmax2;
}

// The tensor-core test accepts  s <= r^2 + 2 delta.  When 2 delta is not small against r^2 (a wide, sparse
// cloud with a small radius) nearly every 32-column chunk becomes a candidate and the exact recheck does the
// all-pairs work: correct, but slower than the FP32 CUDA-core sweep, whose slack scales with |x| r instead of
// |x|^2.  The caller then takes that path.
bool tc_band_is_tight(const mpb200_samples *s, double r) {
    double center[16];
    const double Rmax2 = tc_center(s, s->d, center);
    return 2.0 * (double)tc_delta(r, s->d, Rmax2) <= 0.25 * r * r;
}

template <int D>
int tc_prepare_operands(mpb200_samples *s, double r, TcPlan *plan) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const int64_t N = s->N, Npad = (N + 127) & ~int64_t(127);
    double center[16];
    const double Rmax2 = tc_center(s, D, center);
    if (int rc = s->sorted_pos.reserve(sizeof(float) * (size_t)(kTcK + 1) * (size_t)Npad + 256)) return rc;
    if (int rc = s->aux.reserve(sizeof(double) * 32)) return rc;
    float *opB = s->sorted_pos.as<float>();
    float *nrm_half = opB + (size_t)kTcK * (size_t)Npad;
    double *d_center = s->aux.as<double>();
    MPB_CUDA(cudaMemcpyAsync(d_center, center, sizeof(double) * D, cudaMemcpyHostToDevice, st));
    tc_prepare<D><<<(unsigned)ceil_div(Npad, 256), 256, 0, st>>>(s->V.as<double>(), N, Npad, d_center, r * r, opB, nrm_half);
    MPB_LAUNCHED();
    plan->opB = opB;
    plan->nrm_half = nrm_half;
    plan->Npad = Npad;
    plan->delta = tc_delta(r, D, Rmax2);
    return 0;
}

template <int D>
int tc_sweep(mpb200_samples *s, const TcPlan &P, double r, int64_t nq_run, int *counts, int cap, int *slab_j,
             double *slab_s, int *rcounts) {
    cudaStream_t st = ctx().stream;
    if (P.Npad >= (int64_t(1) << 27))  // queued candidates are packed as (sample << 5 | row)
        return fail(MPB200_EARG, "the tensor-core all-pairs sweep supports fewer than 2^27 samples");
    const unsigned nb = (unsigned)ceil_div(nq_run > 0 ? nq_run : 1, kTcM);
    if (rcounts) {  // symmetric: full range, q0 == 0, slabs present; rcounts = atomic counters of the partner appends
        MPB_CUDA(cudaMemsetAsync(rcounts, 0, sizeof(int) * (size_t)nq_run, st));
        tc_rball_kernel<D, 3><<<nb, kTcM, 0, st>>>(s->V.as<double>(), P.opB, P.nrm_half, s->N, P.Npad, 0, nq_run, r * r,
                                                   P.delta, counts, cap, slab_j, slab_s, rcounts);
    } else if (cap > 0)
        tc_rball_kernel<D, 2><<<nb, kTcM, 0, st>>>(s->V.as<double>(), P.opB, P.nrm_half, s->N, P.Npad, s->q0, nq_run, r * r,
                                                   P.delta, counts, cap, slab_j, slab_s, nullptr);
    else
        tc_rball_kernel<D, 0><<<nb, kTcM, 0, st>>>(s->V.as<double>(), P.opB, P.nrm_half, s->N, P.Npad, s->q0, nq_run, r * r,
                                                   P.delta, counts, 0, nullptr, nullptr, nullptr);
    MPB_LAUNCHED();
    return 0;
}

#define MPB_TC_INSTANTIATE(D_)                                                                                 \
    template int tc_prepare_operands<D_>(mpb200_samples *, double, TcPlan *);                                   \
    template int tc_sweep<D_>(mpb200_samples *, const TcPlan &, double, int64_t, int *, int, int *, double *, int *);
MPB_TC_INSTANTIATE(4) MPB_TC_INSTANTIATE(5) MPB_TC_INSTANTIATE(6) MPB_TC_INSTANTIATE(7) MPB_TC_INSTANTIATE(8)
MPB_TC_INSTANTIATE(9) MPB_TC_INSTANTIATE(10) MPB_TC_INSTANTIATE(11) MPB_TC_INSTANTIATE(12) MPB_TC_INSTANTIATE(13)
MPB_TC_INSTANTIATE(14)

}  // namespace mpb
