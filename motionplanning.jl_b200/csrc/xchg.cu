// xchg.cu -- the one exchange step of the sharded precompute, as direct peer stores over NVLink.
//
// Every GPU of the box owns a query-column shard of the neighbour table; the host planner (one FMT*
// wavefront, sequential -- fmt.jl:44-100) needs the GLOBAL column pointer and the GLOBAL edge-validity
// BitVector.  Round 1 did this with three torch elementwise kernels + one NCCL all-gather per step,
// which was latency-bound (0.39 ms of a 1.0 ms step on 8 GPUs).  Here each rank's receive buffer is a
// plain cudaMalloc block exported with CUDA IPC and mapped by every peer (NVSwitch: every peer at full
// bandwidth), and ONE kernel per step
//   * turns the shard's Int64 colptr into int32 column lengths on the fly,
//   * stores them and the validity words straight into slot `rank` of EVERY peer's buffer
//     (16-byte coalesced stores; the source is read once),
// followed by a one-warp flag barrier: release-store of the epoch into every peer's flag word, acquire
// spin on the own flag words (bounded; a peer that never arrives raises an error flag instead of
// hanging the GPU).  Two buffer sets alternate by epoch parity, so a rank may still be reading epoch e
// while its peers already push epoch e + 1.
#include "common.cuh"
#include <algorithm>
#include <new>

struct mpb200_xchg {
    int rank = 0, world = 1;
    int64_t max_ncols = 0, word_cap = 0;
    int64_t counts_off = 0, words_off = 0, slot_bytes = 0, set_bytes = 0, data_off = 0;
    char *local = nullptr;              // own block: [flags | set 0 | set 1]
    char *peer[MPB200_XCHG_MAX_WORLD] = {};  // mapped blocks of every rank (peer[rank] == local)
    bool connected = false;
    unsigned long long epoch = 0;       // epochs pushed so far
    // early push of the column lengths (mpb200_xchg_attach): they are final after the count scan, so they travel on
    // a side stream underneath the fill and validity kernels; mpb200_xchg_push then only sends the validity words
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    unsigned long long counts_epoch = 0;  // epoch whose column lengths are already on their way
    const mpb200_table *counts_table = nullptr;
};

namespace mpb {

constexpr int kFlagBytes = 4096;  // flags[world] (epoch of the last complete push of each rank) + status word
constexpr int kStatusWord = 64;   // index (in 8-byte words) of the error flag

struct PeerPtrs { char *p[MPB200_XCHG_MAX_WORLD]; };

// header (first 64 bytes of a slot): ncols, nnz, epoch.  what: bit 0 = column lengths, bit 1 = header + validity words
__global__ void __launch_bounds__(256)
xchg_push_kernel(const int64_t *__restrict__ colptr, int64_t ncols, const uint4 *__restrict__ words16, int64_t n_words,
                 int64_t nnz, unsigned long long epoch, PeerPtrs dst, int world, int64_t counts_off, int64_t words_off,
                 int what) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    if (tid == 0 && (what & 2)) {
        for (int p = 0; p < world; ++p) {
            long long *h = reinterpret_cast<long long *>(dst.p[p]);
            h[0] = ncols; h[1] = nnz; h[2] = (long long)epoch;
        }
    }
    // column lengths: four int32 per 16-byte store
    const int64_t n4 = (what & 1) ? ((ncols + 3) >> 2) : 0;
    for (int64_t i = tid; i < n4; i += nthr) {
        int c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t j = 4 * i + k;
            c[k] = j < ncols ? (int)(colptr[j + 1] - colptr[j]) : 0;
        }
        const uint4 v = make_uint4((unsigned)c[0], (unsigned)c[1], (unsigned)c[2], (unsigned)c[3]);
        for (int p = 0; p < world; ++p) reinterpret_cast<uint4 *>(dst.p[p] + counts_off)[i] = v;
    }
    // validity words: two uint64 per 16-byte store (the table's buffer is padded to a whole number of them)
    const int64_t n16 = (what & 2) ? ((n_words + 1) >> 1) : 0;
    for (int64_t i = tid; i < n16; i += nthr) {
        uint4 v = words16[i];
        if (2 * i + 1 >= n_words) { v.z = 0; v.w = 0; }
        for (int p = 0; p < world; ++p) reinterpret_cast<uint4 *>(dst.p[p] + words_off)[i] = v;
    }
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// one warp: lane p announces this rank's epoch in peer p's flag block, then waits for peer p's
// announcement in the own flag block.  Stream order puts it after the push kernel; the fence + release
// make the pushed data visible before the flag.
__global__ void __launch_bounds__(32)
xchg_barrier_kernel(PeerPtrs flags, unsigned long long *__restrict__ own_flags, int rank, int world,
                    unsigned long long epoch, long long timeout_ns) {
    const int lane = threadIdx.x;
    __threadfence_system();
    if (lane < world) st_release_sys(reinterpret_cast<unsigned long long *>(flags.p[lane]) + rank, epoch);
    if (lane < world) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(own_flags + lane) < epoch) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if ((long long)(t1 - t0) > timeout_ns) {  // a peer never arrived: report, do not hang the GPU
                own_flags[kStatusWord] = 1ULL + (unsigned long long)lane;
                break;
            }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

}  // namespace mpb

namespace mpb {
static void xchg_slots(const mpb200_xchg *x, unsigned long long epoch, PeerPtrs *dst, PeerPtrs *flags) {
    const int64_t set = x->data_off + (int64_t)(epoch & 1) * x->set_bytes + (int64_t)x->rank * x->slot_bytes;
    for (int p = 0; p < MPB200_XCHG_MAX_WORLD; ++p) {
        dst->p[p] = p < x->world ? x->peer[p] + set : nullptr;
        flags->p[p] = p < x->world ? x->peer[p] : nullptr;
    }
}

// Called by the table builds right after the column-pointer scan (the table is attached to an exchange): the
// column lengths of the UPCOMING epoch go out on the side stream while the main stream runs the fill.
int xchg_push_counts_early(mpb200_xchg *x, const mpb200_table *t, const int64_t *colptr, int64_t ncols) {
    if (!x || !(x->connected || x->world == 1) || ncols > x->max_ncols) return 0;  // push() reports misuse
    Context &c = ctx();
    const unsigned long long epoch = x->epoch + 1;
    PeerPtrs dst, flags;
    xchg_slots(x, epoch, &dst, &flags);
    MPB_CUDA(cudaEventRecord(x->ev_fork, c.stream));
    MPB_CUDA(cudaStreamWaitEvent(x->side, x->ev_fork, 0));
    const unsigned grid = (unsigned)std::min<int64_t>(std::max<int64_t>(ceil_div((ncols + 3) / 4, 256), 1), (int64_t)c.sm_count);
    xchg_push_kernel<<<grid, 256, 0, x->side>>>(colptr, ncols, nullptr, 0, 0, epoch, dst, x->world, x->counts_off,
                                                x->words_off, 1);
    MPB_LAUNCHED();
    MPB_CUDA(cudaEventRecord(x->ev_join, x->side));
    x->counts_epoch = epoch;
    x->counts_table = t;
    return 0;
}
}  // namespace mpb

using namespace mpb;

extern "C" {

int mpb200_xchg_create(int rank, int world, int64_t max_ncols, int64_t word_cap, mpb200_xchg **out, void *ipc_handle) {
    MPB_REQUIRE_INIT();
    MPB_CHECK_ARG(out && ipc_handle, "NULL argument");
    MPB_CHECK_ARG(world >= 1 && world <= MPB200_XCHG_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world size");
    MPB_CHECK_ARG(max_ncols >= 0 && word_cap >= 0, "negative capacity");
    static_assert(sizeof(cudaIpcMemHandle_t) == MPB200_IPC_HANDLE_BYTES, "IPC handle size");
    mpb200_xchg *x = new (std::nothrow) mpb200_xchg();
    if (!x) return fail(MPB200_ENOMEM, "out of host memory");
    x->rank = rank;
    x->world = world;
    x->max_ncols = max_ncols;
    x->word_cap = word_cap;
    x->counts_off = 64;
    x->words_off = x->counts_off + ((4 * max_ncols + 15) / 16) * 16;
    x->slot_bytes = ((x->words_off + ((8 * word_cap + 15) / 16) * 16 + 255) / 256) * 256;
    x->set_bytes = x->slot_bytes * world;
    x->data_off = kFlagBytes;
    // a dedicated cudaMalloc block (not the block cache: IPC exports whole allocations)
    cudaError_t e = cudaMalloc(&x->local, (size_t)(kFlagBytes + 2 * x->set_bytes));
    if (e != cudaSuccess) { delete x; return fail(MPB200_ENOMEM, "cudaMalloc of the exchange buffer failed: %s", cudaGetErrorString(e)); }
    e = cudaMemset(x->local, 0, (size_t)(kFlagBytes + 2 * x->set_bytes));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // zeroed flags before any peer can map the block
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(ipc_handle), x->local);
    if (e != cudaSuccess) {
        cudaFree(x->local);
        delete x;
        return fail(MPB200_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    x->peer[rank] = x->local;
    if (cudaStreamCreateWithFlags(&x->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&x->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&x->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        mpb200_xchg_destroy(x);
        return fail(MPB200_ECUDA, "could not create the exchange's side stream");
    }
    *out = x;
    return MPB200_OK;
}

int mpb200_xchg_attach(mpb200_xchg *x, mpb200_table *t) {
    MPB_CHECK_ARG(t != nullptr, "table handle is NULL");
    t->xchg = x;  // NULL detaches
    return MPB200_OK;
}

int mpb200_xchg_connect(mpb200_xchg *x, const void *handles) {
    MPB_REQUIRE_INIT();
    MPB_CHECK_ARG(x && handles, "NULL argument");
    if (x->connected) return MPB200_OK;
    const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(handles);
    for (int p = 0; p < x->world; ++p) {
        if (p == x->rank) continue;
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h[p], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(MPB200_ECUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s (no peer access between these GPUs?)", p,
                        cudaGetErrorString(e));
        }
        x->peer[p] = static_cast<char *>(ptr);
    }
    x->connected = true;
    return MPB200_OK;
}

int mpb200_xchg_push(mpb200_xchg *x, const mpb200_table *t) {
    MPB_REQUIRE_INIT();
    MPB_CHECK_ARG(x && t, "NULL argument");
    MPB_CHECK_ARG(x->connected || x->world == 1, "mpb200_xchg_connect has not been called");
    MPB_CHECK_ARG(t->ncols <= x->max_ncols, "table has more columns than the exchange was created for");
    const int64_t n_words = ceil_div(t->nnz, 64);
    MPB_CHECK_ARG(n_words <= x->word_cap, "table has more validity words than the exchange was created for");
    if (t->nnz > 0 && !(t->edge_bits_valid && t->edge_bits_nnz == t->nnz))
        return fail(MPB200_ESTATE, "no edge validity has been computed for the current contents of this table");
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const unsigned long long epoch = ++x->epoch;
    PeerPtrs dst, flags;
    xchg_slots(x, epoch, &dst, &flags);
    phase_bank(MPB200_OP_OTHER);
    phase_mark(0);
    // column lengths already under way on the side stream (attached table, pushed right after its count scan)?
    const bool early = x->counts_epoch == epoch && x->counts_table == t;
    if (early) MPB_CUDA(cudaStreamWaitEvent(st, x->ev_join, 0));
    const int64_t units = std::max<int64_t>(early ? 0 : (t->ncols + 3) / 4, (n_words + 1) / 2);
    const unsigned grid = (unsigned)std::min<int64_t>(std::max<int64_t>(ceil_div(units, 256), 1), (int64_t)c.sm_count * 4);
    xchg_push_kernel<<<grid, 256, 0, st>>>(t->colptr.as<int64_t>(), t->ncols, t->edge_bits.as<uint4>(), n_words, t->nnz,
                                           epoch, dst, x->world, x->counts_off, x->words_off, early ? 2 : 3);
    MPB_LAUNCHED();
    xchg_barrier_kernel<<<1, 32, 0, st>>>(flags, reinterpret_cast<unsigned long long *>(x->local), x->rank, x->world,
                                          epoch, 20LL * 1000 * 1000 * 1000);
    MPB_LAUNCHED();
    phase_mark(1);
    phases_collect(1);
    return MPB200_OK;
}

int mpb200_xchg_view(const mpb200_xchg *x, void **recv, int64_t *slot_bytes, int64_t *counts_off, int64_t *words_off,
                     int64_t *status) {
    MPB_REQUIRE_INIT();
    MPB_CHECK_ARG(x != nullptr, "exchange handle is NULL");
    if (recv) *recv = x->local + x->data_off + (int64_t)(x->epoch & 1) * x->set_bytes;  // set of the newest epoch
    if (slot_bytes) *slot_bytes = x->slot_bytes;
    if (counts_off) *counts_off = x->counts_off;
    if (words_off) *words_off = x->words_off;
    if (status) {  // waits for the pushes enqueued so far; non-zero: 1 + the rank that never arrived
        unsigned long long s = 0;
        cudaStream_t st = ctx().stream;
        MPB_CUDA(cudaMemcpyAsync(&s, reinterpret_cast<unsigned long long *>(x->local) + kStatusWord, sizeof(s),
                                 cudaMemcpyDeviceToHost, st));
        MPB_CUDA(cudaStreamSynchronize(st));
        *status = (int64_t)s;
    }
    return MPB200_OK;
}

int mpb200_xchg_destroy(mpb200_xchg *x) {
    if (!x) return MPB200_OK;
    if (ctx().ready) cudaStreamSynchronize(ctx().stream);
    for (int p = 0; p < x->world; ++p)
        if (p != x->rank && x->peer[p]) cudaIpcCloseMemHandle(x->peer[p]);
    if (x->side) { cudaStreamSynchronize(x->side); cudaStreamDestroy(x->side); }
    if (x->ev_fork) cudaEventDestroy(x->ev_fork);
    if (x->ev_join) cudaEventDestroy(x->ev_join);
    if (x->local) cudaFree(x->local);
    delete x;
    return MPB200_OK;
}

}  // extern "C"
