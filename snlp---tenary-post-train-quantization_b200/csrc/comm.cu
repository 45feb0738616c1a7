// The one real exchange step of the row-sharded column sweep: with SSR every block needs the
// column statistics summed over ALL rows, i.e. an all-reduce of 2*rem+1 floats across the row shards.
// The communicator lives inside the library so the C driver loop (sweep.cu) can enqueue the NCCL
// all-reduce on the caller's stream between two of its own kernels, with no return to Python.
//
// NCCL is resolved at run time from the copy already loaded into the process (torch's bundled
// libnccl.so.2), so libtq100.so has no link-time dependency on it and loads on a CPU-only box.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#define TQ_P2P_MAX_RANKS 16

namespace tq {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};

static NcclApi g_nccl = {};
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_nranks = 1;

static bool load_nccl() {
    if (g_nccl.ok) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // prefer the copy torch already mapped
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (!h) {
        set_error("NCCL not found: %s", dlerror());
        return false;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy &&
                g_nccl.GetErrorString;
    if (!g_nccl.ok) set_error("NCCL symbols missing in libnccl.so.2");
    return g_nccl.ok;
}

#define TQ_NCCL(call)                                                              \
    do {                                                                           \
        ncclResult_t r__ = (call);                                                 \
        if (r__ != ncclSuccess) {                                                  \
            set_error("%s failed: %s", #call, g_nccl.GetErrorString(r__));         \
            return TQ_E_UNSUPPORTED;                                               \
        }                                                                          \
    } while (0)

bool comm_active() { return g_comm != nullptr && g_nranks > 1; }

int comm_allreduce_sum_f32(float* buf, int64_t count, cudaStream_t st) {
    if (!comm_active()) return 0;
    TQ_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat32, ncclSum, g_comm, st));
    return 0;
}

}  // namespace tq

extern "C" int tq_comm_unique_id(void* out128_host) {
    using namespace tq;
    TQ_CHECK_ARG(out128_host, "tq_comm_unique_id: null buffer");
    if (!load_nccl()) return TQ_E_UNSUPPORTED;
    ncclUniqueId id;
    TQ_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(out128_host, &id, sizeof(id));
    return 0;
}

extern "C" int tq_comm_init(const void* id128_host, int rank, int nranks) {
    using namespace tq;
    TQ_CHECK_ARG(id128_host && nranks >= 1 && rank >= 0 && rank < nranks, "tq_comm_init: bad arguments");
    TQ_CHECK_ARG(g_comm == nullptr, "tq_comm_init: communicator already initialised");
    if (!load_nccl()) return TQ_E_UNSUPPORTED;
    ncclUniqueId id;
    memcpy(&id, id128_host, sizeof(id));
    TQ_NCCL(g_nccl.CommInitRank(&g_comm, nranks, id, rank));
    g_rank = rank;
    g_nranks = nranks;
    return 0;
}

extern "C" int tq_comm_ready(void) { return tq::g_comm != nullptr ? tq::g_nranks : 0; }

extern "C" int tq_comm_destroy(void) {
    using namespace tq;
    if (g_comm) {
        TQ_NCCL(g_nccl.CommDestroy(g_comm));
        g_comm = nullptr;
        g_nranks = 1;
        g_rank = 0;
    }
    return 0;
}

extern "C" int tq_comm_allreduce_f32(float* buf, int64_t count, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(buf && count >= 0, "tq_comm_allreduce_f32: bad arguments");
    TQ_CHECK_ARG(comm_active(), "tq_comm_allreduce_f32: communicator not initialised (tq_comm_init)");
    return comm_allreduce_sum_f32(buf, count, (cudaStream_t)stream);
}

// ---- one-shot all-reduce over peer memory (NVLink / NVSwitch) -------------------------------------------------------
// The per-block exchange of the row-sharded SSR sweep is 2*rem+1 floats (<= 110 KB): far below the size where a ring or
// tree pays off, and NCCL's launch + protocol latency (~20-30 us at 8 ranks) sits on the sweep's critical path 86 times
// per down_proj.  Here the collective is fused into the two kernels around it:
//   send   (was: ssr_fold_kernel)   folds this rank's per-chunk partials and STORES the result straight into a mailbox
//                                   slot [parity][my rank] of EVERY rank's buffer (peer memory mapped through CUDA IPC),
//                                   then the last CTA publishes a sequence number in every rank's flag array
//   gather (was: ncclAllReduce)     waits until all ranks' flags carry the sequence number (spinning on LOCAL memory),
//                                   sums the slots in rank order -- identical bits on every rank, so every rank selects
//                                   the same block -- and leaves the vector where ssr_select expects it
// Mailboxes are double-buffered by the parity of the sequence number: a rank can only be one exchange ahead of the slowest
// (it needs that rank's flag to finish its own gather), so slot [p] of exchange s+2 is never written before s was read.
namespace tq {

struct P2P {
    float* local = nullptr;            // this rank's buffer (cudaMalloc): mail[2][nranks][cap] floats, then flags[2][nranks]
    float* peer[TQ_P2P_MAX_RANKS] = {};   // every rank's buffer in this process's address space (peer[rank] == local)
    int64_t cap = 0;
    int nranks = 0, rank = 0;
    unsigned seq = 0;                  // exchanges issued so far (identical on every rank: sweeps are collective)
    unsigned* counter = nullptr;       // last-CTA-done ticket of the send kernel
    bool ready = false;
};
static P2P g_p2p;

struct P2PPeers {
    float* base[TQ_P2P_MAX_RANKS];
};

__device__ __forceinline__ unsigned* p2p_flags(float* base, int64_t cap, int nranks) {
    return reinterpret_cast<unsigned*>(base + 2 * (int64_t)nranks * cap);
}

__global__ void __launch_bounds__(256)
p2p_fold_send_kernel(const float* __restrict__ partials, int num_chunks, const float* __restrict__ rowmean, int n, int rem,
                     P2PPeers peers, int64_t cap, int nranks, int rank, unsigned seq, unsigned* __restrict__ counter) {
    const int parity = seq & 1;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t slot = ((int64_t)parity * nranks + rank) * cap;
    if (j < 2 * rem) {
        float s = 0.f;
        for (int c = 0; c < num_chunks; ++c) s += partials[(int64_t)c * 2 * rem + j];
        for (int r = 0; r < nranks; ++r) peers.base[r][slot + j] = s;
    }
    if (blockIdx.x == 0) {
        __shared__ float red[8];
        float s = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(rowmean[i], rowmean[i], s);
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
            v = warp_sum(v);
            if (threadIdx.x == 0)
                for (int r = 0; r < nranks; ++r) peers.base[r][slot + 2 * rem] = v;
        }
    }
    // publish: every CTA's stores are fenced system-wide before its ticket; the CTA that draws the last ticket raises the
    // flags (release) in every rank's buffer
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < nranks) {
            unsigned* f = p2p_flags(peers.base[threadIdx.x], cap, nranks) + parity * nranks + rank;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
        }
        if (threadIdx.x == 0) *counter = 0;           // next launch on this stream starts from zero
    }
}

__global__ void __launch_bounds__(256)
p2p_gather_kernel(float* __restrict__ local, int64_t cap, int nranks, unsigned seq, int count, float* __restrict__ out) {
    const int parity = seq & 1;
    if (threadIdx.x < nranks) {
        const unsigned* f = p2p_flags(local, cap, nranks) + parity * nranks + threadIdx.x;
        unsigned v;
        const unsigned long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int)(v - seq) >= 0) break;
            if (clock64() - t0 > 20000000000ull) __trap();          // ~10 s: a peer died; do not hang the box
        }
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < count) {
        float s = 0.f;
        for (int r = 0; r < nranks; ++r) s += local[((int64_t)parity * nranks + r) * cap + j];   // rank order: same bits everywhere
        out[j] = s;
    }
}

bool p2p_active() { return g_p2p.ready; }

// folded statistics of this rank -> sum over all ranks in `out` (2 * rem + 1 floats), on the caller's stream
int p2p_fold_allreduce(const float* partials, int64_t num_chunks, const float* rowmean, int64_t n, int64_t rem, float* out,
                       cudaStream_t st) {
    const int64_t count = 2 * rem + 1;
    if (count > g_p2p.cap) {
        set_error("p2p all-reduce: %lld floats exceed the mailbox capacity %lld", (long long)count, (long long)g_p2p.cap);
        return TQ_E_WORKSPACE;
    }
    P2PPeers peers;
    for (int r = 0; r < TQ_P2P_MAX_RANKS; ++r) peers.base[r] = g_p2p.peer[r];
    const unsigned seq = ++g_p2p.seq;
    p2p_fold_send_kernel<<<(unsigned)ceil_div(2 * rem, 256), 256, 0, st>>>(partials, (int)num_chunks, rowmean, (int)n, (int)rem,
                                                                           peers, g_p2p.cap, g_p2p.nranks, g_p2p.rank, seq,
                                                                           g_p2p.counter);
    TQ_LAUNCH_CHECK("p2p_fold_send_kernel");
    p2p_gather_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(g_p2p.local, g_p2p.cap, g_p2p.nranks, seq, (int)count, out);
    TQ_LAUNCH_CHECK("p2p_gather_kernel");
    return 0;
}

}  // namespace tq

// Allocate this rank's mailbox buffer (capacity `cap_floats` per slot) and return its CUDA IPC handle (64 bytes).
extern "C" int tq_comm_p2p_alloc(int64_t cap_floats, int rank, int nranks, void* handle64_out) {
    using namespace tq;
    TQ_CHECK_ARG(handle64_out && cap_floats > 0 && nranks >= 1 && nranks <= TQ_P2P_MAX_RANKS && rank >= 0 && rank < nranks,
                 "tq_comm_p2p_alloc: bad arguments (at most %d ranks)", TQ_P2P_MAX_RANKS);
    TQ_CHECK_ARG(g_p2p.local == nullptr, "tq_comm_p2p_alloc: already allocated");
    const size_t bytes = sizeof(float) * 2 * (size_t)nranks * cap_floats + sizeof(unsigned) * 2 * nranks + 64;
    TQ_CUDA(cudaMalloc(&g_p2p.local, bytes));
    TQ_CUDA(cudaMemset(g_p2p.local, 0, bytes));
    TQ_CUDA(cudaMalloc(&g_p2p.counter, sizeof(unsigned)));
    TQ_CUDA(cudaMemset(g_p2p.counter, 0, sizeof(unsigned)));
    TQ_CUDA(cudaDeviceSynchronize());
    g_p2p.cap = cap_floats;
    g_p2p.nranks = nranks;
    g_p2p.rank = rank;
    cudaIpcMemHandle_t h;
    TQ_CUDA(cudaIpcGetMemHandle(&h, g_p2p.local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
    memcpy(handle64_out, &h, sizeof(h));
    return 0;
}

// Map every rank's buffer (handles: nranks x 64 bytes, in rank order).  After this the row-sharded sweep exchanges its
// SSR statistics through peer memory instead of NCCL.
extern "C" int tq_comm_p2p_open(const void* handles) {
    using namespace tq;
    TQ_CHECK_ARG(handles && g_p2p.local != nullptr, "tq_comm_p2p_open: call tq_comm_p2p_alloc first");
    for (int r = 0; r < g_p2p.nranks; ++r) {
        if (r == g_p2p.rank) {
            g_p2p.peer[r] = g_p2p.local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + 64 * r, sizeof(h));
        void* p = nullptr;
        TQ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        g_p2p.peer[r] = static_cast<float*>(p);
    }
    g_p2p.ready = true;
    return 0;
}

extern "C" int tq_comm_p2p_ready(void) { return tq::g_p2p.ready ? tq::g_p2p.nranks : 0; }

