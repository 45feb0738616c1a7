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
