// Gradient all-reduce inside the library: the step's ONE collective - sum of the 431 706-float
// gradient over the time shards (the reference differentiates the SUM over rows, AR.py:228-229, so no rescaling) - is
// issued by the library itself on a side stream, one call per flow section as soon as that flow's backward kernels are
// queued, so it runs under the backward pass of the earlier flows; nma_train_step joins the side stream before the
// (replicated) Adamax update.  All of it is stream-ordered and CUDA-graph capturable (fork / join through events).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 - inside a PyTorch process that is the copy torch already
// loaded), so the library has no link-time dependency on it and a single-GPU user never touches it.
#include <dlfcn.h>
#include <string.h>
#include "nma_common.cuh"

namespace {
// the slice of nccl.h this file needs (NCCL 2.x ABI)
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t*);
typedef int (*fn_CommInitRank)(void**, int, ncclUniqueId_t, int);
typedef int (*fn_CommDestroy)(void*);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);
typedef int (*fn_GetVersion)(int*);
const int NCCL_FLOAT32 = 7, NCCL_SUM = 0;

struct Nccl {
    void* lib;
    fn_GetUniqueId GetUniqueId;
    fn_CommInitRank CommInitRank;
    fn_CommDestroy CommDestroy;
    fn_AllReduce AllReduce;
    fn_GetErrorString GetErrorString;
    fn_GetVersion GetVersion;
} g_nccl = {};

int nccl_load() {
    if (g_nccl.lib) return 0;
    const char* names[] = {getenv("NMA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n || !n[0]) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { nma_set_error("nma_comm: cannot load libnccl.so.2 (%s)", dlerror()); return -4; }
    g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_CommInitRank)dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (fn_CommDestroy)dlsym(lib, "ncclCommDestroy");
    g_nccl.AllReduce = (fn_AllReduce)dlsym(lib, "ncclAllReduce");
    g_nccl.GetErrorString = (fn_GetErrorString)dlsym(lib, "ncclGetErrorString");
    g_nccl.GetVersion = (fn_GetVersion)dlsym(lib, "ncclGetVersion");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
        nma_set_error("nma_comm: libnccl lacks a required symbol");
        return -4;
    }
    g_nccl.lib = lib;
    return 0;
}

int nccl_fail(const char* what, int rc) {
    nma_set_error("nma_comm: %s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return -4;
}

int comm_prepare(nma_handle_s* h) {
    CommState& c = h->comm;
    if (!c.side) {
        NMA_CHECK_CUDA(cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking));
        for (int i = 0; i < NMA_MAX_FLOWS + 2; ++i) NMA_CHECK_CUDA(cudaEventCreateWithFlags(&c.ev_ready[i], cudaEventDisableTiming));
        NMA_CHECK_CUDA(cudaEventCreateWithFlags(&c.ev_done, cudaEventDisableTiming));
    }
    c.pending = 0;
    return 0;
}
}  // namespace

// 128-byte NCCL unique id, to be created on one rank and handed to every rank by the host's own means
// (torch.distributed.broadcast of a byte tensor in viforssms_b200/trainer.py).
extern "C" int nma_comm_unique_id(char* out128) {
    if (!out128) { nma_set_error("nma_comm_unique_id: null argument"); return -1; }
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId_t id;
    if ((rc = g_nccl.GetUniqueId(&id))) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(out128, id.internal, 128);
    return 0;
}

// Collective: every rank calls it with the same id.  The communicator is owned by the handle.
extern "C" int nma_comm_create(nma_handle h, const char* id128, int32_t rank, int32_t world) {
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) { nma_set_error("nma_comm_create: bad argument"); return -1; }
    int rc = nccl_load();
    if (rc) return rc;
    comm_release(h);
    ncclUniqueId_t id;
    memcpy(id.internal, id128, 128);
    void* comm = nullptr;
    if ((rc = g_nccl.CommInitRank(&comm, world, id, rank))) return nccl_fail("ncclCommInitRank", rc);
    h->comm.comm = comm; h->comm.owned = 1; h->comm.world = world; h->comm.rank = rank;
    return comm_prepare(h);
}

// Adopt a communicator the caller owns (an ncclComm_t, passed as void*): SURVEY section 8b's nma_comm_init.
extern "C" int nma_comm_init(nma_handle h, void* nccl_comm) {
    if (!h) { nma_set_error("null handle"); return -1; }
    int rc = nccl_load();
    if (rc) return rc;
    comm_release(h);
    if (!nccl_comm) return 0;         // detaches
    h->comm.comm = nccl_comm; h->comm.owned = 0; h->comm.world = -1; h->comm.rank = -1;
    return comm_prepare(h);
}

void comm_release(nma_handle_s* h) {
    CommState& c = h->comm;
    if (c.side) cudaStreamSynchronize(c.side);
    if (c.comm && c.owned && g_nccl.CommDestroy) g_nccl.CommDestroy(c.comm);
    c.comm = nullptr; c.owned = 0; c.pending = 0;
}

// Call after every CUDA graph that captured this handle's collectives has been destroyed and the device is idle.
extern "C" int nma_comm_destroy(nma_handle h) {
    if (!h) return 0;
    comm_release(h);
    return 0;
}

extern "C" int nma_comm_world(nma_handle h) { return (h && h->comm.comm) ? h->comm.world : 0; }

int comm_allreduce_after(nma_handle_s* h, float* d_buf, int64_t count, int slot, cudaStream_t st) {
    CommState& c = h->comm;
    if (!c.comm || count <= 0) return 0;
    if (slot < 0 || slot >= NMA_MAX_FLOWS + 2) { nma_set_error("nma_comm: bad event slot"); return -1; }
    NMA_CHECK_CUDA(cudaEventRecord(c.ev_ready[slot], st));        // fork: the section is complete at this point of `st`
    NMA_CHECK_CUDA(cudaStreamWaitEvent(c.side, c.ev_ready[slot], 0));
    const int rc = g_nccl.AllReduce(d_buf, d_buf, (size_t)count, NCCL_FLOAT32, NCCL_SUM, c.comm, c.side);
    if (rc) return nccl_fail("ncclAllReduce", rc);
    c.pending += 1;
    return 0;
}

int comm_join(nma_handle_s* h, cudaStream_t st) {
    CommState& c = h->comm;
    if (!c.comm || !c.pending) return 0;
    NMA_CHECK_CUDA(cudaEventRecord(c.ev_done, c.side));
    NMA_CHECK_CUDA(cudaStreamWaitEvent(st, c.ev_done, 0));
    c.pending = 0;
    return 0;
}

// For callers of nma_elbo_fwd_bwd on a handle with a communicator: makes `stream` wait until the gradient
// sections that call issued have been all-reduced.
extern "C" int nma_comm_wait(nma_handle h, void* stream) {
    if (!h) { nma_set_error("null handle"); return -1; }
    return comm_join(h, (cudaStream_t)stream);
}

// all-reduce(sum) of an arbitrary fp32 device buffer on the handle's communicator, ordered after `stream` and joined
// back into it (used for the per-step scalars and by the tests)
extern "C" int nma_comm_allreduce(nma_handle h, float* d_buf, int64_t count, void* stream) {
    if (!h || !d_buf) { nma_set_error("nma_comm_allreduce: null argument"); return -1; }
    if (!h->comm.comm) return 0;
    int rc = comm_allreduce_after(h, d_buf, count, NMA_MAX_FLOWS + 1, (cudaStream_t)stream);
    if (rc) return rc;
    return comm_join(h, (cudaStream_t)stream);
}
