// NCCL plumbing inside the library: the three tiny collectives of a sharded multistart (all-gather of the best
// record, all-gather of the moment accumulator, all-reduce of the histogram) are enqueued on the SAME stream as the
// kernels, without the per-call stream/event hand-over of a framework front end.  Replaces the pickle-file rendezvous
// of consolidate_parallel_sampling_data (IP.py:1130-1155, 1209-1319).
// libnccl is resolved at run time with dlopen (the copy the process already has, e.g. the one PyTorch loaded), so the
// library carries no link-time dependency and loads on machines without NCCL.
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include "common.cuh"

typedef struct { char internal[128]; } pq_ncclUniqueId;
typedef void* pq_ncclComm_t;
typedef int (*fn_GetUniqueId)(pq_ncclUniqueId*);
typedef int (*fn_CommInitRank)(pq_ncclComm_t*, int, pq_ncclUniqueId, int);
typedef int (*fn_CommDestroy)(pq_ncclComm_t);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, pq_ncclComm_t, cudaStream_t);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, pq_ncclComm_t, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);

namespace {
struct NcclApi {
    void* handle = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
} g_nccl;

int load_nccl() {
    if (g_nccl.handle) return PEUQSE_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return pq_set_error(PEUQSE_ERR_INVALID, "libnccl.so.2 could not be loaded");
    g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (fn_AllGather)dlsym(h, "ncclAllGather");
    g_nccl.AllReduce = (fn_AllReduce)dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (fn_GetErrorString)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.AllReduce)
        return pq_set_error(PEUQSE_ERR_INVALID, "libnccl is missing a required symbol");
    g_nccl.handle = h;
    return PEUQSE_OK;
}

int nccl_err(int rc, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "NCCL error %d (%s) in %s", rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?", what);
    return pq_set_error(PEUQSE_ERR_CUDA, buf);
}
}  // namespace

struct peuqse_comm {
    pq_ncclComm_t comm;
    int rank, world;
};

extern "C" int peuqse_comm_unique_id(void* out128) {
    int rc = load_nccl();
    if (rc) return rc;
    pq_ncclUniqueId id;
    const int e = g_nccl.GetUniqueId(&id);
    if (e) return nccl_err(e, "ncclGetUniqueId");
    memcpy(out128, &id, sizeof id);
    return PEUQSE_OK;
}

extern "C" int peuqse_comm_create(const void* id128, int32_t rank, int32_t world, peuqse_comm** out) {
    int rc = load_nccl();
    if (rc) return rc;
    if (!id128 || !out || world < 1 || rank < 0 || rank >= world) return pq_set_error(PEUQSE_ERR_INVALID, "bad comm arguments");
    pq_ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    peuqse_comm* c = new peuqse_comm();
    c->rank = rank;
    c->world = world;
    const int e = g_nccl.CommInitRank(&c->comm, world, id, rank);      // uses the calling thread's current CUDA device
    if (e) { delete c; return nccl_err(e, "ncclCommInitRank"); }
    *out = c;
    return PEUQSE_OK;
}

extern "C" int peuqse_comm_destroy(peuqse_comm* c) {
    if (!c) return PEUQSE_OK;
    if (g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return PEUQSE_OK;
}

// d_recv[world][bytes] <- every rank's d_send[bytes]
extern "C" int peuqse_comm_allgather(peuqse_comm* c, const void* d_send, void* d_recv, int64_t bytes, void* stream) {
    if (!c || !d_send || !d_recv || bytes < 0) return pq_set_error(PEUQSE_ERR_INVALID, "bad allgather arguments");
    const int e = g_nccl.AllGather(d_send, d_recv, (size_t)bytes, /*ncclUint8*/ 1, c->comm, (cudaStream_t)stream);
    return e ? nccl_err(e, "ncclAllGather") : PEUQSE_OK;
}

// in-place sum of int64 counters (histograms)
extern "C" int peuqse_comm_allreduce_sum_i64(peuqse_comm* c, void* d_buf, int64_t count, void* stream) {
    if (!c || !d_buf || count < 0) return pq_set_error(PEUQSE_ERR_INVALID, "bad allreduce arguments");
    const int e = g_nccl.AllReduce(d_buf, d_buf, (size_t)count, /*ncclInt64*/ 4, /*ncclSum*/ 0, c->comm, (cudaStream_t)stream);
    return e ? nccl_err(e, "ncclAllReduce") : PEUQSE_OK;
}
