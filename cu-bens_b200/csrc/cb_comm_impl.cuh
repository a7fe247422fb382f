// cb_comm_impl.cuh - the per-iteration collective of an element-partitioned run, inside the C library
// (included at the end of cb_api.cu).  SURVEY.md 8(e): matrix columns and f_int of the owned joints are
// complete locally (halo elements), so the only data that must cross NVLink every iteration are the sums
// test() needs (misc.c:187-250) and the reaction resultants: eleven doubles, all-reduced by NCCL on the
// handle's stream from device-resident buffers.  ANAFLAG 3 adds the minimum of two element indices.
// NCCL is bound at run time (dlopen of libnccl.so.2; a process that already holds one - torch's - reuses
// it), so hosts that never call cb_comm_init need no NCCL at all.
#include <dlfcn.h>

namespace {
typedef struct ncclComm *cb_ncclComm_t;
typedef struct { char internal[128]; } cb_ncclUniqueId;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(cb_ncclUniqueId *) = nullptr;
    int (*CommInitRank)(cb_ncclComm_t *, int, cb_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(cb_ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, cb_ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
// ncclDataType_t / ncclRedOp_t values of nccl.h (stable across NCCL 2.x)
enum { CB_NCCL_INT32 = 2, CB_NCCL_FLOAT64 = 8, CB_NCCL_SUM = 0, CB_NCCL_MIN = 3 };

int nccl_bind()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return CB_OK;
    void *l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!l) l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!l) l = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!l) return fail(CB_ERR_UNSUPPORTED, "libnccl.so.2 not found: %s", dlerror());
    NcclApi a; a.lib = l;
    a.GetUniqueId = (int (*)(cb_ncclUniqueId *))dlsym(l, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(cb_ncclComm_t *, int, cb_ncclUniqueId, int))dlsym(l, "ncclCommInitRank");
    a.CommDestroy = (int (*)(cb_ncclComm_t))dlsym(l, "ncclCommDestroy");
    a.AllReduce = (int (*)(const void *, void *, size_t, int, int, cb_ncclComm_t, cudaStream_t))dlsym(l, "ncclAllReduce");
    a.GetErrorString = (const char *(*)(int))dlsym(l, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce)
        return fail(CB_ERR_UNSUPPORTED, "libnccl.so.2 lacks the expected entry points");
    g_nccl = a;
    return CB_OK;
}
const char *nccl_err(int e) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "NCCL error"; }
}

extern "C" int cb_comm_unique_id(void *id128)
{
    if (!id128) return fail(CB_ERR_ARG, "null argument");
    int rc = nccl_bind(); if (rc) return rc;
    cb_ncclUniqueId id;
    const int e = g_nccl.GetUniqueId(&id);
    if (e) return fail(CB_ERR_CUDA, "ncclGetUniqueId: %s", nccl_err(e));
    memcpy(id128, &id, sizeof id);
    return CB_OK;
}

extern "C" int cb_comm_init(cb_handle *h, const void *id128, int rank, int world)
{
    if (!h || !id128) return fail(CB_ERR_ARG, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(CB_ERR_ARG, "bad rank / world");
    if (h->comm) return fail(CB_ERR_ARG, "cb_comm_init: the handle already has a communicator");
    int rc = nccl_bind(); if (rc) return rc;
    cudaSetDevice(h->fl.device);
    cb_ncclUniqueId id; memcpy(&id, id128, sizeof id);
    cb_ncclComm_t c = nullptr;
    const int e = g_nccl.CommInitRank(&c, world, id, rank);
    if (e) return fail(CB_ERR_CUDA, "ncclCommInitRank: %s", nccl_err(e));
    h->comm = c; h->comm_rank = rank; h->comm_world = world;
    return CB_OK;
}

extern "C" int cb_comm_destroy(cb_handle *h)
{
    if (!h || !h->comm) return CB_OK;
    cudaSetDevice(h->fl.device);
    cudaStreamSynchronize(h->stream);
    g_nccl.CommDestroy((cb_ncclComm_t)h->comm);
    h->comm = nullptr; h->comm_world = 1; h->comm_rank = 0;
    return CB_OK;
}

// sum over the ranks of the eleven doubles cb_residual_sums left in cb_dev_sums(), in place, on the handle's
// stream (no host synchronisation); a handle without a communicator (one rank) returns at once
extern "C" int cb_residual_allreduce(cb_handle *h)
{
    if (!h || !h->sums.p) return fail(CB_ERR_ARG, "cb_residual_sums has not been called");
    if (!h->comm || h->comm_world == 1) return CB_OK;
    cudaSetDevice(h->fl.device);
    const int e = g_nccl.AllReduce(h->sums.p, h->sums.p, CB_NSUMS, CB_NCCL_FLOAT64, CB_NCCL_SUM, (cb_ncclComm_t)h->comm, h->stream);
    if (e) return fail(CB_ERR_CUDA, "ncclAllReduce: %s", nccl_err(e));
    return CB_OK;
}

// ANAFLAG 3 across ranks (fact 0.8): the lowest global index of a frame / shell that trips, agreed by all
// ranks, between cb_update_forces_begin and cb_update_forces_end
extern "C" int cb_trip_allreduce(cb_handle *h, int *first_fr, int *first_sh)
{
    if (!h || !first_fr || !first_sh) return fail(CB_ERR_ARG, "null argument");
    if (!h->comm || h->comm_world == 1) return CB_OK;
    cudaSetDevice(h->fl.device);
    if (!h->trip_buf.p && h->trip_buf.alloc(2)) return CB_ERR_CUDA;
    const int32_t v[2] = {*first_fr, *first_sh};
    CUDA_TRY(cudaMemcpyAsync(h->trip_buf.p, v, sizeof v, cudaMemcpyHostToDevice, h->stream));
    const int e = g_nccl.AllReduce(h->trip_buf.p, h->trip_buf.p, 2, CB_NCCL_INT32, CB_NCCL_MIN, (cb_ncclComm_t)h->comm, h->stream);
    if (e) return fail(CB_ERR_CUDA, "ncclAllReduce: %s", nccl_err(e));
    int32_t o[2];
    CUDA_TRY(cudaMemcpyAsync(o, h->trip_buf.p, sizeof o, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *first_fr = o[0]; *first_sh = o[1];
    return CB_OK;
}
