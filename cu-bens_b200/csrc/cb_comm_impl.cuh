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
    int (*AllGather)(const void *, void *, size_t, int, cb_ncclComm_t, cudaStream_t) = nullptr;     // optional
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
// ncclDataType_t / ncclRedOp_t values of nccl.h (stable across NCCL 2.x)
enum { CB_NCCL_CHAR = 0, CB_NCCL_INT32 = 2, CB_NCCL_FLOAT64 = 8, CB_NCCL_SUM = 0, CB_NCCL_MIN = 3 };

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
    a.AllGather = (int (*)(const void *, void *, size_t, int, cb_ncclComm_t, cudaStream_t))dlsym(l, "ncclAllGather");
    a.GetErrorString = (const char *(*)(int))dlsym(l, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce)
        return fail(CB_ERR_UNSUPPORTED, "libnccl.so.2 lacks the expected entry points");
    g_nccl = a;
    return CB_OK;
}
const char *nccl_err(int e) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "NCCL error"; }

// Peer-memory mailboxes for the fused sums + all-reduce kernel (cb_api.cu, cb_xchg_allreduce): every rank
// allocates its mailbox, the CUDA IPC handles travel through one ncclAllGather, every rank maps its peers'.
// All ranks agree (ncclAllReduce(min) of a flag) whether everybody succeeded; if not, the NCCL all-reduce
// stays the collective.  One process per GPU on one node (NVLink / NVSwitch peers); CB_COMM_P2P=0 turns it off.
void p2p_teardown(cb_handle *h)
{
    for (void *p : h->peer_open) if (p) cudaIpcCloseMemHandle(p);
    h->peer_open.clear();
    h->mbox.release(); h->peer_tab.release(); h->xerr.release();
    h->p2p = false;
}

void p2p_setup(cb_handle *h)
{
    const int world = h->comm_world, rank = h->comm_rank;
    const char *env = getenv("CB_COMM_P2P");
    int ok = !(env && env[0] == '0') && world >= 2 && world <= 8 && g_nccl.AllGather != nullptr;
    cudaStream_t s = h->stream;
    DevBuf<unsigned char> hd; DevBuf<int32_t> flag;
    std::vector<cudaIpcMemHandle_t> all((size_t)world);
    const size_t words = (size_t)2 * world * CB_X_WORDS;
    if (flag.alloc(1)) return;
    if (ok) ok = !h->mbox.alloc(words) && !h->peer_tab.alloc((size_t)world) && !h->xerr.alloc(1) &&
                 !hd.alloc((size_t)world * sizeof(cudaIpcMemHandle_t));
    if (ok) ok = cudaMemsetAsync(h->mbox.p, 0, words * sizeof(uint2), s) == cudaSuccess &&
                 cudaMemsetAsync(h->xerr.p, 0, sizeof(int32_t), s) == cudaSuccess;
    cudaIpcMemHandle_t mine{};
    if (ok) ok = cudaIpcGetMemHandle(&mine, h->mbox.p) == cudaSuccess;
    // every rank takes part in the two collectives below whatever its own state (they are collective)
    bool gathered = false;
    if (g_nccl.AllGather && hd.p) {
        cudaMemcpyAsync(hd.p + (size_t)rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice, s);
        gathered = g_nccl.AllGather(hd.p + (size_t)rank * sizeof mine, hd.p, sizeof mine, CB_NCCL_CHAR,
                                    (cb_ncclComm_t)h->comm, s) == 0 &&
                   cudaMemcpyAsync(all.data(), hd.p, (size_t)world * sizeof mine, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
                   cudaStreamSynchronize(s) == cudaSuccess;
    }
    ok = ok && gathered;
    std::vector<unsigned long long> tab((size_t)world, 0);
    h->peer_open.assign((size_t)world, nullptr);
    if (ok) {
        for (int r = 0; r < world && ok; ++r) {
            if (r == rank) { tab[r] = (unsigned long long)h->mbox.p; continue; }
            void *ptr = nullptr;
            ok = cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (ok) { h->peer_open[r] = ptr; tab[r] = (unsigned long long)ptr; }
        }
        if (!ok) cudaGetLastError();
        if (ok) ok = cudaMemcpyAsync(h->peer_tab.p, tab.data(), (size_t)world * sizeof(unsigned long long), cudaMemcpyHostToDevice, s) == cudaSuccess;
    }
    int32_t f = ok ? 1 : 0;
    cudaMemcpyAsync(flag.p, &f, sizeof f, cudaMemcpyHostToDevice, s);
    const bool agreed = g_nccl.AllReduce(flag.p, flag.p, 1, CB_NCCL_INT32, CB_NCCL_MIN, (cb_ncclComm_t)h->comm, s) == 0 &&
                        cudaMemcpyAsync(&f, flag.p, sizeof f, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
                        cudaStreamSynchronize(s) == cudaSuccess;
    hd.release(); flag.release();
    if (agreed && f == 1) { h->p2p = true; h->xseq = 0; }
    else p2p_teardown(h);
}
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
    p2p_setup(h);               // optional: peer-memory mailboxes for the fused sums + all-reduce launch
    return CB_OK;
}

// 1 when the ranks of the handle's communicator exchange through mapped peer memory, 0 when through NCCL
extern "C" int cb_comm_peer_memory(cb_handle *h) { return h && h->p2p ? 1 : 0; }

extern "C" int cb_comm_destroy(cb_handle *h)
{
    if (!h || !h->comm) return CB_OK;
    cudaSetDevice(h->fl.device);
    cudaStreamSynchronize(h->stream);
    p2p_teardown(h);
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
    if (h->p2p) {               // own exchange over NVLink peer memory (one small launch; see cb_xchg_allreduce)
        k_xchg_sums<<<1, 256, 0, h->stream>>>(h->sums.p, xchg_args(h, true));
        h->launches += 1;
        CUDA_TRY(cudaGetLastError());
        return CB_OK;
    }
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
