// cb_plan_device.cuh - the sorted element-to-nonzero map and the CSC pattern built ON THE DEVICE (north_star:
// "assembly into a CSC pattern precomputed once on the device ... segmented reduction over a sorted
// element-to-nonzero map"; reference counterpart: codes / skylin, model.c:937-1142, 1204-1281, and the dense scan
// of solve.c:110-119).  Included by cb_api.cu; used for shell-only models with the CSC layout, the others keep the
// host builder (build_plan), which also serves as the checker: tests compare the two bit for bit.
//
//   corners        (joint, element, local joint) of every element corner, radix-sorted by joint (stable: element
//                  order kept = the order the reference adds contributions in) -> node_cstart / corners
//   contributions  (B, A, element, a, b) for every ordered pair of an element's joints with free DOFs and B owned,
//                  radix-sorted by (B, A) -> the contribution list of every joint-pair block, in reference order
//   blocks         run-length encoding of the sorted keys -> one CbPair per (A, B), its contribution range
//   CSC geometry   column height of joint B = sum of the free DOFs of its neighbours (per-joint loop), first Ax
//                  index of every joint by an exclusive scan, Ap and Ai written by one thread per joint
//   tiles          cb_plan_pack.h, the code the host planner runs, one thread per segment of CB_PK_SEG joints:
//                  count pass, exclusive scan of the counts, emit pass
// Only the joint adjacency, the per-joint column geometry and a few counters cross PCIe afterwards (the host
// keeps them for cb_csc_pattern and the symmetric hand-off).
#include <cub/cub.cuh>

namespace devplan {

__global__ void k_gen_corners(long ne, const int32_t *__restrict__ nodes, uint32_t *__restrict__ keys, uint64_t *__restrict__ vals)
{
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= ne * 3) return;
    const long e = i / 3; const int a = (int)(i - e * 3);
    keys[i] = (uint32_t)nodes[e * 4 + a];
    CbCorner c; c.e = (int32_t)e; c.type = CB_T_SHELL; c.b = (uint8_t)a; c.pad[0] = c.pad[1] = 0;
    uint64_t v; memcpy(&v, &c, 8);
    vals[i] = v;
}
// first index with keys[idx] >= j, for every j in [0, n]
template <typename K>
__global__ void k_lower_bound(long nq, const K *__restrict__ keys, long nk, int shift, int32_t *__restrict__ out)
{
    const long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j > nq) return;
    long lo = 0, hi = nk;
    while (lo < hi) { const long mid = (lo + hi) >> 1; if ((long)(keys[mid] >> shift) < j) lo = mid + 1; else hi = mid; }
    out[j] = (int32_t)lo;
}
__global__ void k_gen_contribs(long ne, const int32_t *__restrict__ nodes, const int32_t *__restrict__ nfree, long j0, long j1,
                               uint64_t *__restrict__ keys, uint64_t *__restrict__ vals)
{
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= ne * 9) return;
    const long e = i / 9; const int ab = (int)(i - e * 9), a = ab / 3, b = ab - 3 * a;
    const int32_t A = nodes[e * 4 + a], B = nodes[e * 4 + b];
    const bool ok = nfree[A] > 0 && nfree[B] > 0 && B >= j0 && B < j1;
    keys[i] = ok ? (((uint64_t)(uint32_t)B << 32) | (uint32_t)A) : ~0ull;
    CbContrib c; c.e = (int32_t)e; c.type = CB_T_SHELL; c.a = (uint8_t)a; c.b = (uint8_t)b; c.pad = 0;
    uint64_t v; memcpy(&v, &c, 8);
    vals[i] = v;
}
// column height and row offsets of the blocks of every joint
__global__ void k_joint_geometry(long NJ, const int32_t *__restrict__ jpair, const uint64_t *__restrict__ ukeys,
                                 const int32_t *__restrict__ nfree, int32_t *__restrict__ colh, int32_t *__restrict__ rowoff,
                                 int64_t *__restrict__ width)
{
    const long B = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (B >= NJ) return;
    int h = 0;
    for (int q = jpair[B]; q < jpair[B + 1]; ++q) { rowoff[q] = h; h += nfree[(uint32_t)ukeys[q]]; }
    colh[B] = h;
    width[B] = (int64_t)nfree[B] * h;
}
__global__ void k_fill_pairs(long np, const uint64_t *__restrict__ ukeys, const int32_t *__restrict__ counts,
                             const int32_t *__restrict__ cstart, const int32_t *__restrict__ rowoff,
                             const int64_t *__restrict__ base, const int32_t *__restrict__ colh, const int32_t *__restrict__ first,
                             const uint8_t *__restrict__ mask, CbPair *__restrict__ pairs, unsigned long long *__restrict__ parity)
{
    const long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (q >= np) return;
    const uint32_t A = (uint32_t)ukeys[q], B = (uint32_t)(ukeys[q] >> 32);
    CbPair p;
    p.off = (int32_t)(base[B] + rowoff[q]); p.colh = colh[B]; p.cstart = cstart[q];
    p.eqA0 = first[A]; p.eqB0 = first[B]; p.ccount = (uint16_t)counts[q]; p.maskA = mask[A]; p.maskB = mask[B];
    pairs[q] = p;
    if (!(p.colh & 1)) atomicAdd(parity + (p.off & 1), (unsigned long long)p.ccount);   // integer counters: order-free
}
__global__ void k_pattern(long NJ, long j0, long j1, long nnz, const int32_t *__restrict__ jpair, const int32_t *__restrict__ adj,
                          const int32_t *__restrict__ nfree, const int32_t *__restrict__ first, const int64_t *__restrict__ base,
                          const int32_t *__restrict__ colh, int *__restrict__ Ap, int *__restrict__ Ai)
{
    const long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j >= NJ) return;
    const bool own = j >= j0 && j < j1;
    for (int cc = 0; cc < nfree[j]; ++cc) {
        long p = own ? base[j] + (long)cc * colh[j] : (j < j0 ? 0 : nnz);
        Ap[first[j] - 1 + cc] = (int)p;
        if (!Ai || !own) continue;
        for (int q = jpair[j]; q < jpair[j + 1]; ++q) {
            const int32_t A = adj[q];
            for (int rr = 0; rr < nfree[A]; ++rr) Ai[p++] = first[A] - 1 + rr;
        }
    }
}
__global__ void k_own_bits(long ne, const int32_t *__restrict__ nodes, const int32_t *__restrict__ cstart,
                           const CbCorner *__restrict__ corners, uint8_t *__restrict__ own)
{
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    unsigned m = 0;
    for (int a = 0; a < 3; ++a) {
        const CbCorner c = corners[cstart[nodes[e * 4 + a]]];
        if (c.type == CB_T_SHELL && c.e == e && c.b == a) m |= 1u << a;
    }
    own[e] = (uint8_t)m;
}
__global__ void k_pack_count(PkIn in, long j0, long j1, long nseg, long *__restrict__ counts)
{
    const long s = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const long a0 = j0 + s * CB_PK_SEG, a1 = a0 + CB_PK_SEG < j1 ? a0 + CB_PK_SEG : j1;
    counts[s] = pk_segment(in, a0, a1, 0, nullptr);
}
__global__ void k_pack_emit(PkIn in, long j0, long j1, long nseg, const long *__restrict__ t0, PkOut out)
{
    const long s = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const long a0 = j0 + s * CB_PK_SEG, a1 = a0 + CB_PK_SEG < j1 ? a0 + CB_PK_SEG : j1;
    pk_segment(in, a0, a1, t0[s], &out);
}

struct IsValidKey { __host__ __device__ int operator()(const uint64_t &k) const { return k != ~0ull; } };
__global__ void k_adjacency(long np, const uint64_t *__restrict__ ukeys, int32_t *__restrict__ adj)
{
    const long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (q < np) adj[q] = (int32_t)(uint32_t)ukeys[q];
}
template <typename T> struct Tmp {       // scratch that lives for the duration of the build
    T *p = nullptr;
    int alloc(size_t n) { return n ? cudaMalloc((void **)&p, n * sizeof(T)) != cudaSuccess : 0; }
    ~Tmp() { if (p) cudaFree(p); }
};
static inline unsigned grid_of(long n, int tpb = 256) { return (unsigned)((n + tpb - 1) / tpb); }

}  // namespace devplan

// returns CB_OK (plan ready), a negative value when the model is not one the device builder handles (the
// caller runs the host builder), or an error code
static int build_plan_device(cb_handle *h)
{
    using namespace devplan;
    const long NJ = h->sz.NJ, SH = h->sz.NE_SH;
    const char *kt_env = getenv("CB_KT"), *pl_env = getenv("CB_PLAN");
    if (g_host_only || (pl_env && strcmp(pl_env, "host") == 0) || (kt_env && strcmp(kt_env, "duo") == 0)) return -1;
    if (h->layout != CB_MAT_CSC || !SH || h->sz.NE_TR || h->sz.NE_FR || h->NE_BR || h->fl.ANAFLAG == 3) return -1;
    for (long j = 0; j < NJ; ++j) if (h->h_mask[j] >> 6) return -1;            // a free seventh DOF: general kernels
    if (SH * 9 > 0x7fffffffL || NJ >= (1L << 31)) return -1;
    const auto t_begin = std::chrono::steady_clock::now();
    cudaStream_t s = h->stream;
    const CbStreamShape shapes[5] = {CB_S_SHAPE_WIDE, CB_S_SHAPE_NARROW, CB_S_SHAPE_MINI, CB_S_SHAPE_NARROW12, CB_S_SHAPE_MINI};
    // default: the narrow plan on 12 warps (CB_KT = wide | narrow | mini | mini16 select the other compiled kernels)
    const int shape_id = !kt_env ? 3 : (strcmp(kt_env, "wide") == 0 ? 0 : (strcmp(kt_env, "narrow") == 0 ? 1 : (strcmp(kt_env, "mini") == 0 ? 2 : (strcmp(kt_env, "mini16") == 0 ? 4 : 3))));
    const CbStreamShape shp = shapes[shape_id];
#define DP_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(CB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

    // per-joint tables
    DevBuf<int32_t> d_nfree, d_first; DevBuf<uint8_t> d_mask;
    struct Rel { DevBuf<int32_t> *a, *b; DevBuf<uint8_t> *c; ~Rel() { a->release(); b->release(); c->release(); } } rel{&d_nfree, &d_first, &d_mask};
    if (d_nfree.upload(h->h_nfree) || d_first.upload(h->h_first) || d_mask.upload(h->h_mask)) return CB_ERR_CUDA;
    int jbits = 1; while ((1L << jbits) < NJ) ++jbits;

    // ---- corners: sorted by joint ------------------------------------------------------------------
    const long nc = SH * 3;
    Tmp<uint32_t> ck_in, ck_out; Tmp<uint64_t> cv_in;
    if (ck_in.alloc(nc) || ck_out.alloc(nc) || cv_in.alloc(nc) || h->corners.alloc(nc) || h->node_cstart.alloc(NJ + 1)) return CB_ERR_CUDA;
    k_gen_corners<<<grid_of(nc), 256, 0, s>>>(SH, h->sh_nodes.p, ck_in.p, cv_in.p);
    {
        size_t tb = 0;
        DP_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, ck_in.p, ck_out.p, cv_in.p, (uint64_t *)h->corners.p, (int)nc, 0, jbits, s));
        Tmp<unsigned char> tmp; if (tmp.alloc(tb)) return CB_ERR_CUDA;
        DP_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, tb, ck_in.p, ck_out.p, cv_in.p, (uint64_t *)h->corners.p, (int)nc, 0, jbits, s));
        k_lower_bound<uint32_t><<<grid_of(NJ + 1), 256, 0, s>>>(NJ, ck_out.p, nc, 0, h->node_cstart.p);
        DP_TRY(cudaStreamSynchronize(s));
    }
    // ---- contributions: sorted by (B, A) -------------------------------------------------------------
    const long nall = SH * 9;
    Tmp<uint64_t> kk_in, kk_out, kv_in, kv_out;
    if (kk_in.alloc(nall) || kk_out.alloc(nall) || kv_in.alloc(nall) || kv_out.alloc(nall)) return CB_ERR_CUDA;
    k_gen_contribs<<<grid_of(nall), 256, 0, s>>>(SH, h->sh_nodes.p, d_nfree.p, h->j0, h->j1, kk_in.p, kv_in.p);
    {
        size_t tb = 0;
        DP_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, kk_in.p, kk_out.p, kv_in.p, kv_out.p, (int)nall, 0, 64, s));
        Tmp<unsigned char> tmp; if (tmp.alloc(tb)) return CB_ERR_CUDA;
        DP_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, tb, kk_in.p, kk_out.p, kv_in.p, kv_out.p, (int)nall, 0, 64, s));
    }
    // number of valid contributions (the others carry the all-ones key and sort to the end)
    int32_t nvalid = 0;
    {
        // binary search on the host side of a tiny probe kernel would cost a launch per step; instead count with cub
        Tmp<int32_t> cnt; if (cnt.alloc(1)) return CB_ERR_CUDA;
        size_t tb = 0;
        cub::TransformInputIterator<int, IsValidKey, const uint64_t *> it(kk_out.p, IsValidKey());
        DP_TRY(cub::DeviceReduce::Sum(nullptr, tb, it, cnt.p, (int)nall, s));
        Tmp<unsigned char> tmp; if (tmp.alloc(tb)) return CB_ERR_CUDA;
        DP_TRY(cub::DeviceReduce::Sum(tmp.p, tb, it, cnt.p, (int)nall, s));
        DP_TRY(cudaMemcpyAsync(&nvalid, cnt.p, sizeof nvalid, cudaMemcpyDeviceToHost, s));
        DP_TRY(cudaStreamSynchronize(s));
    }
    if (nvalid <= 0) return -1;
    // ---- blocks: run-length encoding of the keys ---------------------------------------------------------
    Tmp<uint64_t> ukeys; Tmp<int32_t> counts, cstart, d_np;
    if (ukeys.alloc(nvalid) || counts.alloc(nvalid) || cstart.alloc(nvalid + 1) || d_np.alloc(1)) return CB_ERR_CUDA;
    int32_t np = 0;
    {
        size_t tb = 0;
        DP_TRY(cub::DeviceRunLengthEncode::Encode(nullptr, tb, kk_out.p, ukeys.p, counts.p, d_np.p, nvalid, s));
        Tmp<unsigned char> tmp; if (tmp.alloc(tb)) return CB_ERR_CUDA;
        DP_TRY(cub::DeviceRunLengthEncode::Encode(tmp.p, tb, kk_out.p, ukeys.p, counts.p, d_np.p, nvalid, s));
        DP_TRY(cudaMemcpyAsync(&np, d_np.p, sizeof np, cudaMemcpyDeviceToHost, s));
        DP_TRY(cudaStreamSynchronize(s));
        size_t tb2 = 0;
        DP_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb2, counts.p, cstart.p, np, s));
        Tmp<unsigned char> tmp2; if (tmp2.alloc(tb2)) return CB_ERR_CUDA;
        DP_TRY(cub::DeviceScan::ExclusiveSum(tmp2.p, tb2, counts.p, cstart.p, np, s));
        DP_TRY(cudaStreamSynchronize(s));
    }
    // ---- CSC geometry -----------------------------------------------------------------------------------
    Tmp<int32_t> jpair, colh, rowoff; Tmp<int64_t> width, base;
    if (jpair.alloc(NJ + 1) || colh.alloc(NJ) || rowoff.alloc(np) || width.alloc(NJ + 1) || base.alloc(NJ + 1)) return CB_ERR_CUDA;
    k_lower_bound<uint64_t><<<grid_of(NJ + 1), 256, 0, s>>>(NJ, ukeys.p, np, 32, jpair.p);
    DP_TRY(cudaMemsetAsync(width.p, 0, (NJ + 1) * sizeof(int64_t), s));
    k_joint_geometry<<<grid_of(NJ), 256, 0, s>>>(NJ, jpair.p, ukeys.p, d_nfree.p, colh.p, rowoff.p, width.p);
    {
        size_t tb = 0;
        DP_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, width.p, base.p, (int)(NJ + 1), s));
        Tmp<unsigned char> tmp; if (tmp.alloc(tb)) return CB_ERR_CUDA;
        DP_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, width.p, base.p, (int)(NJ + 1), s));
        DP_TRY(cudaStreamSynchronize(s));
    }
    h->base.assign(NJ + 1, 0); h->colh.assign(NJ, 0);
    DP_TRY(cudaMemcpy(h->base.data(), base.p, (NJ + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    DP_TRY(cudaMemcpy(h->colh.data(), colh.p, NJ * sizeof(int32_t), cudaMemcpyDeviceToHost));
    h->ax_base = h->base[h->j0];
    const long nnz = h->base[h->j1] - h->base[h->j0];
    if (nnz > 0x7fffffffL) return fail(CB_ERR_OVERFLOW, "nnz=%ld exceeds the 32-bit CSC indices umfpack_di_* takes", nnz);
    h->nnz = nnz;
    Tmp<CbPair> pairs; Tmp<unsigned long long> parity;
    if (pairs.alloc(np) || parity.alloc(2)) return CB_ERR_CUDA;
    DP_TRY(cudaMemsetAsync(parity.p, 0, 2 * sizeof(unsigned long long), s));
    k_fill_pairs<<<grid_of(np), 256, 0, s>>>(np, ukeys.p, counts.p, cstart.p, rowoff.p, base.p, colh.p, d_first.p, d_mask.p,
                                              pairs.p, parity.p);
    unsigned long long par[2] = {0, 0};
    DP_TRY(cudaMemcpyAsync(par, parity.p, sizeof par, cudaMemcpyDeviceToHost, s));
    DP_TRY(cudaStreamSynchronize(s));
    h->ax_pad = par[1] > par[0] ? 1 : 0;
    // pattern: Ap now; the joint adjacency stays on the device so that cb_dev_Ai can write Ai when asked
    if (h->Ap.alloc(h->sz.NEQ + 1) || h->Ax.alloc((size_t)nnz + 1) || h->d_adj.alloc(np) || h->d_jpair.alloc(NJ + 1) ||
        h->d_nfree.alloc(NJ) || h->d_first.alloc(NJ) || h->d_base.alloc(NJ + 1) || h->d_colh.alloc(NJ))
        return CB_ERR_CUDA;
    dev_zero(h->Ax.p, ((size_t)nnz + 1) * sizeof(double));
    k_adjacency<<<grid_of(np), 256, 0, s>>>(np, ukeys.p, h->d_adj.p);
    DP_TRY(cudaMemcpyAsync(h->d_jpair.p, jpair.p, (NJ + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    DP_TRY(cudaMemcpyAsync(h->d_nfree.p, d_nfree.p, NJ * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    DP_TRY(cudaMemcpyAsync(h->d_first.p, d_first.p, NJ * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    DP_TRY(cudaMemcpyAsync(h->d_base.p, base.p, (NJ + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    DP_TRY(cudaMemcpyAsync(h->d_colh.p, colh.p, NJ * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    {
        k_pattern<<<grid_of(NJ), 256, 0, s>>>(NJ, h->j0, h->j1, nnz, h->d_jpair.p, h->d_adj.p, h->d_nfree.p, h->d_first.p,
                                               h->d_base.p, h->d_colh.p, h->Ap.p, nullptr);
        const int last = (int)nnz;
        DP_TRY(cudaMemcpyAsync(h->Ap.p + h->sz.NEQ, &last, sizeof last, cudaMemcpyHostToDevice, s));
        DP_TRY(cudaStreamSynchronize(s));
    }
    // fused nodal update: the first element at every joint
    if (h->sh_own.alloc(SH)) return CB_ERR_CUDA;
    k_own_bits<<<grid_of(SH), 256, 0, s>>>(SH, h->sh_nodes.p, h->node_cstart.p, h->corners.p, h->sh_own.p);
    // ---- tiles ----------------------------------------------------------------------------------------------
    PkIn in{};
    in.pairs = pairs.p; in.contribs = (const CbContrib *)kv_out.p; in.jpair = jpair.p; in.nfree = d_nfree.p; in.colh = colh.p;
    in.base = base.p; in.ax_base = h->ax_base; in.ax_pad = h->ax_pad; in.cls = h->cls_on ? h->sh_class.p : nullptr; in.shp = shp;
    const long nseg = (h->j1 - h->j0 + CB_PK_SEG - 1) / CB_PK_SEG;
    Tmp<long> d_cnt, d_t0;
    if (d_cnt.alloc(nseg) || d_t0.alloc(nseg + 1)) return CB_ERR_CUDA;
    size_t stack0 = 0;
    cudaDeviceGetLimit(&stack0, cudaLimitStackSize);
    DP_TRY(cudaDeviceSetLimit(cudaLimitStackSize, 16 * 1024));     // the packer's scratch lives on the thread stack
    k_pack_count<<<grid_of(nseg, 32), 32, 0, s>>>(in, h->j0, h->j1, nseg, d_cnt.p);
    std::vector<long> cnt(nseg), t0(nseg + 1, 0);
    DP_TRY(cudaMemcpyAsync(cnt.data(), d_cnt.p, nseg * sizeof(long), cudaMemcpyDeviceToHost, s));
    DP_TRY(cudaStreamSynchronize(s));
    for (long k = 0; k < nseg; ++k) { if (cnt[k] < 0) return -1; t0[k + 1] = t0[k] + cnt[k]; }
    const long nt = t0[nseg];
    if (nt <= 0) return -1;
    Plan &P = h->plan_csc;
    if (P.tilesS.alloc(nt) || P.stepsS.alloc((size_t)nt * shp.steps * 32) || P.pairsS.alloc((size_t)nt * shp.pairs) ||
        P.elemsS.alloc((size_t)nt * shp.slots))
        return CB_ERR_CUDA;
    DP_TRY(cudaMemcpyAsync(d_t0.p, t0.data(), (nseg + 1) * sizeof(long), cudaMemcpyHostToDevice, s));
    PkOut out{P.tilesS.p, P.stepsS.p, P.pairsS.p, P.elemsS.p, nullptr};
    k_pack_emit<<<grid_of(nseg, 32), 32, 0, s>>>(in, h->j0, h->j1, nseg, d_t0.p, out);
    DP_TRY(cudaGetLastError());
    DP_TRY(cudaStreamSynchronize(s));
    cudaDeviceSetLimit(cudaLimitStackSize, stack0);
    P.ntilesS = nt; P.nrowsS = nt * shp.steps; P.shapeS = shape_id; P.shape = shp;
    // ---- what the host keeps: adjacency (block keys), touched ranges, bookkeeping -------------------------------
    {
        h->adj_start.resize(NJ + 1); h->adj.resize(np);
        DP_TRY(cudaMemcpyAsync(h->adj.data(), h->d_adj.p, (size_t)np * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        DP_TRY(cudaMemcpyAsync(h->adj_start.data(), h->d_jpair.p, (NJ + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        DP_TRY(cudaStreamSynchronize(s));
    }
    {
        std::vector<uint8_t> touched(NJ, 0);
        const std::vector<int32_t> &nd = h->h_nodes[CB_T_SHELL];
        for (long e = 0; e < SH; ++e) for (int a = 0; a < 3; ++a) touched[nd[e * 4 + a]] = 1;
        h->jl0 = NJ; h->jl1 = 0; h->ql0 = h->sz.NEQ; h->ql1 = 0;
        for (long j = 0; j < NJ; ++j)
            if (touched[j]) {
                if (j < h->jl0) h->jl0 = j;
                h->jl1 = j + 1;
                if (h->h_nfree[j]) {
                    h->ql0 = std::min<long>(h->ql0, h->h_first[j] - 1);
                    h->ql1 = std::max<long>(h->ql1, h->h_first[j] - 1 + h->h_nfree[j]);
                }
            }
        if (h->jl1 <= h->jl0) h->jl0 = h->jl1 = 0;
        if (h->ql1 <= h->ql0) h->ql0 = h->ql1 = 0;
        const bool whole = h->j0 == 0 && h->j1 == NJ;
        bool all_touched = true;
        for (long j = whole ? 0 : h->jl0; j < (whole ? NJ : h->jl1); ++j) if (!touched[j]) all_touched = false;
        h->fuse_node = h->fl.ANAFLAG == 2 && all_touched && !getenv("CB_NO_FUSED_NODE_UPDATE");
    }
    h->ncontrib = nvalid; h->max_dof = 6; h->mixed = 0;
    h->map_bytes = (long)(nt * sizeof(CbTileS) + ((size_t)nt * shp.steps * 32 + (size_t)nt * shp.pairs + (size_t)nt * shp.slots) * 4);
    DP_TRY(cudaStreamSynchronize(s));
    h->plan_on_device = true;
    h->plan_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    h->plan_ready = true;
    return CB_OK;
#undef DP_TRY
}
