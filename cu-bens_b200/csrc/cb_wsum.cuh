// cb_wsum.cuh - warp-level partial sums of the shell force pass (included by cb_api.cu).
//
// forces_sh scatters every element's 18 global force components into f_temp (shell.c:2388-2397).  The device
// gathers instead (one thread per joint, fixed order), which needs the element forces staged in HBM: 144 bytes
// per shell written by the force kernel and read again by the gather.  Most of those bytes are redundant: the 32
// consecutive shells of a warp touch ~34 distinct joints with their 96 corners.  Here every warp sums its
// corners per joint before staging ("slot" = one joint as seen by one warp, at most six corners; a joint with
// more corners in one warp gets a second slot), writes 48 bytes per slot, and the gather adds a joint's slots in
// warp order - 2.8x fewer staged bytes on a structured plate.  The order of the additions is fixed by the mesh
// (slots by warp, corners by element within a slot): bit-reproducible run to run; it is not the element-by-
// element association of the reference (differences at the 1e-16 level of the sum).
//
// Built on the device at the first force pass: one thread per warp sorts its <= 96 (joint, corner) keys, a scan
// numbers the slots, a radix sort by joint turns the slot list into the per-joint lists of the gather.
namespace wsum {

// (joint << 7 | lane * 3 + a) of the corners of warp w, sorted; returns their number
__device__ inline int sorted_keys(long NE, const int32_t *__restrict__ nodes, long w, unsigned long long *keys)
{
    int n = 0;
    for (int l = 0; l < 32; ++l) {
        const long e = w * 32 + l;
        if (e >= NE) break;
        for (int a = 0; a < 3; ++a) {
            const unsigned long long k = ((unsigned long long)(uint32_t)nodes[e * 4 + a] << 7) | (unsigned)(l * 3 + a);
            int i = n++;
            while (i > 0 && keys[i - 1] > k) { keys[i] = keys[i - 1]; --i; }
            keys[i] = k;
        }
    }
    return n;
}

__global__ void k_count(long NE, long nwarp, const int32_t *__restrict__ nodes, int32_t *__restrict__ cnt)
{
    const long w = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (w > nwarp) return;
    if (w == nwarp) { cnt[w] = 0; return; }
    unsigned long long keys[96];
    const int n = sorted_keys(NE, nodes, w, keys);
    int slots = 0, run = 0;
    for (int i = 0; i < n; ++i) {
        const bool same = i > 0 && (keys[i] >> 7) == (keys[i - 1] >> 7) && run < 6;
        if (!same) { ++slots; run = 0; }
        ++run;
    }
    cnt[w] = slots;
}

// slot record: bits 0-2 number of corners, then 7 bits per corner (lane * 3 + a), corners in element order
__global__ void k_emit(long NE, long nwarp, const int32_t *__restrict__ nodes, const int32_t *__restrict__ start,
                       unsigned long long *__restrict__ corners, uint32_t *__restrict__ slot_joint, int32_t *__restrict__ slot_id)
{
    const long w = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (w >= nwarp) return;
    unsigned long long keys[96];
    const int n = sorted_keys(NE, nodes, w, keys);
    int s = start[w] - 1, run = 0;
    unsigned long long rec = 0;
    for (int i = 0; i < n; ++i) {
        const bool same = i > 0 && (keys[i] >> 7) == (keys[i - 1] >> 7) && run < 6;
        if (!same) {
            if (s >= start[w]) corners[s] = rec | (unsigned)run;
            ++s; run = 0; rec = 0;
            slot_joint[s] = (uint32_t)(keys[i] >> 7); slot_id[s] = s;
        }
        rec |= (keys[i] & 127ull) << (3 + 7 * run);
        ++run;
    }
    if (n > 0) corners[s] = rec | (unsigned)run;
}
}   // namespace wsum

static int build_wsum(cb_handle *h)
{
    h->ws_tried = true;
    if (!h->fuse_node || getenv("CB_NO_WARP_SUMS") || g_host_only) return CB_OK;
    const long SH = h->sz.NE_SH, NJ = h->sz.NJ, nwarp = (SH + 31) / 32;
    if (SH <= 0 || nwarp >= (1L << 30)) return CB_OK;
    cudaStream_t s = h->stream;
#define WS_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(CB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)
    DevBuf<int32_t> cnt, slot_id; DevBuf<uint32_t> sj_in, sj_out; DevBuf<unsigned char> tmp;
    struct Rel { DevBuf<int32_t> *a, *b; DevBuf<uint32_t> *c, *d; DevBuf<unsigned char> *e;
                 ~Rel() { a->release(); b->release(); c->release(); d->release(); e->release(); } } rel{&cnt, &slot_id, &sj_in, &sj_out, &tmp};
    if (cnt.alloc((size_t)nwarp + 1) || h->ws_start.alloc((size_t)nwarp + 1)) return CB_ERR_CUDA;
    const unsigned gw = (unsigned)((nwarp + 1 + 127) / 128);
    wsum::k_count<<<gw, 128, 0, s>>>(SH, nwarp, h->sh_nodes.p, cnt.p);
    size_t tb = 0;
    WS_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, h->ws_start.p, (int)(nwarp + 1), s));
    if (tmp.alloc(tb)) return CB_ERR_CUDA;
    WS_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, h->ws_start.p, (int)(nwarp + 1), s));
    int32_t nslot = 0;
    WS_TRY(cudaMemcpyAsync(&nslot, h->ws_start.p + nwarp, sizeof nslot, cudaMemcpyDeviceToHost, s));
    WS_TRY(cudaStreamSynchronize(s));
    if (nslot <= 0) return CB_OK;
    if (h->ws_corners.alloc((size_t)nslot) || h->ws_fg.alloc((size_t)nslot * 6) || h->js_slots.alloc((size_t)nslot) ||
        h->js_start.alloc((size_t)NJ + 1) || slot_id.alloc((size_t)nslot) || sj_in.alloc((size_t)nslot) || sj_out.alloc((size_t)nslot))
        return CB_ERR_CUDA;
    wsum::k_emit<<<gw, 128, 0, s>>>(SH, nwarp, h->sh_nodes.p, h->ws_start.p, h->ws_corners.p, sj_in.p, slot_id.p);
    int jbits = 1; while ((1L << jbits) < NJ) ++jbits;
    tmp.release(); tb = 0;
    WS_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, sj_in.p, sj_out.p, slot_id.p, h->js_slots.p, (int)nslot, 0, jbits, s));
    if (tmp.alloc(tb)) return CB_ERR_CUDA;
    WS_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, tb, sj_in.p, sj_out.p, slot_id.p, h->js_slots.p, (int)nslot, 0, jbits, s));
    devplan::k_lower_bound<uint32_t><<<(unsigned)((NJ + 1 + 255) / 256), 256, 0, s>>>(NJ, sj_out.p, nslot, 0, h->js_start.p);
    WS_TRY(cudaGetLastError());
    WS_TRY(cudaStreamSynchronize(s));
    h->ws_nslot = nslot; h->ws_nwarp = nwarp;
    h->ws_ready = true;
    h->map_bytes += (size_t)(nwarp + 1) * 4 + (size_t)nslot * 12 + (size_t)(NJ + 1) * 4;
#undef WS_TRY
    return CB_OK;
}
