// cb_sym.cuh - symmetric hand-off of the assembled tangent matrix to the host-side solver (included by
// cb_api.cu; SURVEY 8(e) "To the solver": the device -> host copy of Ax, not the assembly, is what an
// iteration waits for).
//
// K_t is symmetric, umfpack_di_* wants the full (unsymmetric-storage) CSC (solve.c:110-135).  Instead of
// sending all nnz values over PCIe, the device packs the UPPER triangle of the owned column slice - rows
// are ascending inside a column, so that is a prefix of every column (plus, for an element-partitioned
// slice, the rows of neighbour joints beyond the owned range, a suffix, whose mirror images live on another
// rank) - into one contiguous stream, ships it in chunks, and host threads rebuild the full columns while
// the next chunk is still on the wire: the upper part of a column is copied, the lower part is the
// transpose of blocks found in the packed columns of the higher-numbered neighbour joints.
// The packed stream with cb_csc_upper_pattern is itself a standard upper-triangular CSC for symmetric
// solvers (the built-in LDL^T of host/cb_sparse.c needs nothing else).
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

struct SymPlan {
    bool ready = false;
    long nnzu = 0;                          // packed length
    std::vector<int32_t> ulen0;             // [NJ] rows above the diagonal block of joint j's columns
    std::vector<int32_t> slen;              // [NJ] rows of neighbour joints >= j1 (suffix), 0 for a whole model
    std::vector<int64_t> ubase;             // [NJ+1] packed offset of joint j's first column
    std::vector<int32_t> mirror;            // per adjacency entry (A, B > A, B owned): row offset of A inside B's columns
    DevBuf<int32_t> d_ulen0, d_slen, d_nfree, d_colh;
    DevBuf<int64_t> d_ubase, d_base;
    DevBuf<double> packed;                  // device staging of the packed stream
    cudaStream_t copy_stream = nullptr;
    // The stream is cut into chunks of consecutive joints.  Most cross PCIe as packed upper triangles and are
    // rebuilt by host threads; every `full_every`-th chunk crosses as its full columns straight into place
    // (no host work), which balances the copy engine against the host's memory bandwidth: with 16 threads the
    // rebuild of a chunk takes about 1.5 x as long as its packed copy.
    static const int NCHUNK = 27;
    int full_every = 0;                     // 0: every chunk packed
    bool chunk_full[NCHUNK] = {};
    std::vector<uint8_t> joint_full;        // [NJ] the joint's chunk crosses as full columns
    cudaEvent_t ev_pack = nullptr, ev_chunk[NCHUNK] = {};
    long chunk_j[NCHUNK + 1] = {};          // joint ranges of the chunks
    // asynchronous rebuild
    std::thread worker;
    bool busy = false;
    int rc = 0;
};

__global__ void __launch_bounds__(256)
k_pack_upper(long j0, long j1, const int32_t *__restrict__ nfree, const int32_t *__restrict__ colh,
             const int32_t *__restrict__ ulen0, const int32_t *__restrict__ slen, const int64_t *__restrict__ base,
             const int64_t *__restrict__ ubase, int64_t ax_base, const double *__restrict__ Ax, double *__restrict__ out)
{
    // one warp per joint: its columns' prefix (rows <= column) and suffix (rows of joints past the owned range)
    const long j = j0 + (blockIdx.x * (long)blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    if (j >= j1) return;
    const int nf = nfree[j], ch = colh[j], u0 = ulen0[j], sl = slen[j];
    const double *src = Ax + (base[j] - ax_base);
    double *dst = out + ubase[j];
    for (int cc = 0; cc < nf; ++cc) {
        const int len = u0 + cc + 1;
        for (int k = lane; k < len; k += 32) dst[k] = src[k];
        for (int k = lane; k < sl; k += 32) dst[len + k] = src[ch - sl + k];
        dst += len + sl; src += ch;
    }
}
