// cb_sym_impl.cuh - see cb_sym.cuh.  Included at the end of cb_api.cu.

static inline long sym_poff(int cc, int u0, int sl) { return (long)cc * (u0 + sl) + (long)cc * (cc + 1) / 2; }

// host-side structure of the packed upper triangle (once per model)
static int sym_prepare(cb_handle *h)
{
    SymPlan &S = h->sym;
    if (S.ready) return CB_OK;
    int rc = build_plan(h); if (rc) return rc;
    if (!(h->layout & CB_MAT_CSC)) return fail(CB_ERR_ARG, "handle has no CSC layout");
    const long NJ = h->sz.NJ;
    for (long j = h->j0; j < h->j1; ++j)
        if (h->colh[j] > 255) return fail(CB_ERR_UNSUPPORTED, "symmetric hand-off: a joint couples to more than 255 equations");
    S.ulen0.assign(NJ, 0); S.slen.assign(NJ, 0); S.ubase.assign(NJ + 1, 0);
    S.mirror.assign(h->adj.size(), 0);
    long pos = 0;
    for (long j = 0; j < NJ; ++j) {
        S.ubase[j] = pos;
        if (j < h->j0 || j >= h->j1 || !h->h_nfree[j]) continue;
        int u0 = 0, sl = 0;
        for (int k = h->adj_start[j]; k < h->adj_start[j + 1]; ++k) {
            const int32_t A = h->adj[k];
            if (A < j) u0 += h->h_nfree[A];
            else if (A >= h->j1) sl += h->h_nfree[A];
        }
        S.ulen0[j] = u0; S.slen[j] = sl;
        const int nf = h->h_nfree[j];
        pos += (long)nf * (u0 + sl) + (long)nf * (nf + 1) / 2;
    }
    S.ubase[NJ] = pos; S.nnzu = pos;
    // where the rows of joint A start inside the columns of a higher-numbered owned neighbour B
    for (long A = h->j0; A < h->j1; ++A)
        for (int k = h->adj_start[A]; k < h->adj_start[A + 1]; ++k) {
            const int32_t B = h->adj[k];
            if (B <= A || B >= h->j1) continue;
            int r = 0;
            for (int q = h->adj_start[B]; q < h->adj_start[B + 1] && h->adj[q] < A; ++q) r += h->h_nfree[h->adj[q]];
            S.mirror[k] = r;
        }
    // chunks of (roughly) equal packed size
    for (int c = 0; c <= SymPlan::NCHUNK; ++c) {
        const long target = h->j0 == h->j1 ? 0 : (long)((double)S.nnzu * c / SymPlan::NCHUNK);
        long j = std::lower_bound(S.ubase.begin() + h->j0, S.ubase.begin() + h->j1, target) - S.ubase.begin();
        S.chunk_j[c] = (c == SymPlan::NCHUNK) ? h->j1 : std::min(std::max(j, h->j0), h->j1);
    }
    {   // which chunks cross as full columns (CB_SYM_FULL_EVERY = k: every k-th; 0: none; default 5)
        const char *e = getenv("CB_SYM_FULL_EVERY");
        S.full_every = e ? atoi(e) : 5;
        S.joint_full.assign(NJ, 0);
        for (int c = 0; c < SymPlan::NCHUNK; ++c) {
            S.chunk_full[c] = S.full_every > 0 && (c % S.full_every) == S.full_every - 1;
            if (S.chunk_full[c]) for (long j = S.chunk_j[c]; j < S.chunk_j[c + 1]; ++j) S.joint_full[j] = 1;
        }
    }
    if (!g_host_only) {
        std::vector<int64_t> b64(h->base.begin(), h->base.end());
        if (S.d_ulen0.upload(S.ulen0) || S.d_slen.upload(S.slen) || S.d_nfree.upload(h->h_nfree) || S.d_colh.upload(h->colh) ||
            S.d_ubase.upload(S.ubase) || S.d_base.upload(b64) || S.packed.alloc((size_t)S.nnzu + 1))
            return CB_ERR_CUDA;
        if (cudaStreamCreateWithFlags(&S.copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&S.ev_pack, cudaEventDisableTiming) != cudaSuccess)
            return fail(CB_ERR_CUDA, "copy stream creation failed");
        for (int c = 0; c < SymPlan::NCHUNK; ++c)
            if (cudaEventCreateWithFlags(&S.ev_chunk[c], cudaEventDisableTiming) != cudaSuccess)
                return fail(CB_ERR_CUDA, "event creation failed");
        CUDA_TRY(cudaDeviceSynchronize());
    }
    S.ready = true;
    return CB_OK;
}

// full columns of the joints [ja, jb) from the packed stream P (whole stream addressable): the upper part
// of a column is a copy, the rest the mirror image of upper blocks.  A joint's columns are one contiguous
// run of Ax: they are put together in a small cache-resident buffer and leave with streaming stores (the
// destination is written exactly once and not read back by this thread - no write-allocate traffic).
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
static inline void sym_stream_out(double *dst, const double *src, long n)
{
#if defined(__SSE2__)
    long i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15)) { dst[i] = src[i]; ++i; }
    for (; i + 2 <= n; i += 2) _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
    for (; i < n; ++i) dst[i] = src[i];
#else
    memcpy(dst, src, (size_t)n * sizeof(double));
#endif
}

static void sym_expand_range(const cb_handle *h, const double *P, double *Ax, long ja, long jb)
{
    const SymPlan &S = h->sym;
    double buf[7 * 256];                           // 7 columns of at most 255 rows (colh is checked in sym_prepare)
    const int32_t *adj = h->adj.data(), *adj_start = h->adj_start.data(), *nfree = h->h_nfree.data();
    const uint8_t *jfull = S.joint_full.data();
    for (long A = ja; A < jb; ++A) {
        const int nA = nfree[A];
        if (!nA || jfull[A]) continue;             // a joint of a full chunk is already in place
        const int u0 = S.ulen0[A], sl = S.slen[A], hA = h->colh[A];
        const double *PA = P + S.ubase[A];
        long po[8];
        for (int cc = 0; cc <= nA; ++cc) po[cc] = sym_poff(cc, u0, sl);
        for (int cc = 0; cc < nA; ++cc) {
            double *F = buf + cc * hA;
            const double *pc = PA + po[cc];
            for (int k = 0; k < u0 + cc + 1; ++k) F[k] = pc[k];                          // rows <= column
            for (int rr = cc + 1; rr < nA; ++rr) F[u0 + rr] = PA[po[rr] + u0 + cc];       // diagonal block
        }
        int ro = u0 + nA, so = 0;
        for (int k = adj_start[A]; k < adj_start[A + 1]; ++k) {
            const int32_t B = adj[k];
            const int nB = nfree[B];
            if (B <= A || !nB) continue;
            if (B < h->j1 && jfull[B]) {           // block (A, B) sits in the full columns of joint B, already in Ax
                const int hB = h->colh[B];
                const double *FB = Ax + (h->base[B] - h->ax_base) + S.mirror[k];
                for (int c2 = 0; c2 < nB; ++c2)
                    for (int cc = 0; cc < nA; ++cc) buf[cc * hA + ro + c2] = FB[(long)c2 * hB + cc];
            } else if (B < h->j1) {                // ... or in the packed columns of joint B
                const int uB = S.ulen0[B], sB = S.slen[B];
                const double *PB = P + S.ubase[B] + S.mirror[k];
                if (nA == 6 && nB == 6) {
                    long off = 0;
#pragma unroll
                    for (int c2 = 0; c2 < 6; ++c2) {
                        const double *src = PB + off;                                    // rows of A in column c2 of B
                        double *d = buf + ro + c2;
                        d[0] = src[0]; d[hA] = src[1]; d[2 * hA] = src[2]; d[3 * hA] = src[3]; d[4 * hA] = src[4]; d[5 * hA] = src[5];
                        off += uB + sB + c2 + 1;
                    }
                } else {
                    for (int c2 = 0; c2 < nB; ++c2) {
                        const double *src = PB + sym_poff(c2, uB, sB);
                        for (int cc = 0; cc < nA; ++cc) buf[cc * hA + ro + c2] = src[cc];
                    }
                }
            } else {                               // beyond the owned range: shipped as a suffix of A's own columns
                for (int cc = 0; cc < nA; ++cc)
                    for (int c2 = 0; c2 < nB; ++c2) buf[cc * hA + ro + c2] = PA[po[cc] + u0 + cc + 1 + so + c2];
                so += nB;
            }
            ro += nB;
        }
        sym_stream_out(Ax + (h->base[A] - h->ax_base), buf, (long)nA * hA);
    }
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

// worker pool of one rebuild: thread t expands its share of every chunk as soon as the chunk's sources
// (the chunk itself and, for the mirrored blocks, the next ones up to the matrix bandwidth) have arrived
namespace {
struct SymJob {
    cb_handle *h; const double *P; double *Ax; int nthreads;
    std::atomic<int> arrived{0};              // chunks on the host so far
    std::mutex mu; std::condition_variable cv;
    int need[SymPlan::NCHUNK];                // chunks that must have arrived before chunk c can be expanded
};
void sym_worker(SymJob *J, int t)
{
    const SymPlan &S = J->h->sym;
    for (int c = 0; c < SymPlan::NCHUNK; ++c) {
        {
            std::unique_lock<std::mutex> lk(J->mu);
            J->cv.wait(lk, [&] { return J->arrived.load() >= J->need[c]; });
        }
        const long a = S.chunk_j[c], b = S.chunk_j[c + 1], n = b - a;
        sym_expand_range(J->h, J->P, J->Ax, a + n * t / J->nthreads, a + n * (t + 1) / J->nthreads);
    }
}
void sym_needs(const cb_handle *h, int *need)
{
    // chunk c's joints reach forward to their largest owned neighbour
    const SymPlan &S = h->sym;
    for (int c = 0; c < SymPlan::NCHUNK; ++c) {
        long far = S.chunk_j[c + 1];
        for (long A = S.chunk_j[c]; A < S.chunk_j[c + 1]; ++A)
            if (h->adj_start[A + 1] > h->adj_start[A]) {
                const long B = h->adj[h->adj_start[A + 1] - 1];
                if (B < h->j1 && B + 1 > far) far = B + 1;
            }
        int k = c + 1;
        while (k < SymPlan::NCHUNK && S.chunk_j[k] < far) ++k;
        need[c] = k;
    }
}
}

extern "C" long cb_csc_upper_nnz(cb_handle *h)
{
    if (!h) return -1;
    if (!g_host_only) cudaSetDevice(h->fl.device);
    if (sym_prepare(h)) return -1;
    return h->sym.nnzu;
}

// upper-triangular CSC pattern of the owned column slice (plus, for a partition, the rows past the owned
// range): Apu[NEQ+1], Aiu[cb_csc_upper_nnz]
extern "C" int cb_csc_upper_pattern(cb_handle *h, int *Apu, int *Aiu)
{
    if (!h || !Apu) return fail(CB_ERR_ARG, "null argument");
    if (!g_host_only) cudaSetDevice(h->fl.device);
    int rc = sym_prepare(h); if (rc) return rc;
    const SymPlan &S = h->sym;
    if (S.nnzu > 0x7fffffffL) return fail(CB_ERR_OVERFLOW, "upper nnz exceeds 32-bit indices");
    const long NJ = h->sz.NJ;
    for (long j = 0; j < NJ; ++j) {
        const bool own = j >= h->j0 && j < h->j1;
        const int nf = h->h_nfree[j], u0 = S.ulen0[j], sl = S.slen[j];
        for (int cc = 0; cc < nf; ++cc) {
            long p = own ? S.ubase[j] + sym_poff(cc, u0, sl) : (j < h->j0 ? 0 : S.nnzu);
            Apu[h->h_first[j] - 1 + cc] = (int)p;
            if (!Aiu || !own) continue;
            for (int k = h->adj_start[j]; k < h->adj_start[j + 1]; ++k) {
                const int32_t A = h->adj[k];
                if (A < j) for (int rr = 0; rr < h->h_nfree[A]; ++rr) Aiu[p++] = h->h_first[A] - 1 + rr;
            }
            for (int rr = 0; rr <= cc; ++rr) Aiu[p++] = h->h_first[j] - 1 + rr;
            for (int k = h->adj_start[j]; k < h->adj_start[j + 1]; ++k) {
                const int32_t A = h->adj[k];
                if (A >= h->j1) for (int rr = 0; rr < h->h_nfree[A]; ++rr) Aiu[p++] = h->h_first[A] - 1 + rr;
            }
        }
    }
    Apu[h->sz.NEQ] = (int)S.nnzu;
    return CB_OK;
}

// pack on the device and ship the chunks (asynchronous: events mark the arrival of each chunk)
static int sym_launch_transfer(cb_handle *h, double *Axu_host, double *Ax_host)
{
    SymPlan &S = h->sym;
    const long nj = h->j1 - h->j0;
    if (nj > 0) {
        const long warps = nj, threads = warps * 32;
        k_pack_upper<<<(unsigned)((threads + 255) / 256), 256, 0, h->stream>>>(
            h->j0, h->j1, S.d_nfree.p, S.d_colh.p, S.d_ulen0.p, S.d_slen.p, S.d_base.p, S.d_ubase.p, h->ax_base,
            h->Ax.p + h->ax_pad, S.packed.p);
        ++h->launches;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(S.ev_pack, h->stream));
    CUDA_TRY(cudaStreamWaitEvent(S.copy_stream, S.ev_pack, 0));
    for (int c = 0; c < SymPlan::NCHUNK; ++c) {
        if (Ax_host && S.chunk_full[c]) {          // full columns straight into place
            const long a = h->base[S.chunk_j[c]] - h->ax_base, b = h->base[S.chunk_j[c + 1]] - h->ax_base;
            if (b > a)
                CUDA_TRY(cudaMemcpyAsync(Ax_host + a, h->Ax.p + h->ax_pad + a, (size_t)(b - a) * sizeof(double),
                                         cudaMemcpyDeviceToHost, S.copy_stream));
        } else {
            const long a = S.ubase[S.chunk_j[c]], b = S.ubase[S.chunk_j[c + 1]];
            if (b > a)
                CUDA_TRY(cudaMemcpyAsync(Axu_host + a, S.packed.p + a, (size_t)(b - a) * sizeof(double), cudaMemcpyDeviceToHost,
                                         S.copy_stream));
        }
        CUDA_TRY(cudaEventRecord(S.ev_chunk[c], S.copy_stream));
    }
    return CB_OK;
}

extern "C" int cb_get_csc_upper_values(cb_handle *h, double *Axu)
{
    if (!h || !Axu) return fail(CB_ERR_ARG, "null argument");
    if (!(h->layout & CB_MAT_CSC) || !h->Ax.p) return fail(CB_ERR_ARG, "no CSC matrix assembled");
    cudaSetDevice(h->fl.device);
    int rc = sym_prepare(h); if (rc) return rc;
    if (h->sym.busy) return fail(CB_ERR_ARG, "a mirrored transfer is in flight (cb_csc_values_end)");
    rc = sym_launch_transfer(h, Axu, nullptr); if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->sym.copy_stream));
    return CB_OK;
}

// full CSC values on the host from half the PCIe traffic.  Ax [cb_csc_nnz] receives the full columns,
// Axu_staging [cb_csc_upper_nnz] (page-locked: cb_host_alloc) the packed upper triangle.  _begin returns once
// the work is queued (pack kernel, chunked copies on a second stream, nthreads host threads rebuilding
// the columns as chunks arrive) - the caller may go on with cb_update_forces; _end waits for the matrix.
extern "C" int cb_csc_values_begin(cb_handle *h, double *Ax, double *Axu_staging, int nthreads)
{
    if (!h || !Ax || !Axu_staging) return fail(CB_ERR_ARG, "null argument");
    if (!(h->layout & CB_MAT_CSC) || !h->Ax.p) return fail(CB_ERR_ARG, "no CSC matrix assembled");
    cudaSetDevice(h->fl.device);
    int rc = sym_prepare(h); if (rc) return rc;
    SymPlan &S = h->sym;
    if (S.busy) return fail(CB_ERR_ARG, "cb_csc_values_begin: previous transfer not ended");
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    rc = sym_launch_transfer(h, Axu_staging, Ax); if (rc) return rc;
    S.busy = true; S.rc = 0;
    const int dev = h->fl.device;
    S.worker = std::thread([h, Ax, Axu_staging, nthreads, dev]() {
        SymPlan &S = h->sym;
        cudaSetDevice(dev);
        SymJob J; J.h = h; J.P = Axu_staging; J.Ax = Ax; J.nthreads = nthreads;
        sym_needs(h, J.need);
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(sym_worker, &J, t);
        for (int c = 0; c < SymPlan::NCHUNK; ++c) {
            if (cudaEventSynchronize(S.ev_chunk[c]) != cudaSuccess) S.rc = 1;
            { std::lock_guard<std::mutex> lk(J.mu); J.arrived.store(c + 1); }
            J.cv.notify_all();
        }
        for (auto &th : pool) th.join();
    });
    return CB_OK;
}

// bytes one cb_csc_values_begin moves device -> host (full chunks + packed chunks)
extern "C" long cb_csc_values_d2h_bytes(cb_handle *h)
{
    if (!h || sym_prepare(h)) return -1;
    const SymPlan &S = h->sym;
    long n = 0;
    for (int c = 0; c < SymPlan::NCHUNK; ++c)
        n += S.chunk_full[c] ? (h->base[S.chunk_j[c + 1]] - h->base[S.chunk_j[c]]) : (S.ubase[S.chunk_j[c + 1]] - S.ubase[S.chunk_j[c]]);
    return n * (long)sizeof(double);
}

extern "C" int cb_csc_values_end(cb_handle *h)
{
    if (!h) return fail(CB_ERR_ARG, "null handle");
    SymPlan &S = h->sym;
    if (!S.busy) return CB_OK;
    S.worker.join();
    S.busy = false;
    return S.rc ? fail(CB_ERR_CUDA, "mirrored transfer failed") : CB_OK;
}

extern "C" int cb_get_csc_values_mirrored(cb_handle *h, double *Ax, double *Axu_staging, int nthreads)
{
    const int rc = cb_csc_values_begin(h, Ax, Axu_staging, nthreads);
    return rc ? rc : cb_csc_values_end(h);
}

// Host-only self test of the packed layout and the rebuild (no device): a synthetic symmetric matrix on
// the model's CSC pattern, packed by the same rule as k_pack_upper, expanded with nthreads threads and
// compared entry by entry.  seconds (may be NULL) receives the wall time of the rebuild alone.
extern "C" int cb_sym_selftest(const cb_sizes *sz, const cb_flags *fl, const cb_model *m, long j0, long j1,
                               int nthreads, double *seconds)
{
    g_host_only = true;
    cb_handle *h = nullptr;
    int rc = cb_create(sz, fl, m, &h);
    if (rc == CB_OK && (j0 != 0 || j1 != 0)) rc = cb_set_owned_joints(h, j0, j1);
    if (rc == CB_OK) rc = sym_prepare(h);
    if (rc == CB_OK) {
        const SymPlan &S = h->sym;
        std::vector<int> Ap(h->sz.NEQ + 1), Ai((size_t)h->nnz);
        host_pattern(h, Ap.data(), Ai.data());
        std::vector<double> full((size_t)h->nnz), packed((size_t)S.nnzu), out((size_t)h->nnz, -1.0);
        auto val = [](long i, long j) {        // symmetric in (i, j)
            const long a = std::min(i, j), b = std::max(i, j);
            return (double)((a * 1315423911L + b * 2654435761L) % 1000003L) + 0.5;
        };
        for (long c = 0; c < h->sz.NEQ; ++c)
            for (int p = Ap[c]; p < Ap[c + 1]; ++p) full[p] = val(Ai[p], c);
        for (long j = h->j0; j < h->j1; ++j) {       // the rule of k_pack_upper
            const int nf = h->h_nfree[j], ch = h->colh[j], u0 = S.ulen0[j], sl = S.slen[j];
            const double *src = full.data() + (h->base[j] - h->ax_base);
            double *dst = packed.data() + S.ubase[j];
            for (int cc = 0; cc < nf; ++cc) {
                const int len = u0 + cc + 1;
                for (int k = 0; k < len; ++k) dst[k] = src[k];
                for (int k = 0; k < sl; ++k) dst[len + k] = src[ch - sl + k];
                dst += len + sl; src += ch;
            }
        }
        // the upper pattern must describe exactly the packed stream
        std::vector<int> Apu(h->sz.NEQ + 1), Aiu((size_t)S.nnzu);
        rc = cb_csc_upper_pattern(h, Apu.data(), Aiu.data());
        for (long c = 0; rc == CB_OK && c < h->sz.NEQ; ++c)
            for (int p = Apu[c]; p < Apu[c + 1]; ++p)
                if (packed[p] != val(Aiu[p], c)) { rc = fail(CB_ERR_ARG, "upper pattern does not match the packed stream at column %ld", c); break; }
        if (rc == CB_OK) {
            // what the copy engine does: full chunks land in place, the packed stream only holds the others
            for (int c = 0; c < SymPlan::NCHUNK; ++c)
                if (S.chunk_full[c]) {
                    const long a = h->base[S.chunk_j[c]] - h->ax_base, b = h->base[S.chunk_j[c + 1]] - h->ax_base;
                    std::copy(full.begin() + a, full.begin() + b, out.begin() + a);
                    std::fill(packed.begin() + S.ubase[S.chunk_j[c]], packed.begin() + S.ubase[S.chunk_j[c + 1]], -7.0);
                }
            SymJob J; J.h = h; J.P = packed.data(); J.Ax = out.data(); J.nthreads = nthreads < 1 ? 1 : nthreads;
            sym_needs(h, J.need);
            J.arrived.store(SymPlan::NCHUNK);
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> pool;
            for (int t = 0; t < J.nthreads; ++t) pool.emplace_back(sym_worker, &J, t);
            for (auto &th : pool) th.join();
            if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            for (long p = 0; p < h->nnz; ++p)
                if (out[p] != full[p]) { rc = fail(CB_ERR_ARG, "rebuilt matrix differs at entry %ld", p); break; }
            // every chunk's sources must be covered by its `need`
            for (int c = 0; c < SymPlan::NCHUNK; ++c)
                if (J.need[c] < c + 1 || J.need[c] > SymPlan::NCHUNK) rc = fail(CB_ERR_ARG, "bad chunk dependency");
        }
    }
    if (h) cb_destroy(h);
    g_host_only = false;
    return rc;
}
