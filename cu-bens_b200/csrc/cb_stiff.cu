// cb_stiff.cu - tangent stiffness K_t and its assembly, fused.
//
// Replaces `ss = 0; stiff_tr; stiff_fr; stiff_sh; [stiff_br]` (main.c:1899-1921; truss.c:82-204,
// frame.c:226-362, shell.c:110-344, brick.c:79-397) together with the generic dense
// transform() (misc.c:41-69) and the skyline / dense scatter inlined in each of them.
//
// One thread owns one node-pair block of the global matrix (row node A, column node B) and walks
// the block's sorted contribution list (element type, element, local row node a, local column
// node b - the order the reference adds them in).  For each contribution it evaluates only the
// 6x6 (7x7 frame, 3x3 truss/brick) sub-block K_ab = T_a^T k_ab T_b of that element - the
// transformation is block diagonal, so the reference's 2*18^3 dense products collapse to four
// 3x3 triple products - accumulates in registers and finally writes the block once.  No atomics,
// no colouring, no staging of element matrices in HBM: the only HBM write is the matrix itself.
// Threads are bucketed by contribution count on the host so warps do uniform work.
#include "cb_internal.h"

#define CB_TPB_K 128

// ---- R^T S R for the four 3x3 sub-blocks of a shell node-pair block -----------------------
// R rows are the local axes e1,e2,e3 (R[3*r+c]); S is given in local axes.
__device__ __forceinline__ void rtsr_add(const double *R, const double S[3][3], double *acc,
                                         int r0, int c0, int ld)
{
    double W[3][3];                      // W = S R
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            W[r][q] = S[r][0] * R[q] + S[r][1] * R[3 + q] + S[r][2] * R[6 + q];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            acc[(r0 + p) * ld + c0 + q] += R[p] * W[0][q] + R[3 + p] * W[1][q] + R[6 + p] * W[2][q];
}

// shape-function gradients of the CST in the reference local axes: node 0,1,2 -> (bx, by)*2A
__device__ __forceinline__ void cst_grad(int n, double X2, double X3, double Y3, double &bx,
                                         double &by)
{
    bx = (n == 0) ? -Y3 : (n == 1 ? Y3 : 0.0);
    by = (n == 0) ? (X3 - X2) : (n == 1 ? -X3 : X2);
}

// K_ab (6x6, global axes) of shell e, added into acc[6][6] (row-major, ld = 7 for frame mixing)
__device__ __forceinline__ void shell_block(const CbStiffArgs &A, int e, int a, int b, double *acc,
                                            int ld)
{
    const double *sc = A.d.sh_const + (long)e * CB_SH_CONST;
    const double *fr = A.sh_frame + (long)e * CB_SH_FRAME;
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = fr[i];
    const double E = sc[0], nu = sc[1], t = sc[2], A0 = sc[4], X2 = sc[5], X3 = sc[6], Y3 = sc[7];
    const double C00 = E / (1 - nu * nu), C01 = C00 * nu, C22 = C00 * (1 - nu) / 2;
    double bxa, bya, bxb, byb;
    cst_grad(a, X2, X3, Y3, bxa, bya);
    cst_grad(b, X2, X3, Y3, bxb, byb);
    // membrane (shell.c:487-531): t*A0 * Bm_a^T C Bm_b with Bm = grad/(2 A0)
    const double sm = t * A0 / (4 * A0 * A0);
    double m00 = sm * (C00 * bxa * bxb + C22 * bya * byb);
    double m01 = sm * (C01 * bxa * byb + C22 * bya * bxb);
    double m10 = sm * (C01 * bya * bxb + C22 * bxa * byb);
    double m11 = sm * (C00 * bya * byb + C22 * bxa * bxb);
    double g = 0.0;
    if (A.d.ANAFLAG == 2) {
        // geometric (shell.c:660-840): A_def * Bnl^T N Bnl, Bnl = grad/(2 A_def)
        const double Ad = fr[9];
        const double *Nm = A.d.sh_Nm + (long)e * 4;
        g = (bxa * (Nm[0] * bxb + Nm[2] * byb) + bya * (Nm[2] * bxb + Nm[1] * byb)) / (4 * Ad);
    }
    // bending 3x3 sub-block of the precomputed DKT matrix (rows w,tx,ty of a; cols of b)
    const double *kb = A.d.sh_keb + (long)e * 81 + (3 * a) * 9 + 3 * b;
    const double k00 = kb[0], k01 = kb[1], k02 = kb[2];
    const double k10 = kb[9], k11 = kb[10], k12 = kb[11];
    const double k20 = kb[18], k21 = kb[19], k22 = kb[20];
    double drill = 0.0;
    if (a == b) drill = A.d.sh_keb[(long)e * 81 + (3 * a + 1) * 10] / 10000;   // shell.c:482-484
    {   // translation-translation
        const double S[3][3] = {{m00 + g, m01, 0}, {m10, m11 + g, 0}, {0, 0, k00 + g}};
        rtsr_add(R, S, acc, 0, 0, ld);
    }
    {   // translation(row) - rotation(col): w row couples to theta_x, theta_y
        const double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {k01, k02, 0}};
        rtsr_add(R, S, acc, 0, 3, ld);
    }
    {   // rotation(row) - translation(col)
        const double S[3][3] = {{0, 0, k10}, {0, 0, k20}, {0, 0, 0}};
        rtsr_add(R, S, acc, 3, 0, ld);
    }
    {   // rotation - rotation, drilling stiffness on theta_z
        const double S[3][3] = {{k11, k12, 0}, {k21, k22, 0}, {0, 0, drill}};
        rtsr_add(R, S, acc, 3, 3, ld);
    }
}

// K_ab (3x3) of truss e (truss.c:102-166)
__device__ __forceinline__ void truss_block(const CbStiffArgs &A, int e, int a, int b, double *acc,
                                            int ld)
{
    const double *tc = A.d.tr_const + (long)e * CB_TR_CONST;
    const double *fr = A.tr_frame + (long)e * CB_TR_FRAME;
    const double c[3] = {fr[0], fr[1], fr[2]};
    const double dl = fr[3];
    double k, gN = 0.0;
    if (A.d.ANAFLAG == 1) {
        k = tc[0] * tc[1] / tc[2];
    } else {
        k = tc[0] * tc[1] * (dl * dl) / tc[3];
        gN = A.tr_ef[(long)e * 2] / dl;
    }
    const double s = (a == b) ? 1.0 : -1.0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            acc[p * ld + q] += s * (k * c[p] * c[q] + ((p == q) ? gN : 0.0));
}

// ------------------------------------------------------------------------------------------
// the assembly kernel: one thread per node-pair block
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CB_TPB_K)
k_assemble_blocks(CbStiffArgs A)
{
    const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= A.npairs) return;
    const CbPair pr = A.pairs[p];
    double acc[49];
#pragma unroll
    for (int i = 0; i < 49; ++i) acc[i] = 0.0;
    for (int c = 0; c < pr.ccount; ++c) {
        const CbContrib ct = A.contribs[pr.cstart + c];
        if (ct.type == CB_T_SHELL) shell_block(A, ct.e, ct.a, ct.b, acc, 7);
        else if (ct.type == CB_T_TRUSS) truss_block(A, ct.e, ct.a, ct.b, acc, 7);
    }
    if (!A.skyline) {
        // CSC: column cc of node B is contiguous; rows of node A start at pr.off inside it
        int cc = 0;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (!((pr.maskB >> c) & 1)) continue;
            double *col = A.out + (long)pr.off + (long)cc * pr.colh;
            int rr = 0;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                if (!((pr.maskA >> r) & 1)) continue;
                col[rr] = acc[r * 7 + c];
                ++rr;
            }
            ++cc;
        }
    } else {
        // skyline: entry (i <= j) lives at ss[maxa[j-1] + (j-i) - 1] (model.c:1269-1278)
        int cc = 0;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (!((pr.maskB >> c) & 1)) continue;
            const long j = pr.eqB0 + cc;
            const long dj = A.maxa[j - 1] - 1;
            int rr = 0;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                if (!((pr.maskA >> r) & 1)) continue;
                const long i = pr.eqA0 + rr;
                if (i <= j) A.out[dj + (j - i)] = acc[r * 7 + c];
                ++rr;
            }
            ++cc;
        }
    }
}

int cbk_stiff(const CbStiffArgs &a, cudaStream_t s, long *launches)
{
    if (a.npairs == 0) return 0;
    unsigned g = (unsigned)((a.npairs + CB_TPB_K - 1) / CB_TPB_K);
    k_assemble_blocks<<<g, CB_TPB_K, 0, s>>>(a);
    ++*launches;
    return cudaGetLastError() != cudaSuccess;
}
