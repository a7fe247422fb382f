// cb_stiff.cu - tangent stiffness K_t and its assembly, fused.
//
// Replaces `ss = 0; stiff_tr; stiff_fr; stiff_sh; [stiff_br]` (main.c:1899-1921; truss.c:82-204,
// frame.c:226-362, shell.c:110-344, brick.c:79-397) together with the generic dense
// transform() (misc.c:41-69) and the skyline / dense scatter inlined in each of them.
//
// The global matrix is tiled by joint pairs (row joint A, column joint B).  A block receives the
// sub-blocks K_ab = T_a^T k_ab T_b of the elements containing both joints; the contribution list
// is sorted by (block, element type, element) - the order the reference adds them in.  The
// transformation is block diagonal, so a sub-block costs four sparse 3x3 triple products instead
// of a share of the reference's 2*18^3 dense MACs.
//
//   k_assemble_tiles   (CSC)      one CTA owns a run of consecutive joints, i.e. one contiguous
//                                 range of Ax.  Phase 1: one thread per contribution evaluates its
//                                 sub-block into shared memory (uniform work per thread).
//                                 Phase 2: segmented reduction over the sorted contribution list,
//                                 one thread per (block, column), into an output image of the
//                                 tile in shared memory.  Phase 3: the image is streamed to HBM with
//                                 fully coalesced stores.  No atomics, no colouring, no element
//                                 matrices staged in HBM, bit-reproducible.
//   k_assemble_blocks  (skyline)  one thread owns one block and loops its contributions; used for
//                                 the reference's skyline layout (small / medium models that keep
//                                 the SLVFLAG=0 host solver), where the output is not contiguous.
#include "cb_internal.h"
#include "cb_frame_math.cuh"

#define CB_TPB_K 128

// ---- sparse R^T S R pieces ------------------------------------------------------------------
// R rows are the local axes e1,e2,e3 (R[3*r+c]).  out is row-major with leading dimension ld.

// S = [[s00,s01,0],[s10,s11,0],[0,0,s22]]
__device__ __forceinline__ void rtsr_diag(const double *R, double s00, double s01, double s10,
                                          double s11, double s22, double *out, int ld)
{
    double W0[3], W1[3], W2[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        W0[q] = s00 * R[q] + s01 * R[3 + q];
        W1[q] = s10 * R[q] + s11 * R[3 + q];
        W2[q] = s22 * R[6 + q];
    }
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            out[p * ld + q] = R[p] * W0[q] + R[3 + p] * W1[q] + R[6 + p] * W2[q];
}

// S = [[0,0,0],[0,0,0],[a,b,0]]  (row w couples to theta_x, theta_y):  out = e3^T (a e1 + b e2)
__device__ __forceinline__ void rtsr_row(const double *R, double a, double b, double *out, int ld)
{
    double w[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) w[q] = a * R[q] + b * R[3 + q];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) out[p * ld + q] = R[6 + p] * w[q];
}

// S = [[0,0,a],[0,0,b],[0,0,0]]:  out = (a e1 + b e2)^T e3
__device__ __forceinline__ void rtsr_col(const double *R, double a, double b, double *out, int ld)
{
    double v[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) v[p] = a * R[p] + b * R[3 + p];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) out[p * ld + q] = v[p] * R[6 + q];
}

// shape-function gradients of the CST in the reference local axes: node 0,1,2 -> (bx, by)*2A
__device__ __forceinline__ void cst_grad(int n, double X2, double X3, double Y3, double &bx,
                                         double &by)
{
    bx = (n == 0) ? -Y3 : (n == 1 ? Y3 : 0.0);
    by = (n == 0) ? (X3 - X2) : (n == 1 ? -X3 : X2);
}

// K_ab (6x6, global axes) of a yielded shell (ANAFLAG 3): the local elasto-plastic block left in
// sh_kpl by k_shell_plastic_prep (stiffm_sh, shell.c:842-1133) is full - membrane and bending
// couple - plus the geometric term g on the translations (stiffg_sh), rotated by blockdiag(R, R).
// out(p, q) is written through the caller's indexing: out[p * ldr + q * ldc].
__device__ __noinline__ void shell_block_plastic(const CbDev &d, int e, int a, int b, double *out,
                                                 int ldr, int ldc)
{
    const double *kr = d.sh_Nm + (long)e * CB_SH_KREC;
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = kr[i];
    double bxa, bya, bxb, byb;
    cst_grad(a, kr[9], kr[10], kr[11], bxa, bya);
    cst_grad(b, kr[9], kr[10], kr[11], bxb, byb);
    const double g = bxa * (kr[15] * bxb + kr[17] * byb) + bya * (kr[17] * bxb + kr[16] * byb);
    const double *S = d.sh_kpl + (long)e * 324 + (6 * a) * 18 + 6 * b;
    for (int bi = 0; bi < 2; ++bi)
        for (int bj = 0; bj < 2; ++bj) {
            double s[3][3], W[3][3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    s[i][j] = S[(3 * bi + i) * 18 + 3 * bj + j] + ((bi == 0 && bj == 0 && i == j) ? g : 0.0);
            for (int i = 0; i < 3; ++i)
                for (int q = 0; q < 3; ++q) W[i][q] = s[i][0] * R[q] + s[i][1] * R[3 + q] + s[i][2] * R[6 + q];
            for (int p = 0; p < 3; ++p)
                for (int q = 0; q < 3; ++q)
                    out[(3 * bi + p) * ldr + (3 * bj + q) * ldc] = R[p] * W[0][q] + R[3 + p] * W[1][q] + R[6 + p] * W[2][q];
        }
}

// K_ab (6x6, global axes) of shell e -> blk (row-major, leading dimension ld)
// membrane  shell.c:487-531, DKT bending (precomputed) shell.c:533-658, drilling shell.c:482-484,
// geometric shell.c:660-840, rotation shell.c:285-305 + misc.c:41-69
__device__ __forceinline__ void shell_block(const CbDev &d, int e, int a, int b, double *blk, int ld)
{
    if (d.ANAFLAG == 3 && d.sh_yv[e] != 0) { shell_block_plastic(d, e, a, b, blk, ld, 1); return; }
    const double2 *kr2 = reinterpret_cast<const double2 *>(d.sh_Nm + (long)e * CB_SH_KREC);
    double kr[CB_SH_KREC];
#pragma unroll
    for (int i = 0; i < CB_SH_KREC / 2; ++i) { double2 v = kr2[i]; kr[2 * i] = v.x; kr[2 * i + 1] = v.y; }
    const double *R = kr;
    double bxa, bya, bxb, byb;
    cst_grad(a, kr[9], kr[10], kr[11], bxa, bya);
    cst_grad(b, kr[9], kr[10], kr[11], bxb, byb);
    const double m00 = kr[12] * bxa * bxb + kr[14] * bya * byb;
    const double m01 = kr[13] * bxa * byb + kr[14] * bya * bxb;
    const double m10 = kr[13] * bya * bxb + kr[14] * bxa * byb;
    const double m11 = kr[12] * bya * byb + kr[14] * bxa * bxb;
    const double g = bxa * (kr[15] * bxb + kr[17] * byb) + bya * (kr[17] * bxb + kr[16] * byb);
    // bending 3x3 sub-block (rows w,tx,ty of a; cols of b), stored block-contiguous
    double kbv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) kbv[i] = SOA(d.sh_keb, (3 * a + b) * 9 + i, e, d.NE_SH);
    const double k00 = kbv[0], k01 = kbv[1], k02 = kbv[2];
    const double k10 = kbv[3], k11 = kbv[4], k12 = kbv[5];
    const double k20 = kbv[6], k21 = kbv[7], k22 = kbv[8];
    const double drill = (a == b) ? k11 / 10000 : 0.0;
    rtsr_diag(R, m00 + g, m01, m10, m11 + g, k00 + g, blk, ld);             // translation-translation
    rtsr_row(R, k01, k02, blk + 3, ld);                                     // translation-rotation
    rtsr_col(R, k10, k20, blk + 3 * ld, ld);                                // rotation-translation
    rtsr_diag(R, k11, k12, k21, k22, drill, blk + 3 * ld + 3, ld);          // rotation-rotation
}

// K_ab (7x7, global axes w.r.t. the joints) of frame e: local tangent (frame.c:364-579, releases
// 798-900), rotation by blockdiag(R,R,1) (frame.c:286-299) and, with member-end offsets, the
// rigid-link transformation (frame.c:304-323)
__device__ __noinline__ void frame_block(const CbStiffArgs &A, int e, int a, int b, double *blk, int ld)
{
    double k[14][14], eft[14];
    const double *fr = A.fr_frame + (long)e * CB_FR_FRAME;
    frame_local_k(A.d, e, A.fr_ef, A.fr_efFE, fr[9], k, eft);
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = fr[i];
    double W[7][7], K[7][7];
    for (int i = 0; i < 7; ++i) {                     // W = k_ab T_b
        const double *kr = &k[7 * a + i][7 * b];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                W[i][3 * q + j] = kr[3 * q] * R[j] + kr[3 * q + 1] * R[3 + j] + kr[3 * q + 2] * R[6 + j];
        W[i][6] = kr[6];
    }
    for (int j = 0; j < 7; ++j) {                     // K = T_a^T W
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                K[3 * p + i][j] = R[i] * W[3 * p][j] + R[3 + i] * W[3 * p + 1][j] + R[6 + i] * W[3 * p + 2][j];
        K[6][j] = W[6][j];
    }
    if (A.d.fr_osflag[e] != 0) {
        // L = [[I, S],[0, I]] (+1 on warping), S from the end offsets; K <- L_a^T K L_b
        const double *oa = A.d.fr_offset + (long)e * 6 + 3 * a, *ob = A.d.fr_offset + (long)e * 6 + 3 * b;
        const double Sa[3][3] = {{0, oa[2], -oa[1]}, {-oa[2], 0, oa[0]}, {oa[1], -oa[0], 0}};
        const double Sb[3][3] = {{0, ob[2], -ob[1]}, {-ob[2], 0, ob[0]}, {ob[1], -ob[0], 0}};
        for (int i = 0; i < 7; ++i)
            for (int j = 0; j < 3; ++j)
                K[i][3 + j] += K[i][0] * Sb[0][j] + K[i][1] * Sb[1][j] + K[i][2] * Sb[2][j];
        for (int j = 0; j < 7; ++j)
            for (int i = 0; i < 3; ++i)
                K[3 + i][j] += Sa[0][i] * K[0][j] + Sa[1][i] * K[1][j] + Sa[2][i] * K[2][j];
    }
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) blk[i * ld + j] = K[i][j];
}

// Same block with the local joints known at compile time: every index into the 14x14 local matrix
// is static, so it lives in registers and the three quarters of it this block does not need are
// never computed (the generic version above keeps it in local memory: 4 KB of stack per thread).
// Members with end releases (static condensation, runtime pivots) take the generic path.
// OFF = false: the model has no member-end offsets (CbDev::fr_simple) - the rigid-link branch and its
// osflag load are compiled out.
// Entry (i, j) of the block goes to stg[i * sr + j * sc]: a stage column (sr = 7 * STR, sc = STR) or
// the tile's output image itself (sr = 1, sc = column height).
template <int LA, int LB, bool PL, bool OFF = true>
__device__ __noinline__ void frame_block_t(const CbStiffArgs &A, int e, double *stg, int sr, int sc)
{
    double k[14][14], eft[14], fr[CB_FR_FRAME], fc[10];
    {
        double ea[14], eb[14];                        // 16-byte vector loads of the AoS records
        ldv2<CB_FR_FRAME>(A.fr_frame + (long)e * CB_FR_FRAME, fr);
        ldv2<10>(A.d.fr_const + (long)e * CB_FR_CONST, fc);
        ldv2<14>(A.fr_ef + (long)e * 14, ea);
        ldv2<14>(A.fr_efFE + (long)e * 14, eb);
#pragma unroll
        for (int i = 0; i < 14; ++i) {
#pragma unroll
            for (int j = 0; j < 14; ++j) k[i][j] = 0;
            eft[i] = ea[i] + eb[i];
        }
    }
    frame_elastic_rcp(k, fc);
    if (A.d.ANAFLAG >= 2) frame_geometric_rcp(k, eft, fr[9], fc[2], fc[8]);
    if (PL) {                     // ANAFLAG 3 only: its own instantiation (register pressure)
        const int y0 = A.d.fr_yldflag[(long)e * 2], y1 = A.d.fr_yldflag[(long)e * 2 + 1];
        if (y0 != 2 || y1 != 2) frame_plastic(k, eft, y0, y1, A.d.fr_plast + (long)e * 3);
    }
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = fr[i];
    double W[7][7], K[7][7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {                     // W = k_ab T_b
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                W[i][3 * q + j] = k[7 * LA + i][7 * LB + 3 * q] * R[j] + k[7 * LA + i][7 * LB + 3 * q + 1] * R[3 + j] +
                                  k[7 * LA + i][7 * LB + 3 * q + 2] * R[6 + j];
        W[i][6] = k[7 * LA + i][7 * LB + 6];
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {                     // K = T_a^T W
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                K[3 * p + i][j] = R[i] * W[3 * p][j] + R[3 + i] * W[3 * p + 1][j] + R[6 + i] * W[3 * p + 2][j];
        K[6][j] = W[6][j];
    }
    if (OFF && A.d.fr_osflag[e] != 0) {
        const double *oa = A.d.fr_offset + (long)e * 6 + 3 * LA, *ob = A.d.fr_offset + (long)e * 6 + 3 * LB;
        const double Sa[3][3] = {{0, oa[2], -oa[1]}, {-oa[2], 0, oa[0]}, {oa[1], -oa[0], 0}};
        const double Sb[3][3] = {{0, ob[2], -ob[1]}, {-ob[2], 0, ob[0]}, {ob[1], -ob[0], 0}};
#pragma unroll
        for (int i = 0; i < 7; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                K[i][3 + j] += K[i][0] * Sb[0][j] + K[i][1] * Sb[1][j] + K[i][2] * Sb[2][j];
#pragma unroll
        for (int j = 0; j < 7; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                K[3 + i][j] += Sa[0][i] * K[0][j] + Sa[1][i] * K[1][j] + Sa[2][i] * K[2][j];
    }
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) stg[i * sr + j * sc] = K[i][j];
}

// K_ab (3x3) of 8-node brick e: 2x2x2 Gauss, B_a^T C B_b detJ (brick.c:79-397, jacob 541-699).
// For the isotropic C of brick.c:127-141 (lambda = e1, mu = e2, lambda + 2 mu = e3) the 6x24
// strain-displacement product collapses to
//   K_ab[i][j] = detJ * (lambda g_a[i] g_b[j] + mu g_a[j] g_b[i] + mu delta_ij g_a.g_b),
// g_n = J^-1 dN_n/d(r,s,t).  Bricks are linear and assembled once.
// The Jacobian part is common to the 64 joint pairs of a brick: k_brick_prep evaluates it once per
// (brick, Gauss point) - inverse Jacobian (9) and detJ, same expressions and order as jacob() - and
// brick_block only forms g_a, g_b and the 3x3 update (5x fewer flops per contribution, no coordinate
// gather, no local arrays).
#define CB_BR_PREP 10            // doubles per (brick, Gauss point): J^-1 row-major, detJ
__constant__ int c_br_sg[8] = {+1, -1, +1, -1, -1, +1, -1, +1};
__constant__ int c_br_sr[8] = {+1, -1, -1, +1, +1, -1, -1, +1};
__constant__ int c_br_ss[8] = {+1, +1, -1, -1, +1, +1, -1, -1};
__constant__ int c_br_st[8] = {+1, +1, +1, +1, -1, -1, -1, -1};

__global__ void __launch_bounds__(256)
k_brick_prep(CbStiffArgs A, double *__restrict__ prep)
{
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= A.d.NE_BR * 8) return;
    const long e = i >> 3;
    const int q = (int)(i & 7);
    const double gp = 0.57735026918962576451;      // 1/sqrt(3)
    const double R = (q & 4) ? -gp : gp, S = (q & 2) ? -gp : gp, T = (q & 1) ? -gp : gp;
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const long jt = A.d.br_nodes[e * 8 + n];
        const double X0 = A.x[jt * 3], X1 = A.x[jt * 3 + 1], X2 = A.x[jt * 3 + 2];
        const double dr = c_br_sg[n] * (S + c_br_ss[n]) * (T + c_br_st[n]) / 8.0;
        const double ds = c_br_sg[n] * (R + c_br_sr[n]) * (T + c_br_st[n]) / 8.0;
        const double dt = c_br_sg[n] * (R + c_br_sr[n]) * (S + c_br_ss[n]) / 8.0;
        J[0][0] += dr * X0; J[1][0] += ds * X0; J[2][0] += dt * X0;
        J[0][1] += dr * X1; J[1][1] += ds * X1; J[2][1] += dt * X1;
        J[0][2] += dr * X2; J[1][2] += ds * X2; J[2][2] += dt * X2;
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                 c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    double *o = prep + i * CB_BR_PREP;
    o[0] = c00 / det; o[1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det; o[2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    o[3] = c01 / det; o[4] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; o[5] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    o[6] = c02 / det; o[7] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det; o[8] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    o[9] = det;
}

__device__ __noinline__ void brick_block(const CbStiffArgs &A, int e, int a, int b, double *blk, int ld)
{
    const double E = A.d.br_const[(long)e * 4], v = A.d.br_const[(long)e * 4 + 1];
    const double lam = E * v / ((1 + v) * (1 - 2 * v)), mu = .5 * (E / (1 + v));
    // natural-coordinate signs of local joints a, b (the c_br_* tables as arithmetic: a table indexed by a
    // lane-varying joint serialises in the constant cache)
    const int sra = 1 - 2 * (((a + 1) >> 1) & 1), ssa = 1 - (a & 2), sta = 1 - ((a & 4) >> 1), sga = sra * ssa * sta;
    const int srb = 1 - 2 * (((b + 1) >> 1) & 1), ssb = 1 - (b & 2), stb = 1 - ((b & 4) >> 1), sgb = srb * ssb * stb;
    double K[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const double gp = 0.57735026918962576451;      // 1/sqrt(3)
    const double2 *pr = reinterpret_cast<const double2 *>(A.br_prep + (long)e * 8 * CB_BR_PREP);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const double R = (q & 4) ? -gp : gp, S = (q & 2) ? -gp : gp, T = (q & 1) ? -gp : gp;
        double Ji[10];
#pragma unroll
        for (int i = 0; i < 5; ++i) { const double2 t2 = pr[q * 5 + i]; Ji[2 * i] = t2.x; Ji[2 * i + 1] = t2.y; }
        const double det = Ji[9];
        const double da[3] = {sga * (S + ssa) * (T + sta) / 8.0, sga * (R + sra) * (T + sta) / 8.0,
                              sga * (R + sra) * (S + ssa) / 8.0};
        const double db[3] = {sgb * (S + ssb) * (T + stb) / 8.0, sgb * (R + srb) * (T + stb) / 8.0,
                              sgb * (R + srb) * (S + ssb) / 8.0};
        double ga[3], gb[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            ga[i] = Ji[3 * i] * da[0] + Ji[3 * i + 1] * da[1] + Ji[3 * i + 2] * da[2];
            gb[i] = Ji[3 * i] * db[0] + Ji[3 * i + 1] * db[1] + Ji[3 * i + 2] * db[2];
        }
        const double gg = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                K[i][j] += det * (lam * ga[i] * gb[j] + mu * ga[j] * gb[i] + ((i == j) ? mu * gg : 0.0));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) blk[i * ld + j] = K[i][j];
}

__device__ __noinline__ void brick_mass_block(const CbStiffArgs &A, int e, int a, int b, double *blk, int ld)
{
    const int sr[8] = {+1, -1, -1, +1, +1, -1, -1, +1};
    const int ss[8] = {+1, +1, -1, -1, +1, +1, -1, -1};
    const int st[8] = {+1, +1, +1, +1, -1, -1, -1, -1};
    const double rho = A.d.br_const[(long)e * 4 + 2];
    double X[8][3];
    for (int n = 0; n < 8; ++n) {
        const long jt = A.d.br_nodes[(long)e * 8 + n];
        for (int m = 0; m < 3; ++m) X[n][m] = A.x[jt * 3 + m];
    }
    const double gp = 0.57735026918962576451;
    double mab = 0;
    for (int q = 0; q < 8; ++q) {
        const double R = (q & 4) ? -gp : gp, S = (q & 2) ? -gp : gp, T = (q & 1) ? -gp : gp;
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int n = 0; n < 8; ++n) {
            const double dr = sr[n] * (1 + ss[n] * S) * (1 + st[n] * T) / 8.0;
            const double ds = ss[n] * (1 + sr[n] * R) * (1 + st[n] * T) / 8.0;
            const double dt = st[n] * (1 + sr[n] * R) * (1 + ss[n] * S) / 8.0;
            for (int m = 0; m < 3; ++m) { J[0][m] += dr * X[n][m]; J[1][m] += ds * X[n][m]; J[2][m] += dt * X[n][m]; }
        }
        const double det = J[0][0] * J[1][1] * J[2][2] - J[0][0] * J[1][2] * J[2][1] - J[0][1] * J[1][0] * J[2][2] +
                           J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] - J[0][2] * J[1][1] * J[2][0];
        const double ha = (1 + sr[a] * R) * (1 + ss[a] * S) * (1 + st[a] * T) / 8.0;
        const double hb = (1 + sr[b] * R) * (1 + ss[b] * S) * (1 + st[b] * T) / 8.0;
        mab += rho * (ha * hb) * det;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) blk[i * ld + j] = (i == j) ? mab : 0.0;
}

// K_ab (3x3) of truss e (truss.c:102-166)
__device__ __noinline__ void truss_block(const CbStiffArgs &A, int e, int a, int b, double *blk,
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
    double kab = s * k;
    if (A.d.ANAFLAG == 3) {
        // stiffm_tr (truss.c:206-229): plastic reduction of the 2x2 axial matrix beyond the squash
        // load; the geometric term only once the member is on the yield surface (truss.c:155-156)
        const double Py = A.d.tr_py[e], f0 = A.tr_ef[(long)e * 2], f1 = A.tr_ef[(long)e * 2 + 1];
        const double r2 = (f0 / Py) * (f0 / Py);
        if (r2 > 1 + 1e-4) {
            const double G0 = 2 * f0 / (Py * Py), G1 = 2 * f1 / (Py * Py);
            const double kg0 = k * G0 - k * G1, kg1 = -k * G0 + k * G1;
            const double gkg = kg0 * G0 + kg1 * G1;
            kab -= (a == 0 ? kg0 : kg1) * (b == 0 ? kg0 : kg1) / gkg;
        }
        if (!(r2 >= 1 - 1e-4)) gN = 0.0;
    }
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            blk[p * ld + q] = kab * c[p] * c[q] + ((p == q) ? s * gN : 0.0);
}

// ------------------------------------------------------------------------------------------
// CSC: tile kernel.  ND = largest DOF count per joint among the model's element types.
// Persistent CTAs (grid = resident CTAs), software-pipelined over tiles: the header, the
// contribution record and the element inputs of the NEXT tile are requested while the current
// tile is being reduced and written out, so no HBM round trip sits on the critical path.
// shared memory: stage[ND*ND][CB_TILE_T+1] doubles | obuf[tile_smem_out] doubles |
//                spair[CB_TILE_T] pair records | ndof[CB_TILE_T]
// ------------------------------------------------------------------------------------------
struct ShellIn {               // inputs of one shell contribution
    double kr[CB_SH_KREC];     // stiffness-pass record of the element
    double kb[10];             // its 3x3 DKT sub-block for (a,b), contribution-ordered copy
};

__device__ __forceinline__ void shell_load(const CbStiffArgs &A, const CbContrib &ct, long cidx,
                                           ShellIn &in)
{
    const double2 *kr2 = reinterpret_cast<const double2 *>(A.d.sh_Nm + (long)ct.e * CB_SH_KREC);
#pragma unroll
    for (int i = 0; i < CB_SH_KREC / 2; ++i) { double2 v = __ldg(kr2 + i); in.kr[2 * i] = v.x; in.kr[2 * i + 1] = v.y; }
    const double2 *kb2 = reinterpret_cast<const double2 *>(A.kebc + cidx * 10);
#pragma unroll
    for (int i = 0; i < 5; ++i) { double2 v = __ldg(kb2 + i); in.kb[2 * i] = v.x; in.kb[2 * i + 1] = v.y; }
}

// K_ab (6x6, global axes) from preloaded inputs, written straight to the stage column of thread t
template <int ND>
__device__ __forceinline__ void shell_block_stage(const ShellIn &in, int a, int b, double *stg)
{
    constexpr int STR = CB_TILE_T + 1;
    const double *R = in.kr;
    double bxa, bya, bxb, byb;
    cst_grad(a, in.kr[9], in.kr[10], in.kr[11], bxa, bya);
    cst_grad(b, in.kr[9], in.kr[10], in.kr[11], bxb, byb);
    const double m00 = in.kr[12] * bxa * bxb + in.kr[14] * bya * byb;
    const double m01 = in.kr[13] * bxa * byb + in.kr[14] * bya * bxb;
    const double m10 = in.kr[13] * bya * bxb + in.kr[14] * bxa * byb;
    const double m11 = in.kr[12] * bya * byb + in.kr[14] * bxa * bxb;
    const double g = bxa * (in.kr[15] * bxb + in.kr[17] * byb) + bya * (in.kr[17] * bxb + in.kr[16] * byb);
    const double *kb = in.kb;
    const double drill = (a == b) ? kb[4] / 10000 : 0.0;
    double s[9];
    rtsr_diag(R, m00 + g, m01, m10, m11 + g, kb[0] + g, s, 3);
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) stg[(p * ND + q) * STR] = s[p * 3 + q];
    rtsr_row(R, kb[1], kb[2], s, 3);
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) stg[(p * ND + 3 + q) * STR] = s[p * 3 + q];
    rtsr_col(R, kb[3], kb[6], s, 3);
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) stg[((3 + p) * ND + q) * STR] = s[p * 3 + q];
    rtsr_diag(R, kb[4], kb[5], kb[7], kb[8], drill, s, 3);
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) stg[((3 + p) * ND + 3 + q) * STR] = s[p * 3 + q];
}

// mass mode of the tile kernel (cb_mass with bricks): consistent brick mass, lumped shell mass on
// the diagonal of the joint's own block; returns the DOF count of the contribution's joints
template <int ND>
__device__ __noinline__ int mass_block_stage(const CbStiffArgs &A, const CbContrib &ct, double *stg)
{
    constexpr int STR = CB_TILE_T + 1;
    if (A.mixed && ND > 3 && ct.type != CB_T_SHELL && ct.type != CB_T_FRAME) {     // see the stiffness mode
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) stg[(p * ND + q) * STR] = 0.0;
    } else {
        for (int i = 0; i < ND * ND; ++i) stg[i * STR] = 0.0;
    }
    if (ct.type == CB_T_BRICK) {
        double blk[9];
        brick_mass_block(A, ct.e, ct.a, ct.b, blk, 3);
        if (ct.e >= A.d.NE_SBR) { stg[0] = blk[8]; return 1; }      // fluid: Q, with dens = 1 / c^2 (brick.c:63-69)
        for (int p = 0; p < 3; ++p) stg[(p * ND + p) * STR] = blk[p * 3 + p];
        return 3;
    }
    if (ct.type == CB_T_COUPLE) {                                    // [M 0; -rho L^T Q] (fsi.c:436-443)
        if (ct.a == 1 && ct.b == 0) {
            const double *L = A.d.cp_L + (long)ct.e * 4;
            for (int q = 0; q < 3; ++q) stg[q * STR] = -1 * A.d.fdens * L[q];
        }
        return 3;
    }
    if (ct.type == CB_T_SHELL && ND >= 6) {
        if (ct.a == ct.b) {             // lumped: rho A t / 3, rotations * t^2 / 12 (shell.c:1548-1558)
            const double th = SOA(A.d.sh_const, 2, ct.e, A.d.NE_SH);
            const double Mtot = A.sh_dens[ct.e] * SOA(A.d.sh_const, 4, ct.e, A.d.NE_SH) * th;
            for (int p = 0; p < 6; ++p)
                stg[(p * ND + p) * STR] = (p < 3) ? Mtot / 3 : Mtot / 3 * (th * th) / 12;
        }
        return 6;
    }
    return (ct.type == CB_T_FRAME) ? 7 : 3;
}

// SHELL_ONLY: the model holds nothing but DKT shells - the other element branches are compiled
// out so that they cannot cost the hot configuration registers.  FRAME_SIMPLE: nothing but frames
// without end releases, rigid offsets or plasticity (CbDev::fr_simple, BASELINE config 4): no
// dependent load of the release flags ahead of the element data, no generic 14x14 path (4 KB of
// stack), no mixed-DOF reduction.
template <int ND, bool SHELL_ONLY, bool FRAME_SIMPLE = false>
__global__ void __launch_bounds__(CB_TILE_T, (ND <= 6) ? 4 : 3)
k_assemble_tiles(CbStiffArgs A)
{
    extern __shared__ double smem[];
    constexpr int STR = CB_TILE_T + 1;
    constexpr int NN = ND * ND;
    double *stage = smem;
    double *obuf = smem + ((NN * STR + 1) & ~1);           // 16-byte aligned
    CbTPair *spair = reinterpret_cast<CbTPair *>(obuf + A.tile_smem_out);       // [CB_TILE_T]
    unsigned char *ndof = reinterpret_cast<unsigned char *>(spair + CB_TILE_T);
    const int t = threadIdx.x;

    long tile = blockIdx.x;
    if (tile >= A.ntiles) return;
    CbTile tl = A.tiles[tile];
    CbContrib ct{}; ct.type = 0xff;
    CbTDst td{};
    if (t < tl.ns) { ct = A.tcontribs[tl.t0 + t]; if (FRAME_SIMPLE) td = A.tdst[tl.t0 + t]; }
    ShellIn in;
    if (!FRAME_SIMPLE && ND >= 6 && ct.type == CB_T_SHELL && !A.mass_mode) shell_load(A, ct, (long)tl.c0 + ct.pad, in);

    for (;;) {
        const long next = tile + gridDim.x;
        const bool has_next = next < A.ntiles;
        CbTile tln = tl;
        if (has_next) tln = A.tiles[next];                 // requested now, needed after phase 1
        if (t < tl.np) spair[t] = A.tpairs[tl.p0 + t];

        // ---- phase 1: one contribution per thread -> its column of `stage` ---------------------
        if (ct.type != 0xff) {
            // FRAME_SIMPLE: the block goes to the column of the THREAD (consecutive lanes, consecutive
            // doubles: conflict-free stores; the planner's slots are not in column order) and ndof,
            // which this instantiation does not need, maps reference-order columns to thread columns
            double *stg = stage + (FRAME_SIMPLE ? t : ct.pad);
            const int col = ct.pad;             // column of `stage` / entry of ndof: reference order
            if constexpr (FRAME_SIMPLE) {
                ndof[col] = (unsigned char)t;
                // the only contribution of a fully free joint pair lands in the output image at once
                int sr = 7 * STR, sc = STR;
                if (td.direct) { stg = obuf + (int)((tl.out0 + A.out_par) & 1) + td.rel; sr = 1; sc = td.colh; }
                if (ct.a == 0) {
                    if (ct.b == 0) frame_block_t<0, 0, false, false>(A, ct.e, stg, sr, sc); else frame_block_t<0, 1, false, false>(A, ct.e, stg, sr, sc);
                } else {
                    if (ct.b == 0) frame_block_t<1, 0, false, false>(A, ct.e, stg, sr, sc); else frame_block_t<1, 1, false, false>(A, ct.e, stg, sr, sc);
                }
            } else if (A.mass_mode) {
                ndof[col] = (unsigned char)mass_block_stage<ND>(A, ct, stg);
            } else if (ct.type == CB_T_SHELL) {
                if constexpr (ND >= 6) {
                    if (ND > 6) {
#pragma unroll
                        for (int i = 0; i < NN; ++i) stg[i * STR] = 0.0;
                    }
                    if (A.d.ANAFLAG == 3 && A.d.sh_yv[ct.e] != 0)
                        shell_block_plastic(A.d, ct.e, ct.a, ct.b, stg, ND * STR, STR);
                    else
                        shell_block_stage<ND>(in, ct.a, ct.b, stg);
                }
                ndof[col] = 6;
            } else if (SHELL_ONLY) {
            } else if (ct.type == CB_T_FRAME) {
                if constexpr (ND >= 7) {
                    if (A.d.fr_mendrel[(long)ct.e * 5] == 1) {
                        double blk[49];
                        frame_block(A, ct.e, ct.a, ct.b, blk, 7);
#pragma unroll
                        for (int i = 0; i < 49; ++i) stg[i * STR] = blk[i];
                    } else if (A.d.ANAFLAG == 3) {
                        if (ct.a == 0) {
                            if (ct.b == 0) frame_block_t<0, 0, true>(A, ct.e, stg, 7 * STR, STR); else frame_block_t<0, 1, true>(A, ct.e, stg, 7 * STR, STR);
                        } else {
                            if (ct.b == 0) frame_block_t<1, 0, true>(A, ct.e, stg, 7 * STR, STR); else frame_block_t<1, 1, true>(A, ct.e, stg, 7 * STR, STR);
                        }
                    } else if (ct.a == 0) {
                        if (ct.b == 0) frame_block_t<0, 0, false>(A, ct.e, stg, 7 * STR, STR); else frame_block_t<0, 1, false>(A, ct.e, stg, 7 * STR, STR);
                    } else {
                        if (ct.b == 0) frame_block_t<1, 0, false>(A, ct.e, stg, 7 * STR, STR); else frame_block_t<1, 1, false>(A, ct.e, stg, 7 * STR, STR);
                    }
                }
                ndof[col] = 7;
            } else {
                double blk[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) blk[i] = 0.0;
                if (ct.type == CB_T_TRUSS) truss_block(A, ct.e, ct.a, ct.b, blk, 3);
                else if (ct.type == CB_T_BRICK) brick_block(A, ct.e, ct.a, ct.b, blk, 3);
                // a 3-DOF contribution in a wider stage column: the mixed reduction reads its leading 3x3 only
                // (r, c < ndof), so the other 27 of 36 slots need no zeros (they were 80 % of this kernel's
                // shared-memory stores on the brick + skin model)
                if (A.mixed && ND > 3) {
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int q = 0; q < 3; ++q) stg[(p * ND + q) * STR] = 0.0;
                } else {
#pragma unroll
                    for (int i = 0; i < NN; ++i) stg[i * STR] = 0.0;
                }
                if (ct.type == CB_T_BRICK && ct.e >= A.d.NE_SBR) {
                    // fluid brick (fsi.c:344, model.c:1089-1140): only the (z, z) entry of the joint block has an
                    // equation - the pressure DOF, DOF 0 of the twin joints; with the reference's E = 1e20,
                    // nu = 0.5e20 (brick.c:72-73) it is detJ grad N_a . grad N_b, the acoustic H matrix
                    stg[0] = blk[8];
                    ndof[col] = 1;
                } else if (ct.type == CB_T_COUPLE) {
                    // [K L; 0 H] (fsi.c:376-383): L = tarea * nnorm couples the joint's translations to its pressure
                    if (ct.a == 0 && ct.b == 1) {
                        const double *L = A.d.cp_L + (long)ct.e * 4;
#pragma unroll
                        for (int p = 0; p < 3; ++p) stg[(p * ND) * STR] = L[p];
                    }
                    ndof[col] = 3;
                } else {
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int q = 0; q < 3; ++q) stg[(p * ND + q) * STR] = blk[p * 3 + q];
                    ndof[col] = 3;
                }
            }
        }
        // request the next tile's contribution record, then its element inputs; they land while
        // this tile is reduced and written out
        CbContrib ctn{}; ctn.type = 0xff;
        CbTDst tdn{};
        if (has_next && t < tln.ns) { ctn = A.tcontribs[tln.t0 + t]; if (FRAME_SIMPLE) tdn = A.tdst[tln.t0 + t]; }
        // pairs with more than one contribution come first (the planner sorts a tile's pairs by count)
        const int npd = __syncthreads_count(t < tl.np && spair[t].cnt > 1);
        if (!FRAME_SIMPLE && ND >= 6 && ctn.type == CB_T_SHELL && !A.mass_mode) shell_load(A, ctn, (long)tln.c0 + ctn.pad, in);

        // ---- phase 2: segmented reduction over the sorted contribution list, one thread per
        // (joint-pair block, column): ND rows of `stage` summed over the block's contributions in
        // reference order, written as one contiguous column run of the tile's output image.  The
        // image is shifted by the parity of out0 so phase 3 can use aligned 16-byte accesses.
        const int shift = (int)((tl.out0 + A.out_par) & 1);
        // FRAME_SIMPLE: the pairs that are reduced (npd of them, the joints' own blocks in a lattice) are walked
        // column-major - consecutive lanes take consecutive PAIRS of one column c, whose staged contributions sit
        // three thread columns apart: sixteen distinct banks.  Pair-major (seven lanes per pair) put the lanes of
        // neighbouring pairs on overlapping banks (a 2.5-way conflict on every load of the reduction).
        const int nred = FRAME_SIMPLE ? npd : 0;
        const int nitems = (tl.np - nred) * ND + nred * ND;
        for (int it = t; it < nitems; it += CB_TILE_T) {
            int p, c;
            if (FRAME_SIMPLE && it < nred * ND) { c = it / nred; p = it - c * nred; }
            else { p = it / ND; c = it - p * ND; }
            const CbTPair pr = spair[p];
            constexpr unsigned FULL = (1u << ND) - 1u;
            if (FRAME_SIMPLE && pr.cnt == 1 && pr.maskA == FULL && pr.maskB == FULL) continue;   // written in phase 1
            if (pr.maskA == FULL && pr.maskB == FULL && (FRAME_SIMPLE || !A.mixed)) {
                const double *src = stage + c * STR + (FRAME_SIMPLE ? 0 : pr.cs);
                double acc[ND];
                const int s0 = FRAME_SIMPLE ? ndof[pr.cs] : 0;
#pragma unroll
                for (int r = 0; r < ND; ++r) acc[r] = src[r * ND * STR + s0];
                for (int q = 1; q < pr.cnt; ++q) {
                    const int sq = FRAME_SIMPLE ? ndof[pr.cs + q] : q;
#pragma unroll
                    for (int r = 0; r < ND; ++r) acc[r] += src[r * ND * STR + sq];
                }
                double *dst = obuf + shift + pr.rel + c * pr.colh;
#pragma unroll
                for (int r = 0; r < ND; ++r) dst[r] = acc[r];
            } else {
                if (!((pr.maskB >> c) & 1)) continue;
                const int cc = __popc(pr.maskB & ((1u << c) - 1));
                double *dst = obuf + shift + pr.rel + cc * pr.colh;
                int rr = 0;
                for (int r = 0; r < ND; ++r) {
                    if (!((pr.maskA >> r) & 1)) continue;
                    const double *src = stage + (r * ND + c) * STR + (FRAME_SIMPLE ? 0 : pr.cs);
                    double sum = 0.0;
                    for (int q = 0; q < pr.cnt; ++q)
                        if (FRAME_SIMPLE) sum += src[ndof[pr.cs + q]];
                        else if (!A.mixed || (r < ndof[pr.cs + q] && c < ndof[pr.cs + q])) sum += src[q];
                    dst[rr++] = sum;
                }
            }
        }
        __syncthreads();

        // ---- phase 3: stream the tile's output image to HBM (coalesced, 16-byte vectors) --------
        {
            double *dst = A.out + tl.out0;
            const double *img = obuf + shift;
            if (shift && t == 0) dst[0] = img[0];
            const int nv = (tl.nout - shift) >> 1;                   // aligned pairs
            const double2 *img2 = reinterpret_cast<const double2 *>(img + shift);
            double2 *dst2 = reinterpret_cast<double2 *>(dst + shift);
            for (int i = t; i < nv; i += CB_TILE_T) dst2[i] = img2[i];
            const int tail = shift + 2 * nv;
            if (tail < tl.nout && t == CB_TILE_T - 1) dst[tail] = img[tail];
        }

        if (!has_next) break;
        tile = next; tl = tln; ct = ctn; td = tdn;
        __syncthreads();          // obuf / spair are rewritten by the next iteration
    }
}

// ------------------------------------------------------------------------------------------
// CSC, shell-only models: "duo" tile kernel.  One persistent CTA walks tiles (runs of consecutive
// joints whose CSC columns are one contiguous slice of Ax), software-pipelined:
//   * the krec records of the tile's distinct shells and its pair records are brought into shared
//     memory by cp.async, double-buffered (the next tile's land while this tile is processed);
//     the work item and the DKT sub-blocks of the next tile are prefetched into registers;
//   * a thread evaluates up to two consecutive contributions of one joint-pair block, three
//     columns at a time, adds them in registers and writes the result into the tile's output
//     image in shared memory with 16-byte stores (lanes arranged by the planner so that the
//     stores of a quarter-warp hit distinct bank groups);
//   * a block with more than two contributions (diagonal blocks: one per adjacent shell) is a
//     group of consecutive lanes of one warp: the leader stores, the followers add their partial
//     sums to the image one after the other (list order, __syncwarp between the rounds);
//   * the finished image leaves as ONE bulk asynchronous copy shared -> global (TMA engine,
//     cp.async.bulk): no thread touches the output on its way to HBM.
// shared memory: image | shell records x2 | DKT entries | pair records x2 | work items | element
// ids | ring of tile records
// ------------------------------------------------------------------------------------------

#ifndef CB_T2_MBAR
#define CB_T2_MBAR 1              // the wait for the previous image's bulk read-out is taken off the CTA barrier
#endif
#ifndef CB_T2_CTAS
#define CB_T2_CTAS 4              // resident CTAs per SM the kernel is compiled for (96 threads each)
#endif
#ifndef CB_T2_CTAS_CLS
#define CB_T2_CTAS_CLS 4          // ... with the class table
#endif

// Columns 0..2 (LEFT) or 3..5 of K_ab (6x6, global axes) of one shell contribution:
// top[9] = rows 0..2 (translations), bot[9] = rows 3..5 (rotations), row-major 3x3; kb = the 3x3
// DKT sub-block for (a,b).  first: overwrite, else accumulate (second contribution of the block).
template <bool LEFT>
__device__ __forceinline__ void shell_half_acc(const double *kr, const double *kb, int a, int b,
                                               double *top, double *bot, bool first)
{
    const double *R = kr;
    double s[9];
#define CB_PUT(dst)                                                                               \
    _Pragma("unroll") for (int i = 0; i < 9; ++i) { if (first) dst[i] = s[i]; else dst[i] += s[i]; }
    if (LEFT) {
        double bxa, bya, bxb, byb;
        cst_grad(a, kr[9], kr[10], kr[11], bxa, bya);
        cst_grad(b, kr[9], kr[10], kr[11], bxb, byb);
        const double m00 = kr[12] * bxa * bxb + kr[14] * bya * byb;
        const double m01 = kr[13] * bxa * byb + kr[14] * bya * bxb;
        const double m10 = kr[13] * bya * bxb + kr[14] * bxa * byb;
        const double m11 = kr[12] * bya * byb + kr[14] * bxa * bxb;
        const double g = bxa * (kr[15] * bxb + kr[17] * byb) + bya * (kr[17] * bxb + kr[16] * byb);
        rtsr_diag(R, m00 + g, m01, m10, m11 + g, kb[0] + g, s, 3);
        CB_PUT(top)
        rtsr_col(R, kb[3], kb[6], s, 3);
        CB_PUT(bot)
    } else {
        const double drill = (a == b) ? kb[4] / 10000 : 0.0;
        rtsr_row(R, kb[1], kb[2], s, 3);
        CB_PUT(top)
        rtsr_diag(R, kb[4], kb[5], kb[7], kb[8], drill, s, 3);
        CB_PUT(bot)
    }
#undef CB_PUT
}

// ---- asynchronous staging (cp.async: no register scoreboard is held while a copy is in flight) ----
#define CB_T2_EIDS ((CB_T2_ELEMS * 9 + CB_T2_T - 1) / CB_T2_T)
__device__ __forceinline__ unsigned t2_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
#define CB_CPA(BYTES, CACHE, dst, src)                                                              \
    asm volatile("cp.async." CACHE ".shared.global [%0], [%1], " #BYTES ";" ::"r"(t2_saddr(dst)), "l"(src))
#define CB_CPA_COMMIT() asm volatile("cp.async.commit_group;")

// element ids of a tile's copy items (item t + k * CB_T2_T -> seid[k][t], private to thread t)
__device__ __forceinline__ void t2_issue_eids(const CbStiffArgs &A, const CbTile2 &tl, int *seid)
{
#pragma unroll
    for (int k = 0; k < CB_T2_EIDS; ++k) {
        const int i = threadIdx.x + k * CB_T2_T;
        if (i < tl.ne * 9) CB_CPA(4, "ca", seid + k * CB_T2_T + threadIdx.x, A.tile_elems + tl.e0 + i / 9);
    }
}
// shell records of a tile (9 chunks of 16 bytes each) + its pair records
__device__ __forceinline__ void t2_issue_stage(const CbStiffArgs &A, const CbTile2 &tl, const int *eid,
                                               double *krec_dst, CbTPair *pair_dst)
{
#pragma unroll
    for (int k = 0; k < CB_T2_EIDS; ++k) {
        const int i = threadIdx.x + k * CB_T2_T;
        if (i < tl.ne * 9) {
            const int es = i / 9, ch = i - es * 9;
            CB_CPA(16, "cg", krec_dst + es * CB_SH_KREC + ch * 2, A.d.sh_Nm + (long)eid[k] * CB_SH_KREC + ch * 2);
        }
    }
    if (threadIdx.x < tl.np) CB_CPA(16, "cg", pair_dst + threadIdx.x, A.tpairs2 + tl.p0 + threadIdx.x);
}
// DKT sub-blocks of the two contributions of work item t of tile T: kebc[(T * 18 + u * 9 + i) *
// CB_T2_T + t] (u = contribution 0/1, i = 3x3 entry) -> skb[u * 9 + i][t], private to thread t.
// The left half of a block takes entries 0,3,6, the right half the other six.
template <bool LEFT>
__device__ __forceinline__ void t2_issue_kb(const CbStiffArgs &A, long T, int t, double *skb)
{
    const double *src = A.kebc + T * (18L * CB_T2_T) + t;
#pragma unroll
    for (int i = 0; i < 18; ++i)
        if (((i % 3) == 0) == LEFT) CB_CPA(8, "ca", skb + i * CB_T2_T + t, src + i * CB_T2_T);
}

// columns c0..c0+2 of a block into the tile image (ACC: added to what is there); top/bot as in
// shell_half_acc
template <bool ACC>
__device__ __forceinline__ void t2_store_half(double *obuf, int shift, const CbTPair &pr, int c0,
                                              const double *top, const double *bot)
{
    double *img = obuf + shift + pr.rel;
    if (pr.maskA == 0x3f && pr.maskB == 0x3f) {
        if ((((shift + pr.rel) | pr.colh) & 1) == 0) {          // 16-byte aligned columns
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                double2 *col = reinterpret_cast<double2 *>(img + (c0 + q) * pr.colh);
                double2 v0 = make_double2(top[q], top[3 + q]);
                double2 v1 = make_double2(top[6 + q], bot[q]);
                double2 v2 = make_double2(bot[3 + q], bot[6 + q]);
                if (ACC) {
                    // one column at a time (keeps the live range of the old values short)
                    asm volatile("" ::: "memory");
                    const double2 o0 = col[0], o1 = col[1], o2 = col[2];
                    v0.x = o0.x + v0.x; v0.y = o0.y + v0.y; v1.x = o1.x + v1.x; v1.y = o1.y + v1.y;
                    v2.x = o2.x + v2.x; v2.y = o2.y + v2.y;
                }
                col[0] = v0; col[1] = v1; col[2] = v2;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                double *col = img + (c0 + q) * pr.colh;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    if (ACC) { col[r] += top[r * 3 + q]; col[3 + r] += bot[r * 3 + q]; }
                    else { col[r] = top[r * 3 + q]; col[3 + r] = bot[r * 3 + q]; }
                }
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int c = c0 + q;
            if (!((pr.maskB >> c) & 1)) continue;
            double *col = img + __popc(pr.maskB & ((1u << c) - 1)) * pr.colh;
            int rr = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r)
                if ((pr.maskA >> r) & 1) { if (ACC) col[rr++] += top[r * 3 + q]; else col[rr++] = top[r * 3 + q]; }
#pragma unroll
            for (int r = 0; r < 3; ++r)
                if ((pr.maskA >> (3 + r)) & 1) { if (ACC) col[rr++] += bot[r * 3 + q]; else col[rr++] = bot[r * 3 + q]; }
        }
    }
}

// spin until the phase of the "image may be overwritten" mbarrier with this parity has completed
__device__ __forceinline__ void t2_wait_image(unsigned long long *bar, int parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "T2_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra T2_WAIT_%=;\n"
        "}\n" ::"r"(t2_saddr(bar)), "r"(parity)
        : "memory");
}

// the DKT staging columns (skb) are not needed when the blocks come from the class table
#define CB_T2_SKB(CLS) ((CLS) ? 0 : 18 * CB_T2_T)
#define CB_T2_SMEM_DOUBLES(CLS) (CB_T2_OUT + 2 + 2 * CB_T2_ELEMS * CB_SH_KREC + CB_T2_SKB(CLS))
#define CB_T2_SMEM_BYTES(CLS) (CB_T2_SMEM_DOUBLES(CLS) * 8 + 2 * CB_T2_T * 16 + CB_T2_T * 16 + CB_T2_EIDS * CB_T2_T * 4 + 4 * 48 + 16)

// CLS: the DKT sub-blocks come from the geometry-class table (L1-resident) instead of the
// work-ordered per-contribution copy in HBM; the classes of a work item's two contributions are
// packed in its c0 field.
template <bool CLS>
__global__ void __launch_bounds__(CB_T2_T, CLS ? CB_T2_CTAS_CLS : CB_T2_CTAS)
k_assemble_shell_tiles(CbStiffArgs A)
{
    extern __shared__ __align__(16) double smem[];
    double *obuf = smem;                                              // [CB_T2_OUT + 2]
    double *skrec = obuf + CB_T2_OUT + 2;                             // [2][CB_T2_ELEMS*18], 16 B aligned
    double *skb = skrec + 2 * CB_T2_ELEMS * CB_SH_KREC;               // [18][CB_T2_T]
    CbTPair *spair2 = reinterpret_cast<CbTPair *>(skb + CB_T2_SKB(CLS));   // [2][CB_T2_T]
    int4 *swork = reinterpret_cast<int4 *>(spair2 + 2 * CB_T2_T);   // [CB_T2_T]
    int *seid = reinterpret_cast<int *>(swork + CB_T2_T);           // [CB_T2_EIDS][CB_T2_T]
    int *sring = seid + CB_T2_EIDS * CB_T2_T;                       // [4][12] tile records
    // mbarrier "the image may be overwritten": thread 0 arrives once the copy engine has read the
    // previous tile's image out; every thread checks it right before its first image store, by which
    // time (half a tile later) the phase has long completed - so nobody waits for the read-out at the
    // CTA barrier at the top of the loop any more
    unsigned long long *sbar = reinterpret_cast<unsigned long long *>(sring + 48);
    const int t = threadIdx.x;
    if (CB_T2_MBAR && t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t2_saddr(sbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const long G = gridDim.x, N = A.ntiles2;
    long tile = blockIdx.x;
    if (tile >= N) return;
    // Software pipeline over the CTA's tiles k, k+G, ...  Everything the loop reads from global
    // memory arrives through cp.async, issued half a tile or more before it is needed:
    //   tile record three ahead (ring of four in shared memory); element ids two ahead; the work
    //   item, the shell + pair records and the DKT entries of the left halves of the next tile in
    //   one group, committed after this tile's left halves; the DKT entries of its right halves in
    //   a second group after this tile's right halves.  At most the newest group is in flight at
    //   the two points where data is needed (cp.async.wait_group 1).
    CbTile2 tl = A.tiles2[tile];
    int it = 0;                                                        // iteration count = ring position
    {
        if (t < 20) {                                                  // records k+1, k+2 into the ring
            const int r = 1 + t / 10, wd = t % 10;
            if (tile + r * G < N) sring[r * 12 + wd] = reinterpret_cast<const int *>(A.tiles2 + tile + r * G)[wd];
        }
        int eid[CB_T2_EIDS];
#pragma unroll
        for (int k = 0; k < CB_T2_EIDS; ++k) {
            const int i = t + k * CB_T2_T;
            eid[k] = (i < tl.ne * 9) ? A.tile_elems[tl.e0 + i / 9] : 0;
        }
        t2_issue_stage(A, tl, eid, skrec, spair2);
        if (tile + G < N) t2_issue_eids(A, A.tiles2[tile + G], seid);
        if (t < tl.nw) {
            CB_CPA(16, "cg", swork + t, A.works + tl.w0 + t);
            if (!CLS) t2_issue_kb<true>(A, tile, t, skb);
        }
        CB_CPA_COMMIT();
        if (!CLS && t < tl.nw) t2_issue_kb<false>(A, tile, t, skb);
        CB_CPA_COMMIT();
    }
    int buf = 0;

    for (;;) {
        const long next = tile + G;
        const bool has_next = next < N, has_next2 = next + G < N, has_next3 = next + 2 * G < N;
        // the previous tile's image must have been read out by the copy engine
        if (!CB_T2_MBAR && t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();            // this tile's shell + pair records visible; image + buffers free
        if (CB_T2_MBAR && t == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t2_saddr(sbar)) : "memory");
        }
        CbWork w;
        *reinterpret_cast<int4 *>(&w) = swork[t];
        CbTile2 tln = tl, tlnn = tl;
        // ring slots are 48 bytes, 16-byte aligned: three 16-byte broadcast loads per record instead of
        // ten 4-byte ones (the kernel is bound by shared-memory wavefronts)
        if (has_next) {
            const int4 *r4 = reinterpret_cast<const int4 *>(sring + ((it + 1) & 3) * 12);
            const int4 q0 = r4[0], q1 = r4[1], q2 = r4[2];
            const int v[10] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y};
#pragma unroll
            for (int i = 0; i < 10; ++i) reinterpret_cast<int *>(&tln)[i] = v[i];
        }
        if (has_next2) {
            const int4 *r4 = reinterpret_cast<const int4 *>(sring + ((it + 2) & 3) * 12);
            const int4 q0 = r4[0], q1 = r4[1], q2 = r4[2];
            const int v[10] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y};
#pragma unroll
            for (int i = 0; i < 10; ++i) reinterpret_cast<int *>(&tlnn)[i] = v[i];
        }
        if (has_next3 && t < 10)
            CB_CPA(4, "ca", sring + ((it + 3) & 3) * 12 + t, reinterpret_cast<const int *>(A.tiles2 + next + 2 * G) + t);
        const int shift = (int)((tl.out0 + A.out_par) & 1);
        const CbTPair *spair = spair2 + buf * CB_T2_T;
        const double *kr0 = skrec + buf * CB_T2_ELEMS * CB_SH_KREC;
        const bool active = t < tl.nw;
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        CbTPair pr{};
        int gmax = 0;
        if (active) {
            if (w.kind != 4) pr = spair[w.dst];
            gmax = __reduce_max_sync(amask, w.kind == 2 ? (int)w.pad0 : 0);   // largest group of the warp
        }

        // ---- phase 1: the block (or partial sum) of this thread, three columns at a time --------
#define CB_T2_HALF(LEFT, C0)                                                                       \
        if (active) {                                                                              \
            double top[9], bot[9];                                                                 \
            if (w.n >= 1) {                                                                        \
                double kb[18];                                                                     \
                if (CLS) {                                                                         \
                    const double *k0 = A.d.keb_tab + (w.c0 & 0xffff) * 81 + (3 * w.a0 + w.b0) * 9;  \
                    const double *k1 = A.d.keb_tab + ((unsigned)w.c0 >> 16) * 81 + (3 * w.a1 + w.b1) * 9; \
                    _Pragma("unroll") for (int i = 0; i < 9; ++i)                                  \
                        if (((i % 3) == 0) == LEFT) {                                              \
                            kb[i] = __ldg(k0 + i);                                                 \
                            kb[9 + i] = (w.n == 2) ? __ldg(k1 + i) : 0.0;                          \
                        }                                                                          \
                } else {                                                                           \
                _Pragma("unroll") for (int i = 0; i < 18; ++i)                                     \
                    if (((i % 3) == 0) == LEFT) kb[i] = skb[i * CB_T2_T + t];                    \
                }                                                                                  \
                shell_half_acc<LEFT>(kr0 + w.s0 * CB_SH_KREC, kb, w.a0, w.b0, top, bot, true);     \
                if (w.n == 2)                                                                      \
                    shell_half_acc<LEFT>(kr0 + w.s1 * CB_SH_KREC, kb + 9, w.a1, w.b1, top, bot, false); \
                if (CB_T2_MBAR && LEFT) t2_wait_image(sbar, it & 1);                               \
                if (w.kind != 3) t2_store_half<false>(obuf, shift, pr, C0, top, bot);              \
            }                                                                                      \
            for (int q = 1; q < gmax; ++q) {        /* warp-uniform */                             \
                __syncwarp(amask);                                                                 \
                if (w.kind == 3 && w.pad1 == q) t2_store_half<true>(obuf, shift, pr, C0, top, bot); \
            }                                                                                      \
        }
        CB_T2_HALF(true, 0)
        // group 1 for the next tile (its left DKT entries go where this thread's were)
        if (has_next) {
            int eid[CB_T2_EIDS];
#pragma unroll
            for (int k = 0; k < CB_T2_EIDS; ++k) eid[k] = seid[k * CB_T2_T + t];
            t2_issue_stage(A, tln, eid, skrec + (buf ^ 1) * CB_T2_ELEMS * CB_SH_KREC,
                           spair2 + (buf ^ 1) * CB_T2_T);
            if (has_next2) t2_issue_eids(A, tlnn, seid);
            if (t < tln.nw) {
                CB_CPA(16, "cg", swork + t, A.works + tln.w0 + t);
                if (!CLS) t2_issue_kb<true>(A, next, t, skb);
            }
        }
        CB_CPA_COMMIT();
        asm volatile("cp.async.wait_group 1;" ::: "memory");     // this tile's right DKT entries
        CB_T2_HALF(false, 3)
#undef CB_T2_HALF
        if (!CLS && has_next && t < tln.nw) t2_issue_kb<false>(A, next, t, skb);
        CB_CPA_COMMIT();
        // writes of the image (generic proxy) ordered before the copy engine's reads (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        // ---- phase 2: the image leaves as one bulk copy (16-byte aligned middle part) -----------
        if (t == 0) {
            double *dst = A.out + tl.out0;
            const double *img = obuf + shift;
            const int nv = (tl.nout - shift) >> 1;
            if (nv > 0)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + shift),
                             "r"(t2_saddr(img + shift)), "r"(nv * 16)
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (shift) dst[0] = img[0];
            const int tail = shift + 2 * nv;
            if (tail < tl.nout) dst[tail] = img[tail];
        }
        if (!has_next) break;
        tile = next; tl = tln; buf ^= 1; ++it;
    }
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <bool CLS>
static int launch_shell_tiles(const CbStiffArgs &a, cudaStream_t s)
{
    const size_t smem = CB_T2_SMEM_BYTES(CLS);
    static CbPerDevice cache{};                      // resident-CTA grid per device (0: not configured)
    int &grid_cache = cache.v[cb_device_slot()];
    if (!grid_cache) {
        if (cudaFuncSetAttribute(k_assemble_shell_tiles<CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return 1;
        int per_sm = 0, dev = 0, nsm = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_shell_tiles<CLS>, CB_T2_T, smem) !=
                cudaSuccess || per_sm < 1)
            per_sm = 1;
        grid_cache = per_sm * nsm;
    }
    long grid = grid_cache;
    if (grid > a.ntiles2) grid = a.ntiles2;
    k_assemble_shell_tiles<CLS><<<(unsigned)grid, CB_T2_T, smem, s>>>(a);
    return cudaGetLastError() != cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// CSC, shell-only models: "stream" tile kernel (plan: cb_internal.h, CbStreamShape / CbTileS).  One
// persistent WARP walks tiles (runs of consecutive joints whose CSC columns are one contiguous slice of
// Ax); the warps of a CTA share nothing, so there is no CTA barrier anywhere.  Per tile:
//   * the records of the tile's distinct shells (krec, 144 B each), its step and pair records arrive by
//     cp.async, double-buffered one tile ahead; element ids and the tile record are register-prefetched
//     (their addresses depend on the tile number only);
//   * every lane walks its steps: one contribution per step, the shell record read ONCE from shared
//     memory (9 x 16 bytes), the DKT sub-block of the NEXT step already on its way, the full 6x6
//     accumulated in registers (FMA chains onto the running sum of the joint-pair block), the block
//     written into the tile image ONCE (18 x 16-byte stores) when its last contribution has been added.
//     The planner hands whole blocks to lanes, so no partial sums meet in shared memory - except the
//     second part of a block too long for one lane, which is added once at the end of the tile;
//   * the finished image leaves as one bulk asynchronous copy shared -> global (TMA engine).
// shared memory per warp: image | shell records x2 | step records x2 | pair records x2
// ------------------------------------------------------------------------------------------
#ifndef CB_S_PREK
#define CB_S_PREK 1               // the next step's shell record is read into registers while this step is evaluated
#endif
#ifndef CB_S_NARROW_WARPS
#define CB_S_NARROW_WARPS 8
#endif
template <int IMG, int SLOTS, int S, int PAIRS, int NBUF = 2>
struct SLayout {
    static constexpr int IMG_BYTES = (IMG + 2) * 8;
    static constexpr int KREC_BYTES = NBUF * SLOTS * CB_SH_KREC * 8;
    static constexpr int REC_BYTES = NBUF * S * 32 * 4;
    static constexpr int PAIR_BYTES = NBUF * PAIRS * 4;
    static constexpr int AUX_BYTES = NBUF == 1 ? 144 : 0;        // next tile's element ids (32 x 4) + tile record (16)
    static constexpr int WARP_BYTES = IMG_BYTES + KREC_BYTES + REC_BYTES + PAIR_BYTES + AUX_BYTES;
    static_assert(WARP_BYTES % 16 == 0 && IMG_BYTES % 16 == 0 && PAIRS % 4 == 0, "16-byte aligned regions");
};

// loads that must be ISSUED where they are written (their results are used an iteration later; a plain
// load would be sunk to its first use and expose the full DRAM latency there)
__device__ __forceinline__ int4 s_ldg16(const void *p)
{
    int4 v;
    asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int s_ldg4(const void *p)
{
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <int SLOTS>
struct SEids { int v[(SLOTS + 31) / 32]; };
template <int SLOTS>
__device__ __forceinline__ SEids<SLOTS> s_load_eids(const CbStiffArgs &A, long tile, int lane)
{
    SEids<SLOTS> e;
#pragma unroll
    for (int h = 0; h < (SLOTS + 31) / 32; ++h)
        e.v[h] = (lane + 32 * h < SLOTS) ? s_ldg4(A.elemsS + tile * SLOTS + lane + 32 * h) : 0;
    return e;
}

// stage one tile: the shell records of the tile's slots are SLOTS x 9 chunks of 16 bytes, chunk q copied by lane
// q % 32 - consecutive lanes read the consecutive bytes of one 144-byte record and write consecutive shared
// memory (one lane per SLOT made every request 28 separate sectors and the shared-memory write of each
// returning sector its own wavefront: 44 % of the kernel's shared-memory wavefronts); the element index of
// a slot is held by lane slot % 32 (s_load_eids) and fetched by shuffle.  Slots past the tile's last shell
// repeat a valid one.  Then 16-byte chunks of the step records and of the pair records.
template <int SLOTS, int S, int PAIRS>
__device__ __forceinline__ void s_issue_stage(const CbStiffArgs &A, long tile, const SEids<SLOTS> &eid, int lane,
                                              double *krec_dst, uint32_t *rec_dst, uint32_t *pair_dst)
{
    constexpr int NCH = SLOTS * 9;
#pragma unroll
    for (int k = 0; k < (NCH + 31) / 32; ++k) {
        const int q = lane + 32 * k;
        const int slot = min(q / 9, SLOTS - 1), ch = q - slot * 9;
        int e = __shfl_sync(0xffffffffu, eid.v[0], slot & 31);
        if constexpr (SLOTS > 32) {
            const int e1 = __shfl_sync(0xffffffffu, eid.v[1], slot & 31);
            if (slot >= 32) e = e1;
        }
        if (q < NCH) CB_CPA(16, "cg", krec_dst + q * 2, A.d.sh_Nm + (long)e * CB_SH_KREC + ch * 2);
    }
    const uint32_t *rs = A.stepsS + tile * (S * 32);
#pragma unroll
    for (int k = 0; k < (S * 8 + 31) / 32; ++k) {
        const int c = lane + 32 * k;
        if (c < S * 8) CB_CPA(16, "cg", rec_dst + c * 4, rs + c * 4);
    }
    if (lane < PAIRS / 4) CB_CPA(16, "cg", pair_dst + lane * 4, A.pairsS + tile * PAIRS + lane * 4);
}

// n == 0 ? v0 : (n == 1 ? v1 : v2) as two selp (the lanes of a warp hold different local joints: the
// compiler's own lowering of the ternaries is a divergent branch)
__device__ __forceinline__ double sel3(int n, double v0, double v1, double v2)
{
    double r;
    asm("{\n"
        ".reg .pred p0, p1;\n"
        ".reg .f64 t;\n"
        "setp.eq.s32 p0, %1, 0;\n"
        "setp.eq.s32 p1, %1, 1;\n"
        "selp.f64 t, %3, %4, p1;\n"
        "selp.f64 %0, %2, t, p0;\n"
        "}\n" : "=d"(r) : "r"(n), "d"(v0), "d"(v1), "d"(v2));
    return r;
}

// acc (column-major 6x6: acc[q * 6 + r] = K[r][q]) += R^T-rotated local block (a, b) of one shell
__device__ __forceinline__ void s_load_krec(const double *kr /*shared*/, double *k)
{
    const double2 *k2 = reinterpret_cast<const double2 *>(kr);
#pragma unroll
    for (int i = 0; i < CB_SH_KREC / 2; ++i) { const double2 v = k2[i]; k[2 * i] = v.x; k[2 * i + 1] = v.y; }
}
__device__ __forceinline__ void s_contrib_core(const double *R, const double *kb, double s00, double s01, double s10,
                                               double s11, double s22, double drill, double *acc)
{
    double W0[3], W1[3], W2[3], U0[3], U1[3], U2[3], w[3], v[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        W0[q] = s00 * R[q] + s01 * R[3 + q];
        W1[q] = s10 * R[q] + s11 * R[3 + q];
        W2[q] = s22 * R[6 + q];
        U0[q] = kb[4] * R[q] + kb[5] * R[3 + q];
        U1[q] = kb[7] * R[q] + kb[8] * R[3 + q];
        U2[q] = drill * R[6 + q];
        w[q] = kb[1] * R[q] + kb[2] * R[3 + q];          // translation-rotation: e3 (x) (k01 e1 + k02 e2)
        v[q] = kb[3] * R[q] + kb[6] * R[3 + q];          // rotation-translation: (k10 e1 + k20 e2) (x) e3
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            // three-FMA chains onto the running sum of the block (no separate add)
            acc[q * 6 + p] = fma(R[6 + p], W2[q], fma(R[3 + p], W1[q], fma(R[p], W0[q], acc[q * 6 + p])));
            acc[q * 6 + 3 + p] = fma(v[p], R[6 + q], acc[q * 6 + 3 + p]);
            acc[(3 + q) * 6 + p] = fma(R[6 + p], w[q], acc[(3 + q) * 6 + p]);
            acc[(3 + q) * 6 + 3 + p] =
                fma(R[6 + p], U2[q], fma(R[3 + p], U1[q], fma(R[p], U0[q], acc[(3 + q) * 6 + 3 + p])));
        }
}
__device__ __forceinline__ void s_contrib(const double *k /*registers*/, const double *kb, int a, int b, double *acc)
{
    // CST shape-function gradients (x 2A) of local joints a, b: (-Y3, X3 - X2), (Y3, -X3), (0, X2)
    const double nY3 = -k[11], dX = k[10] - k[9], nX3 = -k[10];
    const double bxa = sel3(a, nY3, k[11], 0.0), bya = sel3(a, dX, nX3, k[9]);
    const double bxb = sel3(b, nY3, k[11], 0.0), byb = sel3(b, dX, nX3, k[9]);
    const double g = bxa * (k[15] * bxb + k[17] * byb) + bya * (k[17] * bxb + k[16] * byb);
    const double s00 = k[12] * bxa * bxb + k[14] * bya * byb + g;
    const double s01 = k[13] * bxa * byb + k[14] * bya * bxb;
    const double s10 = k[13] * bya * bxb + k[14] * bxa * byb;
    const double s11 = k[12] * bya * byb + k[14] * bxa * bxb + g;
    const double s22 = kb[0] + g;
    const double drill = sel3(a == b ? 0 : 1, kb[4] * 1e-4, 0.0, 0.0);       // shell.c:482-484 (k / 1e4)
    s_contrib_core(k, kb, s00, s01, s10, s11, s22, drill, acc);
}
// geometry classes: everything of the local block (a, b) that does not depend on the shell's state comes from the
// class row (CB_KROW doubles, k_class_tables): [0..8] the DKT 3x3 block, [9] the drilling term, [10..13] the
// material membrane block, [14..16] the gradient products that the membrane forces n0, n1, n2 (the only state
// besides the triad) multiply in the geometric stiffness - no selects on (a, b), no gradients, 14 record doubles
__device__ __forceinline__ void s_contrib_cls(const double *k /*registers: R[9] at 0, n at 15..17*/, const double *kb, double *acc)
{
    const double g = k[15] * kb[14] + k[16] * kb[15] + k[17] * kb[16];
    s_contrib_core(k, kb, kb[10] + g, kb[11], kb[12], kb[13] + g, kb[0] + g, kb[9], acc);
}
// the 14 doubles s_contrib_cls reads of a shell record: 16-byte units 0-4 (triad; k[9] rides along) and 7, 8
__device__ __forceinline__ void s_load_krec_cls(const double *kr /*shared*/, double *k)
{
    const double2 *k2 = reinterpret_cast<const double2 *>(kr);
#pragma unroll
    for (int i = 0; i < CB_SH_KREC / 2; ++i)
        if (i < 5 || i > 6) { const double2 v = k2[i]; k[2 * i] = v.x; k[2 * i + 1] = v.y; }
}

// the block held in acc goes to (ADD: is added to) its place in the tile image
template <bool ADD>
__device__ __forceinline__ void s_store_block(double *obuf, int shift, uint32_t pr, const double *acc)
{
    const int rel = pr & 0xfff, colh = (pr >> 12) & 0xff;
    const unsigned maskA = (pr >> 20) & 0x3f, maskB = pr >> 26;
    double *img = obuf + shift + rel;
    if (maskA == 0x3f && maskB == 0x3f && ((((shift + rel) | colh) & 1) == 0)) {
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            double2 *col = reinterpret_cast<double2 *>(img + q * colh);
            double2 v0 = make_double2(acc[q * 6], acc[q * 6 + 1]);
            double2 v1 = make_double2(acc[q * 6 + 2], acc[q * 6 + 3]);
            double2 v2 = make_double2(acc[q * 6 + 4], acc[q * 6 + 5]);
            if (ADD) {
                const double2 o0 = col[0], o1 = col[1], o2 = col[2];
                v0.x += o0.x; v0.y += o0.y; v1.x += o1.x; v1.y += o1.y; v2.x += o2.x; v2.y += o2.y;
            }
            col[0] = v0; col[1] = v1; col[2] = v2;
        }
    } else {
        int cc = 0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            if (!((maskB >> q) & 1)) continue;
            double *col = img + cc * colh;
            int rr = 0;
#pragma unroll
            for (int p = 0; p < 6; ++p)
                if ((maskA >> p) & 1) { if (ADD) col[rr] += acc[q * 6 + p]; else col[rr] = acc[q * 6 + p]; ++rr; }
            ++cc;
        }
    }
}

// registers: one CTA per SM of at most 8 warps, so every thread may use the full 255 (the register file is
// handed out to warps in groups of four: 10 warps would be budgeted like 12, 168 registers, and spill)
// NBUF 1: one set of record buffers per warp - the next tile's records are requested when this tile's last
// step has read them (their latency is hidden by the other warps, not by the tile's own arithmetic)
template <int IMG, int SLOTS, int S, int PAIRS, int WARPS, bool CLS, int NBUF = 2>
__global__ void __launch_bounds__(32 * WARPS, 1)
k_assemble_shell_stream(CbStiffArgs A)
{
    using L = SLayout<IMG, SLOTS, S, PAIRS, NBUF>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wbase = smem_raw + (size_t)warp * L::WARP_BYTES;
    double *obuf = reinterpret_cast<double *>(wbase);                                           // [IMG + 2]
    double *skrec = reinterpret_cast<double *>(wbase + L::IMG_BYTES);                           // [2][SLOTS][18]
    uint32_t *srec = reinterpret_cast<uint32_t *>(wbase + L::IMG_BYTES + L::KREC_BYTES);         // [2][S][32]
    uint32_t *spair = srec + NBUF * S * 32;                                                      // [2][PAIRS]
    int *seid = reinterpret_cast<int *>(spair + NBUF * PAIRS);                                   // [32]   (NBUF 1)
    int4 *stlr = reinterpret_cast<int4 *>(seid + 32);                                            // [1]    (NBUF 1)
    const long G = (long)gridDim.x * WARPS, N = A.ntilesS;
    long tile = (long)blockIdx.x * WARPS + warp;
    if (tile >= N) return;
    constexpr unsigned FULL = 0xffffffffu;
    // more than 8 warps per SM leave 168 registers per thread: the register prefetches (next step's shell record,
    // next tile's first DKT block) go, the extra warps hide those latencies instead
    constexpr bool LEAN = WARPS > 8;
    constexpr bool PREK = CB_S_PREK && !LEAN;
#ifndef CB_S_LEAN_KBPRE
#define CB_S_LEAN_KBPRE 0         // lean variants: 1 = request the next step's DKT block a step ahead (20 registers; measured 1 % slower)
#endif
    constexpr bool KBPRE = !LEAN || CB_S_LEAN_KBPRE;
#ifndef CB_S_CLSROW
#define CB_S_CLSROW 1             // lean class variants: the state-independent part of a block comes from the class row
#endif
    constexpr bool CLSROW = LEAN && CLS && CB_S_CLSROW && !KBPRE;
    static_assert(NBUF == 2 || LEAN, "single record buffers go with the lean variant");
    static_assert(NBUF == 2 || SLOTS <= 32, "one element id per lane");

    // prologue: this tile staged directly; the next tile's record and element ids into registers - or, with
    // single buffers (168 registers: prefetched registers would be spilled, and the spill store WAITS for the
    // load), into shared memory by cp.async: tile record with the tile's stage, element ids one tile ahead
    int4 tlr, tlr_n;
    SEids<SLOTS> eid_n{};
    if constexpr (NBUF == 1) {
        const SEids<SLOTS> eid = s_load_eids<SLOTS>(A, tile, lane);
        s_issue_stage<SLOTS, S, PAIRS>(A, tile, eid, lane, skrec, srec, spair);
        if (lane == 0) CB_CPA(16, "cg", stlr, A.tilesS + tile);
        CB_CPA_COMMIT();
        tlr = tlr_n = make_int4(0, 0, 0, 0);
    } else {
        tlr = s_ldg16(A.tilesS + tile); tlr_n = tlr;
        const SEids<SLOTS> eid = s_load_eids<SLOTS>(A, tile, lane);
        s_issue_stage<SLOTS, S, PAIRS>(A, tile, eid, lane, skrec, srec, spair);
        CB_CPA_COMMIT();
        if (tile + G < N) { tlr_n = s_ldg16(A.tilesS + tile + G); eid_n = s_load_eids<SLOTS>(A, tile + G, lane); }
    }
    int buf = 0;
    bool img_busy = false;            // a bulk read-out of the image may still be in flight
    // first tile: wait for its records here and fetch step 0's record and DKT block the slow way
    uint32_t r0n;
    double kbC[10], kbN[10];
    {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        r0n = srec[lane];
    }
    if constexpr (!LEAN) {
        if (CLS) {
            const double *k0 = A.d.keb_tab10 + ((r0n >> 20) * 9 + 3 * ((r0n >> 6) & 3) + ((r0n >> 8) & 3)) * 10;
#pragma unroll
            for (int i = 0; i < 10; ++i) kbC[i] = __ldg(k0 + i);
        } else {
            const double *k0 = A.kebc + (tile * (S * 9L)) * 32 + lane;
#pragma unroll
            for (int i = 0; i < 9; ++i) kbC[i] = __ldg(k0 + (long)i * 32);
            kbC[9] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) kbN[i] = kbC[i];
    }

    double acc[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) acc[i] = 0.0;

    for (;;) {
        const bool has_next = tile + G < N, has_next2 = tile + 2 * G < N;
        // the other buffer was last read by the previous tile (the warp re-converged at its end)
        if (NBUF == 2) {
            if (has_next)
                s_issue_stage<SLOTS, S, PAIRS>(A, tile + G, eid_n, lane, skrec + (buf ^ 1) * SLOTS * CB_SH_KREC,
                                               srec + (buf ^ 1) * S * 32, spair + (buf ^ 1) * PAIRS);
            CB_CPA_COMMIT();
        }
        SEids<SLOTS> eid_nn{};
        int4 tlr_nn = tlr_n;
        if constexpr (NBUF == 2) {
            if (has_next2) { eid_nn = s_load_eids<SLOTS>(A, tile + 2 * G, lane); tlr_nn = s_ldg16(A.tilesS + tile + 2 * G); }
        } else {
            if (has_next && lane < SLOTS) CB_CPA(4, "ca", seid + lane, A.elemsS + (tile + G) * SLOTS + lane);
            CB_CPA_COMMIT();
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");       // this tile's records have landed
        __syncwarp();
        if constexpr (NBUF == 1) tlr = *stlr;

        const long out0 = ((long)(unsigned)tlr.x) | ((long)tlr.y << 32);
        const int nout = tlr.z, nsteps = tlr.w & 0xff;
        const double *kr0 = skrec + buf * SLOTS * CB_SH_KREC;
        const uint32_t *rec = srec + buf * S * 32 + lane;
        const uint32_t *pairs = spair + buf * PAIRS;
        const int shift = (int)((out0 + A.out_par) & 1);
        const double *kbs = CLS ? nullptr : A.kebc + (tile * (S * 9L)) * 32 + lane;
        if (!LEAN && !CLS && has_next) {       // step 0 of the next tile: its address depends on the tile number only
            const double *kn0 = A.kebc + ((tile + G) * (S * 9L)) * 32 + lane;
#pragma unroll
            for (int i = 0; i < 9; ++i)
                asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(kbN[i]) : "l"(kn0 + (long)i * 32));
        }

        // DKT sub-block of a step, requested one step ahead of its use: class table (padded to ten
        // doubles per 3x3 block: five 16-byte loads) or the step-ordered copy in HBM (nine coalesced
        // 8-byte loads).  Idle lanes (slot 63: a = b = class 0) read valid memory and compute a block
        // that is never stored - the step has no divergent branch around its arithmetic.
        auto load_kb = [&](uint32_t r, int st, double *kb) {
            if (CLSROW) {
                const double *k0 = A.d.keb_row + ((r >> 20) * 9 + 3 * ((r >> 6) & 3) + ((r >> 8) & 3)) * CB_KROW;
#pragma unroll
                for (int i = 0; i < CB_KROW / 2; ++i)
                    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(kb[2 * i]), "=d"(kb[2 * i + 1]) : "l"(k0 + 2 * i));
            } else if (CLS) {
                const double *k0 = A.d.keb_tab10 + ((r >> 20) * 9 + 3 * ((r >> 6) & 3) + ((r >> 8) & 3)) * 10;
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(kb[2 * i]), "=d"(kb[2 * i + 1]) : "l"(k0 + 2 * i));
            } else {
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(kb[i]) : "l"(kbs + ((long)st * 9 + i) * 32));
            }
        };
        // step 0's record and DKT block were requested at the end of the previous tile (r0n / kbC); the
        // other step records of the tile go to registers now (rows past nsteps are idle)
        uint32_t rs[S];
        rs[0] = LEAN ? rec[0] : r0n;
#pragma unroll
        for (int st = 1; st < S; ++st) rs[st] = rec[st * 32];
        double kbA[CB_KROW], kbB[CB_KROW], krA[CB_SH_KREC], krB[CB_SH_KREC];
        if constexpr (LEAN) {
            load_kb(rs[0], 0, kbA);
        } else {
#pragma unroll
            for (int i = 0; i < 10; ++i) kbA[i] = kbC[i];
        }
        if (PREK && (rs[0] & 63u) != CB_S_IDLE) s_load_krec(kr0 + (rs[0] & 63u) * CB_SH_KREC, krA);
        bool pend = false; uint32_t pend_pr = 0;
#pragma unroll
        for (int st = 0; st < S; ++st) {
            if (st < nsteps) {                              // warp-uniform
                const uint32_t r = rs[st];
                double *kb = (st & 1) ? kbB : kbA, *kbn = (st & 1) ? kbA : kbB;
                double *kr = (st & 1) ? krB : krA, *krn = (st & 1) ? krA : krB;
                const unsigned slot = r & 63u;
                const bool idle = slot == CB_S_IDLE;
                const bool first = (r >> 19) & 1u;
                const bool last = !idle && ((r >> 10) & 1u);
                const bool follow = last && ((r >> 11) & 1u);
                uint32_t pr = 0;
                if (last) pr = pairs[(r >> 12) & 127u];     // needed at the end of the step
                if (st + 1 < S) {                           // the next step's inputs are on their way while
                    const uint32_t rn = rs[st + 1];         // this one is evaluated
                    if ((rn & 63u) != CB_S_IDLE) {
                        if (KBPRE) load_kb(rn, st + 1, kbn);
                        if (PREK) s_load_krec(kr0 + (rn & 63u) * CB_SH_KREC, krn);
                    }
                }
                if (!LEAN && st + 1 == nsteps && has_next) {
                    // last step: the next tile's records have landed long ago - request its first DKT block
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                    r0n = srec[(buf ^ 1) * S * 32 + lane];
                    if (CLS && (r0n & 63u) != CB_S_IDLE) load_kb(r0n, 0, kbC);
                }
                // a lane that starts a block forgets the previous one (its store has long read the registers)
                if (__any_sync(FULL, first)) {
                    if (first) {
#pragma unroll
                        for (int i = 0; i < 36; ++i) acc[i] = 0.0;
                    }
                }
                if (!idle && CLSROW) {
                    if (st > 0) load_kb(r, st, kb);
                    s_load_krec_cls(kr0 + slot * CB_SH_KREC, kr);
                    s_contrib_cls(kr, kb, acc);
                }
                if (!idle && !CLSROW) {
                    if (!KBPRE && st > 0) load_kb(r, st, kb);
                    if (!PREK) s_load_krec(kr0 + slot * CB_SH_KREC, kr);
                    s_contrib(kr, kb, (r >> 6) & 3, (r >> 8) & 3, acc);
                }
                if (follow) { pend = true; pend_pr = pr; }                  // held until the tile's end
                if (__any_sync(FULL, last && !follow)) {
                    if (img_busy) {        // the copy engine must have read the previous image out
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        __syncwarp();
                        img_busy = false;
                    }
                    if (last && !follow) s_store_block<false>(obuf, shift, pr, acc);
                }
            }
        }
        if (__any_sync(FULL, pend)) {          // second parts of the blocks that were cut in two
            __syncwarp();                      // their first parts are in the image
            if (pend) s_store_block<true>(obuf, shift, pend_pr, acc);
        }
        if constexpr (NBUF == 1) {             // every lane has read its last record: request the next tile's
            asm volatile("cp.async.wait_group 0;" ::: "memory");       // its element ids arrived a tile ago
            __syncwarp();
            if (has_next) {
                SEids<SLOTS> e;
                e.v[0] = lane < SLOTS ? seid[lane] : 0;
                s_issue_stage<SLOTS, S, PAIRS>(A, tile + G, e, lane, skrec, srec, spair);
                if (lane == 0) CB_CPA(16, "cg", stlr, A.tilesS + tile + G);
            }
            CB_CPA_COMMIT();
        }
        // writes of the image (generic proxy) ordered before the copy engine's reads (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            double *dst = A.out + out0;
            const double *im = obuf + shift;
            const int nv = (nout - shift) >> 1;
            if (nv > 0)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + shift),
                             "r"(t2_saddr(im + shift)), "r"(nv * 16)
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (shift) dst[0] = im[0];
            const int tail = shift + 2 * nv;
            if (tail < nout) dst[tail] = im[tail];
        }
        img_busy = true;
        if (!has_next) break;
        if (!LEAN && !CLS) {
#pragma unroll
            for (int i = 0; i < 9; ++i) kbC[i] = kbN[i];
        }
        tile += G;
        if constexpr (NBUF == 2) { tlr = tlr_n; tlr_n = tlr_nn; eid_n = eid_nn; buf ^= 1; }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int IMG, int SLOTS, int S, int PAIRS, int WARPS, bool CLS, int NBUF = 2>
static int launch_shell_stream_t(const CbStiffArgs &a, cudaStream_t s)
{
    const size_t smem = (size_t)WARPS * SLayout<IMG, SLOTS, S, PAIRS, NBUF>::WARP_BYTES;
    static CbPerDevice cache{};                      // SM count per device (0: not configured)
    int &nsm = cache.v[cb_device_slot()];
    if (!nsm) {
        const cudaError_t e = cudaFuncSetAttribute(k_assemble_shell_stream<IMG, SLOTS, S, PAIRS, WARPS, CLS, NBUF>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int dev = 0, n = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        nsm = n;
    }
    long grid = nsm;                                 // persistent: one CTA of WARPS independent warps per SM
    const long need = (a.ntilesS + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    k_assemble_shell_stream<IMG, SLOTS, S, PAIRS, WARPS, CLS, NBUF><<<(unsigned)grid, 32 * WARPS, smem, s>>>(a);
    return (int)cudaGetLastError();
}

// the two compiled shapes of cb_internal.h (CB_S_SHAPE_WIDE / CB_S_SHAPE_NARROW)
static int launch_shell_stream(const CbStiffArgs &a, cudaStream_t s)
{
    const bool cls = a.d.keb_tab != nullptr;
    if (a.shapeS == 0)
        return cls ? launch_shell_stream_t<2560, 44, 6, 80, 6, true>(a, s) : launch_shell_stream_t<2560, 44, 6, 80, 6, false>(a, s);
    if (a.shapeS == 4)                      // experiment: 16 warps at 128 registers
        return cls ? launch_shell_stream_t<1024, 20, 3, 32, 16, true, 1>(a, s) : launch_shell_stream_t<1024, 20, 3, 32, 16, false, 1>(a, s);
    if (a.shapeS == 3)                      // the narrow plan on 12 warps with single record buffers
        return cls ? launch_shell_stream_t<1536, 28, 4, 48, 12, true, 1>(a, s) : launch_shell_stream_t<1536, 28, 4, 48, 12, false, 1>(a, s);
    if (a.shapeS == 2)
        return cls ? launch_shell_stream_t<1024, 20, 3, 32, 12, true>(a, s) : launch_shell_stream_t<1024, 20, 3, 32, 12, false>(a, s);
    return cls ? launch_shell_stream_t<1536, 28, 4, 48, CB_S_NARROW_WARPS, true>(a, s)
               : launch_shell_stream_t<1536, 28, 4, 48, CB_S_NARROW_WARPS, false>(a, s);
}

// ------------------------------------------------------------------------------------------
// skyline: one thread per node-pair block
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
        if (ct.type == CB_T_SHELL) {
            double blk[36];
            shell_block(A.d, ct.e, ct.a, ct.b, blk, 6);
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int q = 0; q < 6; ++q) acc[r * 7 + q] += blk[r * 6 + q];
        } else if (ct.type == CB_T_FRAME) {
            double blk[49];
            frame_block(A, ct.e, ct.a, ct.b, blk, 7);
            for (int i = 0; i < 49; ++i) acc[i] += blk[i];
        } else if (ct.type == CB_T_TRUSS || ct.type == CB_T_BRICK) {
            double blk[9];
            if (ct.type == CB_T_TRUSS) truss_block(A, ct.e, ct.a, ct.b, blk, 3);
            else brick_block(A, ct.e, ct.a, ct.b, blk, 3);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 3; ++q) acc[r * 7 + q] += blk[r * 3 + q];
        }
    }
    if (!A.skyline) {
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
            const long dj = A.maxa[j - 1] - 1, dend = A.maxa[j] - 1;   // column j occupies [dj, dend)
            int rr = 0;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                if (!((pr.maskA >> r) & 1)) continue;
                const long i = pr.eqA0 + rr;
                // a joint-pair block may hold DOF pairs no element couples (e.g. rotations of two
                // joints linked only by a truss): they are zero and can lie outside the profile
                if (i <= j && dj + (j - i) < dend) A.out[dj + (j - i)] = acc[r * 7 + c];
                ++rr;
            }
            ++cc;
        }
    }
}

template <int ND, bool SHELL_ONLY, bool FRAME_SIMPLE = false>
static int launch_tiles(const CbStiffArgs &a, cudaStream_t s)
{
    const size_t smem = (size_t)(((ND * ND * (CB_TILE_T + 1) + 1) & ~1) + a.tile_smem_out) * sizeof(double) +
                        CB_TILE_T * sizeof(CbTPair) + CB_TILE_T;
    static CbPerDevice cfg{};
    int &configured = cfg.v[cb_device_slot()];
    if (!configured) {
        if (cudaFuncSetAttribute(k_assemble_tiles<ND, SHELL_ONLY, FRAME_SIMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024) != cudaSuccess)
            return 1;
        configured = 1;
    }
    if (smem > 200 * 1024) return 1;
    int per_sm = 0, dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_tiles<ND, SHELL_ONLY, FRAME_SIMPLE>, CB_TILE_T, smem) !=
            cudaSuccess || per_sm < 1)
        per_sm = 1;
    long grid = (long)per_sm * nsm;                    // persistent: one wave of resident CTAs
    if (grid > a.ntiles) grid = a.ntiles;
    k_assemble_tiles<ND, SHELL_ONLY, FRAME_SIMPLE><<<(unsigned)grid, CB_TILE_T, smem, s>>>(a);
    return cudaGetLastError() != cudaSuccess;
}

int cbk_stiff(const CbStiffArgs &a, cudaStream_t s, long *launches)
{
    if (a.d.NE_BR && !a.mass_mode) {
        if (!a.br_prep) return 1;
        const long n = a.d.NE_BR * 8;
        k_brick_prep<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, a.br_prep);
        ++*launches;
    }
    if (!a.skyline && a.ntilesS > 0 && a.tilesS) {
        ++*launches;
        return launch_shell_stream(a, s);
    }
    if (!a.skyline && a.ntiles2 > 0 && a.tiles2) {
        ++*launches;
        return a.d.keb_tab ? launch_shell_tiles<true>(a, s) : launch_shell_tiles<false>(a, s);
    }
    if (a.skyline || a.tiles == nullptr) {
        if (a.npairs == 0) return 0;
        unsigned g = (unsigned)((a.npairs + CB_TPB_K - 1) / CB_TPB_K);
        k_assemble_blocks<<<g, CB_TPB_K, 0, s>>>(a);
        ++*launches;
        return cudaGetLastError() != cudaSuccess;
    }
    if (a.ntiles == 0) return 0;
    ++*launches;
    const bool shell_only = a.d.NE_SH && !a.d.NE_TR && !a.d.NE_FR && !a.d.NE_BR && !a.mixed;
    if (a.max_dof <= 3) return launch_tiles<3, false>(a, s);
    if (a.max_dof <= 6) return shell_only ? launch_tiles<6, true>(a, s) : launch_tiles<6, false>(a, s);
    const bool frame_simple = a.d.fr_simple && a.d.NE_FR && !a.d.NE_SH && !a.d.NE_TR && !a.d.NE_BR && !a.mixed &&
                              !a.mass_mode;
    return frame_simple ? launch_tiles<7, false, true>(a, s) : launch_tiles<7, false>(a, s);
}
