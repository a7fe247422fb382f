// cb_forces.cu - co-rotational kinematics update + internal-force recovery + f_int / mass gather.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false.  The internal-force path is a chain of
// differences of nearly equal numbers (current minus reference local coordinates, rigid-body
// parts of the displacement increment cancelling inside ke_b * ddb), so agreement with the
// reference to 1e-12 needs the reference's own rounding sequence, not just its formulas.  Every
// routine here therefore keeps the reference's operation ORDER (left-to-right sums, divide not
// reciprocal-multiply) with separate IEEE multiplies and adds - which is what gcc emits for the
// reference on baseline x86-64 (no FMA unit in the target).  Structure is still exploited: the
// 18x18 transformation matrices are block-diagonal copies of one 3x3 triad, so products with
// their structural zeros (which add +-0 and change nothing) are skipped.
//
// Replaces (per Newton iteration):  updatc misc.c:71-185, forces_tr truss.c:231-378,
// forces_fr frame.c:902-1312, forces_sh shell.c:1593-2400 (ANAFLAG 1, 2), the f_temp scatter
// inlined in each of them, mass_tr/fr/sh truss.c:381, frame.c:1314, shell.c:1505; once per model:
// stiffe_b_sh shell.c:533-658 (DKT bending matrix, geometry-constant).
#include "cb_internal.h"
#include <algorithm>
#include "cb_frame_math.cuh"
#include "cb_frame_def_gen.cuh"

#ifndef CB_TPB
#define CB_TPB 128
#endif
#ifndef CB_FORCES_RECOMPUTE_KEB
#define CB_FORCES_RECOMPUTE_KEB 1 // shells without a geometry class: ke_b * ddb as alpha W (alpha^T ddb) from 11 constants
#endif                            // per shell instead of streaming the cached 81-entry matrix (648 B / shell / iteration)
#ifndef CB_FORCES_STAGE_KEB
#define CB_FORCES_STAGE_KEB 0     // 1: per-element DKT matrix staged by cp.async (see k_shell_forces)
#endif
#ifndef CB_FORCES_MINB
#define CB_FORCES_MINB 4
#endif

__device__ __forceinline__ double dot3(const double *a, const double *b)
{   // misc.c:252-262: dp = 0; dp += a[i]*b[i]
    double dp = 0.0;
    dp += a[0] * b[0];
    dp += a[1] * b[1];
    dp += a[2] * b[2];
    return dp;
}

__device__ __forceinline__ void cross3(const double *a, const double *b, double *c, bool unit)
{   // misc.c:264-282
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
    if (unit) {
        double len = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        c[0] /= len; c[1] /= len; c[2] /= len;
    }
}

// x^3 rounded once (double-double product), standing in for libm pow(x,3) when a length has
// to be re-cubed on the device (mass_* overwrite llength)
__device__ __forceinline__ double cube_rn(double x)
{
    const double p = x * x, pe = __fma_rn(x, x, -p);
    const double q = p * x, qe = __fma_rn(p, x, -q);
    return q + (qe + pe * x);
}

// ------------------------------------------------------------------------------------------
// DKT plate-bending stiffness, Batoz explicit form (shell.c:533-658).  ke_b[9][9] row-major.
// ------------------------------------------------------------------------------------------
template <class Store>
__device__ __forceinline__ void dkt_alpha_T_gen(const double *sc, Store st)
{
    const double X2 = sc[5], X3 = sc[6], Y3 = sc[7];
    const double x23 = X2 - X3;
    const double l12 = sc[8] * sc[8], l23 = sc[9] * sc[9], l31 = sc[10] * sc[10];
    const double p4 = -6 * x23 / l23;
    const double p5 = -6 * X3 / l31;
    const double p6 = 6 * X2 / l12;
    const double t4 = 6 * Y3 / l23;
    const double t5 = -6 * Y3 / l31;
    const double q4 = -3 * x23 * Y3 / l23;
    const double q5 = 3 * X3 * Y3 / l31;
    const double r4 = 3 * (Y3 * Y3) / l23;
    const double r5 = 3 * (Y3 * Y3) / l31;
    // transpose of Batoz' alpha (rows = w, theta_x, theta_y of vertices 1, 2, 3)
    st(0, 0, Y3 * p6);        st(0, 1, -(Y3 * p6));      st(0, 2, Y3 * p5);
    st(0, 3, -(X2 * t5));     st(0, 4, 0);               st(0, 5, x23 * t5);
    st(0, 6, -(X3 * p6) - X2 * p5);  st(0, 7, -x23 * p6);  st(0, 8, x23 * p5 + Y3 * t5);

    st(1, 0, 0);              st(1, 1, 0);               st(1, 2, -(Y3 * q5));
    st(1, 3, x23 + X2 * r5);  st(1, 4, x23);             st(1, 5, x23 * (1 - r5));
    st(1, 6, X2 * q5 + Y3);   st(1, 7, Y3);              st(1, 8, -x23 * q5 + Y3 * (1 - r5));

    st(2, 0, -4 * Y3);        st(2, 1, 2 * Y3);          st(2, 2, Y3 * (2 - r5));
    st(2, 3, -(X2 * q5));     st(2, 4, 0);               st(2, 5, x23 * q5);
    st(2, 6, -4 * x23 + X2 * r5);  st(2, 7, 2 * x23);    st(2, 8, x23 * (2 - r5) + Y3 * q5);

    st(3, 0, -(Y3 * p6));     st(3, 1, Y3 * p6);         st(3, 2, Y3 * p4);
    st(3, 3, 0);              st(3, 4, X2 * t4);         st(3, 5, -(X3 * t4));
    st(3, 6, X3 * p6);        st(3, 7, x23 * p6 + X2 * p4);  st(3, 8, -(X3 * p4) + Y3 * t4);

    st(4, 0, 0);              st(4, 1, 0);               st(4, 2, Y3 * q4);
    st(4, 3, X3);             st(4, 4, X3 + X2 * r4);    st(4, 5, X3 * (1 - r4));
    st(4, 6, -Y3);            st(4, 7, -Y3 + X2 * q4);   st(4, 8, Y3 * (r4 - 1) - X3 * q4);

    st(5, 0, -2 * Y3);        st(5, 1, 4 * Y3);          st(5, 2, Y3 * (r4 - 2));
    st(5, 3, 0);              st(5, 4, -(X2 * q4));      st(5, 5, X3 * q4);
    st(5, 6, 2 * X3);         st(5, 7, -4 * X3 + X2 * r4);  st(5, 8, X3 * (2 - r4) - Y3 * q4);

    st(6, 0, 0);              st(6, 1, 0);               st(6, 2, -(Y3 * (p4 + p5)));
    st(6, 3, X2 * t5);        st(6, 4, -(X2 * t4));      st(6, 5, -x23 * t5 + X3 * t4);
    st(6, 6, X2 * p5);        st(6, 7, -(X2 * p4));
    st(6, 8, -x23 * p5 + X3 * p4 - Y3 * (t4 + t5));

    st(7, 0, 0);              st(7, 1, 0);               st(7, 2, Y3 * (q4 - q5));
    st(7, 3, X2 * (r5 - 1));  st(7, 4, X2 * (r4 - 1));   st(7, 5, -x23 * r5 - X3 * r4 - X2);
    st(7, 6, X2 * q5);        st(7, 7, X2 * q4);
    st(7, 8, -x23 * q5 - X3 * q4 + Y3 * (r4 - r5));

    st(8, 0, 0);              st(8, 1, 0);               st(8, 2, Y3 * (r4 - r5));
    st(8, 3, -(X2 * q5));     st(8, 4, -(X2 * q4));      st(8, 5, X3 * q4 + x23 * q5);
    st(8, 6, X2 * (r5 - 2));  st(8, 7, X2 * (r4 - 2));
    st(8, 8, -x23 * r5 - X3 * r4 + 4 * X2 + Y3 * (q5 - q4));
}

__device__ void dkt_alpha_T(const double *sc, double aT[9][9])
{
    dkt_alpha_T_gen(sc, [&](int i, int j, double v) { aT[i][j] = v; });
}
__device__ __forceinline__ void plane_stress(double E, double nu, double &C00, double &C01, double &C22);
__device__ __forceinline__ void membrane_B(const double *sc, double A, double Bm[3][6]);

// membrane data of a shell that never changes: the three columns of ke_m (shell.c:487-531) that the membrane
// displacement vector dm = (0,0,dm2,0,dm4,dm5) multiplies [0..17], the plane-stress coefficients [18..20] and
// t*A0*C/(2 A0)^2 [21..23]
__device__ __forceinline__ void shell_der(const double *sc, double *der)
{
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    plane_stress(sc[0], sc[1], C[0][0], C[0][1], C[2][2]);
    C[1][1] = C[0][0]; C[1][0] = C[0][1];
    double Bm[3][6];
    membrane_B(sc, sc[4], Bm);
    const int col[3] = {2, 4, 5};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double BC[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {           // Bm_C[i][j] (shell.c:513-521)
            double sum = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) sum += Bm[k][i] * C[k][j];
            BC[j] = sum;
        }
#pragma unroll
        for (int jc = 0; jc < 3; ++jc) {
            double sum = 0;                      // ke_m[i][j] (shell.c:522-530)
#pragma unroll
            for (int k = 0; k < 3; ++k) sum += BC[k] * Bm[k][col[jc]];
            der[i * 3 + jc] = sc[2] * sc[4] * sum;
        }
    }
    der[18] = C[0][0]; der[19] = C[0][1]; der[20] = C[2][2];
    const double smc = sc[2] / (4 * sc[4]);
    der[21] = smc * C[0][0]; der[22] = smc * C[0][1]; der[23] = smc * C[2][2];
}

__global__ void __launch_bounds__(CB_TPB)
k_shell_init_keb(CbDev d, double *__restrict__ keb)
{
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_SH) return;
    double sc[CB_SH_CONST];
#pragma unroll
    for (int i = 0; i < CB_SH_CONST; ++i) sc[i] = SOA(d.sh_const, i, e, d.NE_SH);
    const double E = sc[0], nu = sc[1], t3 = sc[3], A0 = sc[4];
    const double E1 = E * t3 / (12 * (1 - nu * nu));
    const double E3 = E1;
    const double E2 = E * t3 / (12 * (1 - nu * nu)) * nu;
    const double E4 = E * t3 / (12 * (1 - nu * nu)) * (1 - nu) / 2;
    double aT[9][9], Q[9][9];
    dkt_alpha_T(sc, aT);
    for (int i = 0; i < 9; ++i) {           // shell.c:610-648, the three row blocks are alike
        double b1 = 0, b2 = 0, b3 = 0;
        for (int j = 0; j < 3; ++j) {
            b1 += E1 * aT[i][j] + E2 * aT[i][j + 3];
            b2 += E2 * aT[i][j] + E3 * aT[i][j + 3];
            b3 += E4 * aT[i][j + 6];
        }
        for (int j = 0; j < 3; ++j) {
            Q[i][j] = (E1 * aT[i][j] + E2 * aT[i][j + 3] + b1) / 24;
            Q[i][j + 3] = (E2 * aT[i][j] + E3 * aT[i][j + 3] + b2) / 24;
            Q[i][j + 6] = (E4 * aT[i][j + 6] + b3) / 24;
        }
    }
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) {
            double sum = 0;
            for (int k = 0; k < 9; ++k) sum += Q[i][k] * aT[j][k];
            SOA(keb, CB_KEB(i, j), e, d.NE_SH) = sum / (2 * A0);
        }
    double der[CB_SH_DER];
    shell_der(sc, der);
    for (int i = 0; i < CB_SH_DER; ++i) SOA(d.sh_der, i, e, d.NE_SH) = der[i];
}

int cbk_shell_init_keb(const CbDev &d, double *keb_out, cudaStream_t s)
{
    if (d.NE_SH == 0) return 0;
    unsigned g = (unsigned)((d.NE_SH + CB_TPB - 1) / CB_TPB);
    k_shell_init_keb<<<g, CB_TPB, 0, s>>>(d, keb_out);
    return cudaGetLastError() != cudaSuccess;
}

// contribution-ordered copy of the 3x3 DKT sub-blocks: the assembly kernel reads it with fully
// coalesced loads and without waiting for the contribution record (static across iterations)
__global__ void __launch_bounds__(256)
k_shell_init_kebc(CbDev d, const CbContrib *__restrict__ contribs, long ncontrib,
                  double *__restrict__ kebc)
{
    long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (c >= ncontrib) return;
    const CbContrib ct = contribs[c];
    double *o = kebc + c * 10;
    if (ct.type != CB_T_SHELL) {
#pragma unroll
        for (int i = 0; i < 10; ++i) o[i] = 0.0;
        return;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i] = SOA(d.sh_keb, (3 * ct.a + ct.b) * 9 + i, ct.e, d.NE_SH);
    o[9] = 0.0;
}

int cbk_shell_init_kebc(const CbDev &d, const CbContrib *contribs, long ncontrib, double *kebc,
                        cudaStream_t s)
{
    if (ncontrib == 0 || d.NE_SH == 0) return 0;
    unsigned g = (unsigned)((ncontrib + 255) / 256);
    k_shell_init_kebc<<<g, 256, 0, s>>>(d, contribs, ncontrib, kebc);
    return cudaGetLastError() != cudaSuccess;
}

// duo plan: the same sub-blocks in work-major order, kebc[(T * 18 + u * 9 + i) * CB_T2_T + t] for
// work item t of tile T (u = its contribution 0/1): one CTA per tile
__global__ void __launch_bounds__(CB_T2_T)
k_shell_init_kebc2(CbDev d, const CbTile2 *__restrict__ tiles, const CbWork *__restrict__ works,
                   const CbContrib *__restrict__ contribs, double *__restrict__ kebc)
{
    const CbTile2 tl = tiles[blockIdx.x];
    const int t = threadIdx.x;
    if (t >= tl.nw) return;
    const CbWork w = works[tl.w0 + t];
    double *o = kebc + blockIdx.x * (18L * CB_T2_T) + t;
    for (int u = 0; u < 2; ++u) {
        const int a = u ? w.a1 : w.a0, b = u ? w.b1 : w.b0;
        const long e = (u < w.n) ? contribs[w.c0 + u].e : -1;
#pragma unroll
        for (int i = 0; i < 9; ++i)
            o[(u * 9 + i) * CB_T2_T] = (e >= 0) ? SOA(d.sh_keb, (3 * a + b) * 9 + i, e, d.NE_SH) : 0.0;
    }
}

int cbk_shell_init_kebc2(const CbDev &d, const CbTile2 *tiles, long ntiles, const CbWork *works,
                         const CbContrib *contribs, double *kebc, cudaStream_t s)
{
    if (ntiles == 0 || d.NE_SH == 0) return 0;
    k_shell_init_kebc2<<<(unsigned)ntiles, CB_T2_T, 0, s>>>(d, tiles, works, contribs, kebc);
    return cudaGetLastError() != cudaSuccess;
}

// stream plan: the sub-block of every step, kebc[((tile * S + step) * 9 + i) * 32 + lane]: a warp's nine
// loads per step are nine contiguous 256-byte runs.  One warp per tile.
__global__ void __launch_bounds__(32)
k_shell_init_kebcS(CbDev d, const CbTileS *__restrict__ tiles, int S, int slots, const uint32_t *__restrict__ steps,
                   const int32_t *__restrict__ elems, double *__restrict__ kebc)
{
    const CbTileS tl = tiles[blockIdx.x];
    const int lane = threadIdx.x;
    for (int st = 0; st < S; ++st) {
        const long row = (long)blockIdx.x * S + st;
        const uint32_t r = (st < tl.nsteps) ? steps[row * 32 + lane] : CB_S_IDLE;
        const unsigned slot = r & 63u;
        const int a = (r >> 6) & 3, b = (r >> 8) & 3;
        const long e = (slot != CB_S_IDLE) ? elems[(long)blockIdx.x * slots + slot] : -1;
#pragma unroll
        for (int i = 0; i < 9; ++i)
            kebc[(row * 9 + i) * 32 + lane] = (e >= 0) ? SOA(d.sh_keb, (3 * a + b) * 9 + i, e, d.NE_SH) : 0.0;
    }
}

int cbk_shell_init_kebcS(const CbDev &d, const CbTileS *tiles, long ntiles, int steps_per_tile, int slots_per_tile,
                         const uint32_t *steps, const int32_t *elems, double *kebc, cudaStream_t s)
{
    if (ntiles == 0 || d.NE_SH == 0) return 0;
    k_shell_init_kebcS<<<(unsigned)ntiles, 32, 0, s>>>(d, tiles, steps_per_tile, slots_per_tile, steps, elems, kebc);
    return cudaGetLastError() != cudaSuccess;
}

// geometry classes: one copy of the DKT matrix / derived membrane data per class, and the work
// records of the duo plan with the classes of their two contributions packed into c0
__global__ void __launch_bounds__(128)
k_class_tables(CbDev d, const int32_t *__restrict__ rep, int ncls, double *__restrict__ keb_tab,
               double *__restrict__ keb_tab10, double *__restrict__ keb_row, double *__restrict__ der_tab)
{
    const int k = blockIdx.x, c = threadIdx.x;
    if (k >= ncls) return;
    const long e = rep[k];
    // keb_row: what the local block (a, b) of a shell of this class holds besides its state (cb_stiff.cu,
    // s_contrib_cls).  CST gradients (x 2A) of the local joints: (-Y3, X3 - X2), (Y3, -X3), (0, X2); membrane
    // coefficients t A0 C / (2 A0)^2 from sh_der[21..23]; stiff_sh's drilling term k_thetay,thetay / 1e4 (shell.c:482)
    for (int i = c; i < 9 * CB_KROW; i += blockDim.x) {
        const int ab = i / CB_KROW, f = i - ab * CB_KROW, a = ab / 3, b = ab - 3 * a;
        const double X2 = SOA(d.sh_const, 5, e, d.NE_SH), X3 = SOA(d.sh_const, 6, e, d.NE_SH), Y3 = SOA(d.sh_const, 7, e, d.NE_SH);
        const double bx[3] = {-Y3, Y3, 0.0}, by[3] = {X3 - X2, -X3, X2};
        const double c00 = SOA(d.sh_der, 21, e, d.NE_SH), c01 = SOA(d.sh_der, 22, e, d.NE_SH), c22 = SOA(d.sh_der, 23, e, d.NE_SH);
        const double bxa = bx[a], bya = by[a], bxb = bx[b], byb = by[b];
        double v = 0.0;
        if (f < 9) v = SOA(d.sh_keb, ab * 9 + f, e, d.NE_SH);
        else if (f == 9) v = (a == b) ? SOA(d.sh_keb, ab * 9 + 4, e, d.NE_SH) * 1e-4 : 0.0;
        else if (f == 10) v = c00 * bxa * bxb + c22 * bya * byb;
        else if (f == 11) v = c01 * bxa * byb + c22 * bya * bxb;
        else if (f == 12) v = c01 * bya * bxb + c22 * bxa * byb;
        else if (f == 13) v = c00 * bya * byb + c22 * bxa * bxb;
        else if (f == 14) v = bxa * bxb;
        else if (f == 15) v = bya * byb;
        else if (f == 16) v = bxa * byb + bya * bxb;
        keb_row[(long)k * 9 * CB_KROW + i] = v;
    }
    if (c < 81) keb_tab[k * 81 + c] = SOA(d.sh_keb, c, e, d.NE_SH);
    if (c < 90) keb_tab10[k * 90 + c] = (c % 10 < 9) ? SOA(d.sh_keb, (c / 10) * 9 + c % 10, e, d.NE_SH) : 0.0;
    if (c < CB_SH_DER) der_tab[k * CB_SH_DER + c] = SOA(d.sh_der, c, e, d.NE_SH);
}
__global__ void __launch_bounds__(256)
k_works_set_class(const CbWork *__restrict__ works, long nworks, const CbContrib *__restrict__ contribs,
                  const int32_t *__restrict__ cls, CbWork *__restrict__ out)
{
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nworks) return;
    CbWork w = works[i];
    int c0 = 0, c1 = 0;
    if (w.kind != 4 && w.n >= 1) c0 = cls[contribs[w.c0].e];
    if (w.kind != 4 && w.n >= 2) c1 = cls[contribs[w.c0 + 1].e];
    w.c0 = c0 | (c1 << 16);
    out[i] = w;
}
int cbk_shell_class_tables(const CbDev &d, const int32_t *rep, int ncls, double *keb_tab, double *keb_tab10, double *keb_row, double *der_tab,
                           const CbWork *works, long nworks, const CbContrib *contribs, CbWork *works_cls,
                           cudaStream_t s)
{
    k_class_tables<<<ncls, 128, 0, s>>>(d, rep, ncls, keb_tab, keb_tab10, keb_row, der_tab);
    if (nworks && works_cls)
        k_works_set_class<<<(unsigned)((nworks + 255) / 256), 256, 0, s>>>(works, nworks, contribs, d.sh_class, works_cls);
    return cudaGetLastError() != cudaSuccess;
}

// plane-stress constitutive coefficients (shell.c:497-501 / 672-676)
__device__ __forceinline__ void plane_stress(double E, double nu, double &C00, double &C01,
                                             double &C22)
{
    C00 = E / (1 - nu * nu);
    C01 = E / (1 - nu * nu) * nu;
    C22 = E / (1 - nu * nu) * (1 - nu) / 2;
}

// membrane strain-displacement matrix with area `A` (shell.c:503-510 / 678-686)
__device__ __forceinline__ void membrane_B(const double *sc, double A, double Bm[3][6])
{
    const double X2 = sc[5], X3 = sc[6], Y3 = sc[7];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) Bm[i][j] = 0;
    Bm[0][0] = Bm[2][1] = -(Y3 / (2 * A));
    Bm[0][2] = Bm[2][3] = Y3 / (2 * A);
    Bm[1][1] = Bm[2][0] = (X3 - X2) / (2 * A);
    Bm[1][3] = Bm[2][2] = -(X3 / (2 * A));
    Bm[1][5] = Bm[2][4] = X2 / (2 * A);
}

// current local membrane coordinates (mem_coord, shell.c:2402-2446) minus the reference ones
// (shell.c:164-167): returns dm[2], dm[4], dm[5]
__device__ __forceinline__ void membrane_dm(const double *xj, const double *xk, const double *xl,
                                            const double *R, const double *sc, double &dm2,
                                            double &dm4, double &dm5)
{
    double Xk[3], Xl[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { Xk[m] = xk[m] - xj[m]; Xl[m] = xl[m] - xj[m]; }
    const double x01 = dot3(R, Xk);        // T[0][:] . X[:,1]
    const double x02 = dot3(R, Xl);        // T[0][:] . X[:,2]
    const double x12 = dot3(R + 3, Xl);    // T[1][:] . X[:,2]
    dm5 = x12 - sc[7];
    dm4 = x02 - sc[6];
    dm2 = x01 - sc[5];
}

// ------------------------------------------------------------------------------------------
// stiffness-pass record (krec) of one shell: triad, reference local coordinates, membrane
// coefficients t*A0*C/(2A0)^2 and geometric coefficients A_def*Nm/(2A_def)^2, with the membrane
// force resultants Nm evaluated exactly as stiffg_sh does (shell.c:688-704).  The record is
// gathered per contribution by the assembly kernel, so it is stored AoS [NE][18]; the 32
// records of a warp are transposed through shared memory and written as one contiguous run.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void shell_krec(double X2, double X3, double Y3, double thick,
                                           const double *R /*[10]*/, const double *cst /*C00,C01,C22,
                                           smc*C00, smc*C01, smc*C22*/, double dm2, double dm4,
                                           double dm5, int anaflag, double *kr)
{
    // membrane strains with the deformed area (Bm of shell.c:678-686 applied to dm), then
    // Nm = t C eps.  Nm only feeds the geometric stiffness (1e-12 tolerance), so it is evaluated
    // in its cheapest form; dm itself is the reference-rounded quantity.
    const double i2a = 1.0 / (2 * R[9]);
    const double exx = Y3 * dm2 * i2a;
    const double eyy = X2 * dm5 * i2a;
    const double gxy = (X2 * dm4 - X3 * dm2) * i2a;
    const double Nx = thick * (cst[0] * exx + cst[1] * eyy);
    const double Ny = thick * (cst[1] * exx + cst[0] * eyy);
    const double Nxy = thick * cst[2] * gxy;
#pragma unroll
    for (int i = 0; i < 9; ++i) kr[i] = R[i];
    kr[9] = X2; kr[10] = X3; kr[11] = Y3;
    kr[12] = cst[3]; kr[13] = cst[4]; kr[14] = cst[5];
    const double gs = (anaflag >= 2) ? 0.5 * i2a : 0.0;          // 1 / (4 A_def)
    kr[15] = gs * Nx; kr[16] = gs * Ny; kr[17] = gs * Nxy;
}

// all 32 lanes of the warp must call this (kr may be garbage for lanes past the last element)
__device__ __forceinline__ void warp_store_krec(double *krec_base, long e_warp0, long ne,
                                                const double *kr, double (*tile)[CB_SH_KREC + 1])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < CB_SH_KREC; ++i) tile[lane][i] = kr[i];
    __syncwarp();
    const long nval = ((ne - e_warp0) < 32 ? (ne - e_warp0) : 32) * CB_SH_KREC;
    double *dst = krec_base + e_warp0 * CB_SH_KREC;
    for (int i = lane; i < nval; i += 32) dst[i] = tile[i / CB_SH_KREC][i % CB_SH_KREC];
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// stand-alone prep for the stiffness pass (used when the record left by the last force pass does
// not describe the state the stiffness is evaluated at: first iteration of an increment,
// committed-state stiffness of the arc-length driver, linear analysis)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CB_TPB)
k_shell_prep(CbDev d, const double *__restrict__ x, const double *__restrict__ frame)
{
    __shared__ double tile[CB_TPB / 32][32][CB_SH_KREC + 1];
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const bool live = e < d.NE_SH;
    double kr[CB_SH_KREC];
    if (live) {
        double sc[CB_SH_CONST], R[CB_SH_FRAME];
#pragma unroll
        for (int i = 0; i < CB_SH_CONST; ++i) sc[i] = SOA(d.sh_const, i, e, d.NE_SH);
#pragma unroll
        for (int i = 0; i < CB_SH_FRAME; ++i) R[i] = SOA(frame, i, e, d.NE_SH);
        const int4 nd = reinterpret_cast<const int4 *>(d.sh_nodes)[e];
        double xj[3], xk[3], xl[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            xj[m] = x[(long)nd.x * 3 + m]; xk[m] = x[(long)nd.y * 3 + m]; xl[m] = x[(long)nd.z * 3 + m];
        }
        double dm2, dm4, dm5;
        membrane_dm(xj, xk, xl, R, sc, dm2, dm4, dm5);
        double cst[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) cst[i] = SOA(d.sh_der, 18 + i, e, d.NE_SH);
        shell_krec(sc[5], sc[6], sc[7], sc[2], R, cst, dm2, dm4, dm5, d.ANAFLAG, kr);
    }
    warp_store_krec(d.sh_Nm, e - (threadIdx.x & 31), d.NE_SH, kr, tile[threadIdx.x >> 5]);
}

int cbk_shell_prep(const CbDev &d, const double *x, const double *sh_frame, cudaStream_t s)
{
    if (d.NE_SH == 0) return 0;
    unsigned g = (unsigned)((d.NE_SH + CB_TPB - 1) / CB_TPB);
    k_shell_prep<<<g, CB_TPB, 0, s>>>(d, x, sh_frame);
    return cudaGetLastError() != cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// updatc, nodal part (misc.c:83-93): x_ip <- x_temp ; x_temp += dd[jcode-1] on free DOFs 1-3
// ------------------------------------------------------------------------------------------
// The same launch also does d_temp += dd (main.c:1949) over the equations [0, nq) handed in through
// ax / ay: two ~20 us kernels of a 1.3 ms step become one.
__global__ void __launch_bounds__(256)
k_node_update(long NJ, const int32_t *__restrict__ jc, const double *__restrict__ dd,
              double *__restrict__ x_temp, double *__restrict__ x_ip, long nq,
              const double *__restrict__ ax, double *__restrict__ ay)
{
    long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (t < nq) ay[t] += ax[t];
    if (t >= NJ * 3) return;
    long i = t / 3; int j = (int)(t - i * 3);
    int k = jc[i * 8 + j];
    double xv = x_temp[t];
    x_ip[t] = xv;
    if (k != 0) x_temp[t] = xv + dd[k - 1];
}

int cbk_node_update(const CbForceArgs &a, cudaStream_t s)
{
    const long nj = a.jl1 - a.jl0, n = std::max(nj * 3, a.axpy_n);
    if (n <= 0) return 0;
    unsigned g = (unsigned)((n + 255) / 256);
    k_node_update<<<g, 256, 0, s>>>(nj > 0 ? nj : 0, a.d.jc + a.jl0 * 8, a.dd, a.x_temp + a.jl0 * 3, a.x_ip + a.jl0 * 3,
                                    a.axpy_n, a.axpy_x, a.axpy_y);
    return cudaGetLastError() != cudaSuccess;
}

// gather the 6 incremental (or total) displacements of one node; fixed DOFs read as 0
__device__ __forceinline__ void gather_node6(const int32_t *jc, const double *v, int node,
                                             double *D)
{
    const int4 a = reinterpret_cast<const int4 *>(jc)[(long)node * 2];
    const int4 b = reinterpret_cast<const int4 *>(jc)[(long)node * 2 + 1];
    D[0] = a.x ? v[a.x - 1] : 0.0; D[1] = a.y ? v[a.y - 1] : 0.0; D[2] = a.z ? v[a.z - 1] : 0.0;
    D[3] = a.w ? v[a.w - 1] : 0.0; D[4] = b.x ? v[b.x - 1] : 0.0; D[5] = b.y ? v[b.y - 1] : 0.0;
}

// updatc, shell part (misc.c:153-184): side lengths, area, triad from the current coordinates
__device__ __forceinline__ void shell_triad(const double *xj, const double *xk, const double *xl,
                                            double *R /*[10]*/, double *dsl /*[3]*/)
{
    double el12[3], el23[3], el31[3], normal[3], lx[3], ly[3], lz[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        el23[m] = xl[m] - xk[m];
        el31[m] = xl[m] - xj[m];
        el12[m] = xk[m] - xj[m];
    }
    dsl[1] = sqrt(dot3(el23, el23));
    dsl[2] = sqrt(dot3(el31, el31));
    dsl[0] = sqrt(dot3(el12, el12));
    cross3(el12, el31, normal, false);
    const double A = 0.5 * sqrt(dot3(normal, normal));
#pragma unroll
    for (int m = 0; m < 3; ++m) { lx[m] = el12[m] / dsl[0]; lz[m] = normal[m] / (2 * A); }
    cross3(lz, lx, ly, true);
#pragma unroll
    for (int m = 0; m < 3; ++m) { R[m] = lx[m]; R[3 + m] = ly[m]; R[6 + m] = lz[m]; }
    R[9] = A;
}

// ------------------------------------------------------------------------------------------
// forces_sh, ANAFLAG 2 (shell.c:1728-1785, 2305-2347, 2386-2397) fused with the shell block of
// updatc.  One thread per element.
// ------------------------------------------------------------------------------------------
#ifndef CB_FORCES_CTAS
#define CB_FORCES_CTAS 4          // class tables: latency-bound, 4 resident CTAs per SM (128 registers)
#endif                            // per-element matrices: HBM-bound, 2 CTAs with 254 registers measure faster
// FUSE: x_temp holds the coordinates BEFORE this iteration's update; the kernel forms x + dd itself and
// the first shell at a joint writes the result to x_new (CbForceArgs::fuse_node)
#ifndef CB_FORCES_CTAS_NOCLS
#define CB_FORCES_CTAS_NOCLS 3
#endif
// threads per CTA of k_shell_forces: one warp.  The kernel's only block-wide barrier (before the warps' tiles
// reuse the staging columns) then waits for nobody else: 128-thread CTAs measured 3.6 % slower (0.443 vs 0.427 ms)
#ifndef CB_SHF_TPB
#define CB_SHF_TPB 32
#endif
#define CB_SHF_SCALE (128 / CB_SHF_TPB)
template <bool CLS, bool FUSE>
__global__ void __launch_bounds__(CB_SHF_TPB, (CLS ? CB_FORCES_CTAS : CB_FORCES_CTAS_NOCLS) * CB_SHF_SCALE)
k_shell_forces(CbDev d, const double *__restrict__ x_temp, const double *__restrict__ dd,
               const double *__restrict__ frame_ip, double *__restrict__ frame_i,
               double *__restrict__ dsl_i, const double *__restrict__ ef_ip,
               double *__restrict__ ef_i, double *__restrict__ x_new)
{
    // [99][CB_SHF_TPB]: the element's DKT matrix (81) and previous end forces (18), copied straight
    // from HBM by cp.async at kernel entry so their latency overlaps the geometry update; the
    // region is reused for the krec transposition at the end.
    extern __shared__ double sbuf[];
    double (*tile)[32][CB_SH_KREC + 1] = reinterpret_cast<double (*)[32][CB_SH_KREC + 1]>(sbuf);
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const bool live = e < d.NE_SH;
    double kr[CB_SH_KREC];
    double *mycol = sbuf + threadIdx.x;
    constexpr bool RECOMP = !CLS && CB_FORCES_RECOMPUTE_KEB;
    constexpr bool STAGE = !CLS && CB_FORCES_STAGE_KEB && !RECOMP;
    constexpr int KOFF = STAGE ? 81 : 0;                           // first column of the staged ef_ip
    // geometry-constant data: per element, or one L1-resident copy per geometry class
    const double *kebsrc = nullptr, *der = nullptr;
    long kstr = 0, dstr = 0;
    if (live) {
        if (CLS) {
            const int cls = __ldg(d.sh_class + e);
            kebsrc = d.keb_tab + (long)cls * 81; kstr = 1;
            der = d.der_tab + (long)cls * CB_SH_DER; dstr = 1;
        } else {
            kebsrc = d.sh_keb + e; kstr = d.NE_SH;
            der = d.sh_der + e; dstr = d.NE_SH;
        }
    }
    if (live) {
        const unsigned sdst = (unsigned)__cvta_generic_to_shared(mycol);
        // per-element matrix: staged in this thread's column.  Class table: read in place later -
        // the lanes of a warp mostly share a class, so those loads are L1 broadcasts, whereas
        // 81 private copies per thread would saturate the L1 data pipe with identical bytes.
        if (STAGE) {
#pragma unroll
        for (int c = 0; c < 81; ++c)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sdst + c * CB_SHF_TPB * 8),
                         "l"(kebsrc + (long)c * kstr));
        }
#pragma unroll
        for (int c = 0; c < 18; ++c)
            if (c % 6 >= 2)                     // the membrane slots of ef_ip are dropped (shell.c:1773)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sdst + (KOFF + c) * CB_SHF_TPB * 8),
                             "l"(ef_ip + (long)c * d.NE_SH + e));
        asm volatile("cp.async.commit_group;");
    }
    double fg[18];            // element force in global axes: staged per corner, or summed per joint by the warp
    if (live) {
    double sc[CB_SH_CONST], Rp[CB_SH_FRAME], Ri[CB_SH_FRAME], dsl[3];
    sc[2] = __ldg(&SOA(d.sh_const, 2, e, d.NE_SH));               // thickness
#pragma unroll
    for (int i = 5; i < 8; ++i) sc[i] = __ldg(&SOA(d.sh_const, i, e, d.NE_SH));   // x2, x3, y3
    if (RECOMP) {
#pragma unroll
        for (int i = 0; i < 11; ++i)
            if (i != 2 && (i < 5 || i > 7)) sc[i] = __ldg(&SOA(d.sh_const, i, e, d.NE_SH));
    }
#pragma unroll
    for (int i = 0; i < CB_SH_FRAME; ++i) Rp[i] = __ldg(&SOA(frame_ip, i, e, d.NE_SH));
    const int4 nd = reinterpret_cast<const int4 *>(d.sh_nodes)[e];
    const int nn[3] = {nd.x, nd.y, nd.z};
    double X[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int m = 0; m < 3; ++m) X[a][m] = __ldg(&x_temp[(long)nn[a] * 3 + m]);
    double DD[3][6];
#pragma unroll
    for (int a = 0; a < 3; ++a) gather_node6(d.jc, dd, nn[a], DD[a]);
    if (FUSE) {
        // updatc, nodal part (misc.c:83-93): x_temp += dd on the free translations.  gather_node6 reads a
        // fixed DOF as 0 and x + 0.0 == x bit for bit, so every shell at a joint forms the same value;
        // the first one stores it
        const unsigned own = d.sh_own[e];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int m = 0; m < 3; ++m) X[a][m] = X[a][m] + DD[a][m];
            if ((own >> a) & 1u) {
#pragma unroll
                for (int m = 0; m < 3; ++m) x_new[(long)nn[a] * 3 + m] = X[a][m];
            }
        }
    }

    shell_triad(X[0], X[1], X[2], Ri, dsl);

    // membrane: total force from the current local coordinates (shell.c:1729-1744, 1767-1775)
    double dm2, dm4, dm5;
    membrane_dm(X[0], X[1], X[2], Ri, sc, dm2, dm4, dm5);
    // record for the next stiffness pass: after `_ip <- _i` this triad / area / dm are exactly
    // what stiff_sh evaluates (shell.c:159-171)
    double derl[RECOMP ? CB_SH_DER : 1];
    if (RECOMP) shell_der(sc, derl);          // same operations as k_shell_init_keb: the cached values bit for bit
    {
        double cst[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) cst[i] = RECOMP ? derl[RECOMP ? 18 + i : 0] : __ldg(der + (18 + i) * dstr);
        shell_krec(sc[5], sc[6], sc[7], sc[2], Ri, cst, dm2, dm4, dm5, d.ANAFLAG, kr);
    }
    // ke_m * dm with the geometry-constant membrane columns (shell.c:1767-1775): the structural
    // zeros of dm add nothing, the remaining three terms keep the reference's order
    double ef_temp[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double acc = 0;
        acc += (RECOMP ? derl[RECOMP ? i * 3 + 0 : 0] : __ldg(der + (i * 3 + 0) * dstr)) * dm2;
        acc += (RECOMP ? derl[RECOMP ? i * 3 + 1 : 0] : __ldg(der + (i * 3 + 1) * dstr)) * dm4;
        acc += (RECOMP ? derl[RECOMP ? i * 3 + 2 : 0] : __ldg(der + (i * 3 + 2) * dstr)) * dm5;
        ef_temp[i] = acc;
    }

    // bending: incremental force ke_b * ddb, ddb = T_ip * DD on the bending DOFs (1746-1785)
    double ddb[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        ddb[3 * a + 0] = dot3(Rp + 6, DD[a]);        // w       : c3_ip . translations
        ddb[3 * a + 1] = dot3(Rp + 0, DD[a] + 3);    // theta_x : c1_ip . rotations
        ddb[3 * a + 2] = dot3(Rp + 3, DD[a] + 3);    // theta_y : c2_ip . rotations
    }
    double defb[9];
    if (RECOMP) {
        // ke_b = alpha W alpha^T / (2 A0) (shell.c:610-658: Q = alpha W, ke_b = Q alpha^T / (2 A0)), so
        // ke_b ddb = alpha (W (alpha^T ddb)) / (2 A0): the entries of alpha are closed forms of six lengths and
        // are evaluated in registers for both products; W is the closed form of the three row blocks of
        // shell.c:610-648
        double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, u[9];
        dkt_alpha_T_gen(sc, [&](int i, int j, double a) { v[j] += a * ddb[i]; });      // v = alpha^T ddb
        const double D0 = sc[0] * sc[3] / (12 * (1 - sc[1] * sc[1]));
        const double E1 = D0, E2 = D0 * sc[1], E4 = D0 * (1 - sc[1]) / 2;
        const double s0 = v[0] + v[1] + v[2], s1 = v[3] + v[4] + v[5], s2 = v[6] + v[7] + v[8];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            u[j] = (E1 * (v[j] + s0) + E2 * (v[j + 3] + s1)) / 24;
            u[j + 3] = (E2 * (v[j] + s0) + E1 * (v[j + 3] + s1)) / 24;
            u[j + 6] = (E4 * (v[j + 6] + s2)) / 24;
        }
        const double inv2A = 1.0 / (2 * sc[4]);
#pragma unroll
        for (int i = 0; i < 9; ++i) defb[i] = 0;
        dkt_alpha_T_gen(sc, [&](int i, int j, double a) { defb[i] += a * u[j]; });     // alpha u
#pragma unroll
        for (int i = 0; i < 9; ++i) defb[i] *= inv2A;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");      // own copies only: no barrier needed
    if (!RECOMP) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double sum = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j)
            sum += (!STAGE ? __ldg(kebsrc + CB_KEB(i, j) * kstr) : mycol[CB_KEB(i, j) * CB_SHF_TPB]) * ddb[j];
        defb[i] = sum;
    }
    }
    double efp[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) efp[i] = (i % 6 >= 2) ? mycol[(KOFF + i) * CB_SHF_TPB] : 0.0;

    // T_i * T_ip^T is block diagonal with M = R_i R_ip^T (shell.c:2326-2338)
    double M[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) M[r][s2] = dot3(Ri + 3 * r, Rp + 3 * s2);

    // ef_i = ef_temp + Ti_Tip (def + ef_ip) with the membrane slots of ef_ip zeroed (1773, 2342-2348)
    double efi[18];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double vt[3], vr[3];
        vt[0] = 0.0; vt[1] = 0.0;
        vt[2] = defb[3 * a] + efp[6 * a + 2];
        vr[0] = defb[3 * a + 1] + efp[6 * a + 3];
        vr[1] = defb[3 * a + 2] + efp[6 * a + 4];
        vr[2] = 0.0 + efp[6 * a + 5];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double et = (r < 2) ? ef_temp[2 * a + r] : 0.0;
            efi[6 * a + r] = et + dot3(M[r], vt);
            efi[6 * a + 3 + r] = 0.0 + dot3(M[r], vr);
        }
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) SOA(ef_i, i, e, d.NE_SH) = efi[i];
    // element force in global axes, T_i^T ef_i (shell.c:2388-2392)
#pragma unroll
    for (int g = 0; g < 6; ++g)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double sum = 0;
            sum += Ri[c] * efi[3 * g];
            sum += Ri[3 + c] * efi[3 * g + 1];
            sum += Ri[6 + c] * efi[3 * g + 2];
            fg[3 * g + c] = sum;
        }
    if (!d.ws_fg) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            double2 *o = reinterpret_cast<double2 *>(CB_FG(d.sh_fg, b, e, d.NE_SH));
#pragma unroll
            for (int i = 0; i < 3; ++i) o[i] = make_double2(fg[6 * b + 2 * i], fg[6 * b + 2 * i + 1]);
        }
    }
#pragma unroll
    for (int i = 0; i < CB_SH_FRAME; ++i) SOA(frame_i, i, e, d.NE_SH) = Ri[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) SOA(dsl_i, i, e, d.NE_SH) = dsl[i];
    }
    __syncthreads();          // everyone is done with sbuf before it becomes the warps' tiles
    if (d.ws_fg) {
        // warp-level partial sums (cb_wsum.cuh): the 32 x 18 force components of the warp's shells meet in its
        // tile, one lane per slot adds the (at most six) corners of its joint in element order and stages 48 bytes
        double (*t)[CB_SH_KREC + 1] = tile[threadIdx.x >> 5];
        const int lane = threadIdx.x & 31;
        if (live) {
#pragma unroll
            for (int i = 0; i < 18; ++i) t[lane][i] = fg[i];
        }
        __syncwarp();
        const long w = (e - lane) >> 5;
        if (w < d.ws_nwarp) {
            const int s1 = d.ws_start[w + 1];
            for (int sl = d.ws_start[w] + lane; sl < s1; sl += 32) {
                const unsigned long long pk = d.ws_corners[sl];
                const int cnt = (int)(pk & 7ull);
                double a6[6] = {0, 0, 0, 0, 0, 0};
                for (int k = 0; k < cnt; ++k) {
                    const int idx = (int)((pk >> (3 + 7 * k)) & 127ull), l = idx / 3, a = idx - 3 * l;
#pragma unroll
                    for (int j = 0; j < 6; ++j) a6[j] += t[l][6 * a + j];
                }
                double2 *o = reinterpret_cast<double2 *>(d.ws_fg + (long)sl * 6);
                o[0] = make_double2(a6[0], a6[1]); o[1] = make_double2(a6[2], a6[3]); o[2] = make_double2(a6[4], a6[5]);
            }
        }
        __syncwarp();         // the tile becomes the krec tile
    }
    warp_store_krec(d.sh_Nm, e - (threadIdx.x & 31), d.NE_SH, kr, tile[threadIdx.x >> 5]);
}

// ------------------------------------------------------------------------------------------
// ANAFLAG 3 shells.  k_shell_plastic_prep: what stiff_sh does before it builds the matrix
// (shell.c:171-283) - Ivanov's criterion at the three vertices from the current stress resultants,
// choice of the controlling vertex, and for a yielded shell the local elasto-plastic matrix
// (stiffm_sh) left in sh_kpl for the assembly kernels.  k_shell_forces_pl: updatc's shell block +
// forces_sh for ANAFLAG 3 (shell.c:1786-2325, 2349-2397), one thread per element; it updates the
// vertex stress resultants / equivalent plastic curvature in place like the reference does.
// ------------------------------------------------------------------------------------------
#include "cb_shell_plastic.cuh"

__global__ void __launch_bounds__(64)
k_shell_plastic_prep(CbDev d, const double *__restrict__ frame, const double *__restrict__ dsl,
                     const double *__restrict__ pl)
{
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_SH) return;
    double sc[CB_SH_CONST];
    for (int i = 0; i < CB_SH_CONST; ++i) sc[i] = SOA(d.sh_const, i, e, d.NE_SH);
    const double E = sc[0], nu = sc[1], t = sc[2], fy = d.sh_yield[e], No = fy * t;
    const double *P = pl + e * CB_SH_PL;
    Ivanov iv[3];
    double phi[3];
    int yv = 0;
    for (int i = 0; i < 3; ++i) {
        iv[i].r = iv[i].s = iv[i].phi = 0; iv[i].h = 0;
        ivanov_eval(iv[i], P + 3 + 3 * i, P + 12 + 3 * i, P[i], fy, t, No);
        phi[i] = (iv[i].q >= 1e-4) ? iv[i].phi : 0;
        if (iv[i].q >= 1e-4 && phi[i] >= 1 - CB_SH_PHITOL) pick_vertex(yv, phi, i);
    }
    d.sh_yv[e] = yv;
    if (yv == 0) return;
    const int v = yv - 1;
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    plane_stress(E, nu, C[0][0], C[0][1], C[2][2]);
    C[1][1] = C[0][0]; C[1][0] = C[0][1];
    IvFlow fl;
    ivanov_flow(fl, iv[v], P + 3 + 3 * v, P + 12 + 3 * v, C, E, t, fy, P[v], No);
    for (int i = 0; i < 3; ++i) sc[8 + i] = SOA(dsl, i, e, d.NE_SH);       // deformed side lengths
    double k[18][18];
    for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) k[i][j] = 0;
    shell_plastic_k(k, fl, C, sc, t, SOA(frame, 9, e, d.NE_SH));
    double *out = d.sh_kpl + e * 324;
    for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) out[i * 18 + j] = k[i][j];
}

int cbk_shell_plastic_prep(const CbDev &d, const double *sh_frame, const double *sh_dsl,
                           const double *sh_pl, cudaStream_t s)
{
    if (d.NE_SH == 0) return 0;
    k_shell_plastic_prep<<<(unsigned)((d.NE_SH + 63) / 64), 64, 0, s>>>(d, sh_frame, sh_dsl, sh_pl);
    return cudaGetLastError() != cudaSuccess;
}

__global__ void __launch_bounds__(64)
k_shell_forces_pl(CbDev d, const double *__restrict__ x_temp, const double *__restrict__ x_ip,
                  const double *__restrict__ dd, const double *__restrict__ frame_ip,
                  double *__restrict__ frame_i, const double *__restrict__ dsl_ip,
                  double *__restrict__ dsl_i, const double *__restrict__ ef_ip, double *__restrict__ ef_i)
{
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_SH) return;
    const int FM[6] = {0, 1, 6, 7, 12, 13}, FB[9] = {2, 3, 4, 8, 9, 10, 14, 15, 16};
    double sc[CB_SH_CONST], Rp[CB_SH_FRAME], Ri[CB_SH_FRAME], dsl[3];
    for (int i = 0; i < CB_SH_CONST; ++i) sc[i] = SOA(d.sh_const, i, e, d.NE_SH);
    for (int i = 0; i < CB_SH_FRAME; ++i) Rp[i] = SOA(frame_ip, i, e, d.NE_SH);
    const double E = sc[0], nu = sc[1], t = sc[2], t3 = t * t * t, fy = d.sh_yield[e], No = fy * t;
    const int4 nd = reinterpret_cast<const int4 *>(d.sh_nodes)[e];
    const int nn[3] = {nd.x, nd.y, nd.z};
    double X[3][3], Xp[3][3];
    for (int a = 0; a < 3; ++a)
        for (int m = 0; m < 3; ++m) { X[a][m] = x_temp[(long)nn[a] * 3 + m]; Xp[a][m] = x_ip[(long)nn[a] * 3 + m]; }
    shell_triad(X[0], X[1], X[2], Ri, dsl);                      // updatc, misc.c:153-184

    // total membrane displacements w.r.t. the reference shape (record for the next stiffness pass)
    double dm2, dm4, dm5;
    membrane_dm(X[0], X[1], X[2], Ri, sc, dm2, dm4, dm5);
    {
        double cst[6], kr[CB_SH_KREC];
        for (int i = 0; i < 6; ++i) cst[i] = SOA(d.sh_der, 18 + i, e, d.NE_SH);
        shell_krec(sc[5], sc[6], sc[7], t, Ri, cst, dm2, dm4, dm5, d.ANAFLAG, kr);
        for (int i = 0; i < CB_SH_KREC; ++i) d.sh_Nm[e * CB_SH_KREC + i] = kr[i];
    }
    // incremental membrane displacements: local coordinates now minus those of the previous
    // iterate in the previous frame (shell.c:1787-1799)
    double ddm[6] = {0, 0, 0, 0, 0, 0};
    {
        const double zero[CB_SH_CONST] = {0};
        double a2, a4, a5, b2, b4, b5;
        membrane_dm(X[0], X[1], X[2], Ri, zero, a2, a4, a5);
        membrane_dm(Xp[0], Xp[1], Xp[2], Rp, zero, b2, b4, b5);
        ddm[2] = a2 - b2; ddm[4] = a4 - b4; ddm[5] = a5 - b5;
    }
    double DD[3][6], ddl[18], ddb[9];
    for (int a = 0; a < 3; ++a) {
        gather_node6(d.jc, dd, nn[a], DD[a]);
        for (int r = 0; r < 3; ++r) {
            ddl[6 * a + r] = dot3(Rp + 3 * r, DD[a]);
            ddl[6 * a + 3 + r] = dot3(Rp + 3 * r, DD[a] + 3);
        }
    }
    for (int i = 0; i < 9; ++i) ddb[i] = ddl[FB[i]];
    double scd[CB_SH_CONST];                                     // constants with the deformed sides
    for (int i = 0; i < CB_SH_CONST; ++i) scd[i] = sc[i];
    for (int i = 0; i < 3; ++i) scd[8 + i] = SOA(dsl_ip, i, e, d.NE_SH);
    const double Adef = Rp[9];
    double strn[3], curv[3][3];
    strain_curvature(strn, curv, ddm, ddb, scd, Adef);
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    plane_stress(E, nu, C[0][0], C[0][1], C[2][2]);
    C[1][1] = C[0][0]; C[1][0] = C[0][1];

    double *P = d.sh_pl + e * CB_SH_PL;
    Ivanov iv[3];
    double phi[3] = {0, 0, 0};
    int yv = 0, code = 0;
    for (int i = 0; i < 3 && code == 0; ++i) {
        double *chi = P + i, *N = P + 3 + 3 * i, *M = P + 12 + 3 * i;
        iv[i].r = iv[i].s = iv[i].phi = 0; iv[i].h = 0;
        ivanov_eval(iv[i], N, M, *chi, fy, t, No);
        if (iv[i].q >= 1e-4 && iv[i].phi >= 1 - CB_SH_PHITOL) {   // shell.c:1877-1992
            IvFlow fl;
            ivanov_flow(fl, iv[i], N, M, C, E, t, fy, *chi, No);
            double fs = 0, fc = 0;
            for (int j = 0; j < 3; ++j) { fs += fl.fnC[j] * strn[j]; fc += fl.fmC[j] * curv[i][j]; }
            fs *= t; fc *= t3 / 12;
            const double lambda = (fs + fc) / (fl.jf + fl.kf - fl.Bf * fl.df_da * fl.da_dchi);
            *chi += sqrt(sq((E * t) / (3 * fy)) * sq(fl.Bf * lambda));
            for (int j = 0; j < 3; ++j) {
                double s1 = 0, s2 = 0;
                for (int q = 0; q < 3; ++q) {
                    s1 += C[j][q] * (strn[q] - lambda * fl.fn[q]);
                    s2 += C[j][q] * (curv[i][q] - lambda * fl.fm[q]);
                }
                N[j] += t * s1; M[j] += t3 * s2 / 12;
            }
        } else {                                                 // elastic increment
            for (int j = 0; j < 3; ++j) {
                double s1 = 0, s2 = 0;
                for (int q = 0; q < 3; ++q) { s1 += C[j][q] * strn[q]; s2 += C[j][q] * curv[i][q]; }
                N[j] += t * s1; M[j] += t3 * s2 / 12;
            }
        }
        ivanov_eval(iv[i], N, M, *chi, fy, t, No);               // shell.c:1994-2041
        if (iv[i].q >= 1e-4) {
            phi[i] = iv[i].phi;
            if (phi[i] > 1 + 10 * CB_SH_PHITOL) {
                code = 1;                                        // shell.c:2044-2046
            } else if (phi[i] > 1 + CB_SH_PHITOL) {              // return to the surface
                int guard = 0;
                do {
                    IvFlow fl;
                    ivanov_flow(fl, iv[i], N, M, C, E, t, fy, *chi, No);
                    const double lambda = (phi[i] - 1) / (fl.jf + fl.kf - fl.Bf * fl.df_da * fl.da_dchi);
                    for (int j = 0; j < 3; ++j) {
                        double s1 = 0, s2 = 0;
                        for (int q = 0; q < 3; ++q) {
                            s1 += C[j][q] * (-lambda * fl.fn[q]);
                            s2 += C[j][q] * (-lambda * fl.fm[q]);
                        }
                        N[j] += t * s1; M[j] += t3 * s2 / 12;
                    }
                    ivanov_eval(iv[i], N, M, *chi, fy, t, No);
                    if (iv[i].q >= 1e-4) phi[i] = iv[i].phi;
                    // the reference would spin for ever on a non-converging return; the device
                    // gives up and asks for a smaller increment instead
                    if (++guard > 200) { code = 1; break; }
                } while (phi[i] > 1 + CB_SH_PHITOL);
                if (code == 0) pick_vertex(yv, phi, i);
            } else if (phi[i] >= 1 - CB_SH_PHITOL) {
                pick_vertex(yv, phi, i);
            }
        } else {
            phi[i] = 0;
        }
    }
    if (code != 0) atomicMin(d.sh_trip, d.sh_gid ? d.sh_gid[e] : (int)e);

    // M = R_i R_ip^T: T_i T_ip^T is block diagonal (shell.c:2352-2362)
    double M3[3][3];
    for (int r = 0; r < 3; ++r)
        for (int s2 = 0; s2 < 3; ++s2) M3[r][s2] = dot3(Ri + 3 * r, Rp + 3 * s2);
    double def[18], eft[18], efp[18], efi[18];
    for (int i = 0; i < 18; ++i) { def[i] = 0; eft[i] = 0; efp[i] = SOA(ef_ip, i, e, d.NE_SH); }
    if (yv == 0) {                                               // shell.c:2215-2278 (= ANAFLAG 2)
        for (int i = 0; i < 6; ++i) {
            double acc = 0;
            acc += SOA(d.sh_der, i * 3 + 0, e, d.NE_SH) * dm2;
            acc += SOA(d.sh_der, i * 3 + 1, e, d.NE_SH) * dm4;
            acc += SOA(d.sh_der, i * 3 + 2, e, d.NE_SH) * dm5;
            eft[FM[i]] = acc;
            efp[FM[i]] = 0;
        }
        for (int i = 0; i < 9; ++i) {
            double sum = 0;
            for (int j = 0; j < 9; ++j) sum += SOA(d.sh_keb, CB_KEB(i, j), e, d.NE_SH) * ddb[j];
            def[FB[i]] = sum;
        }
    } else {                                                     // shell.c:2279-2306
        const int v = yv - 1;
        IvFlow fl;
        ivanov_flow(fl, iv[v], P + 3 + 3 * v, P + 12 + 3 * v, C, E, t, fy, P[v], No);
        double k[18][18];
        for (int i = 0; i < 18; ++i)
            for (int j = 0; j < 18; ++j) k[i][j] = 0;
        shell_plastic_k(k, fl, C, scd, t, Adef);
        for (int i = 0; i < 18; ++i) {
            double sum = 0;
            for (int j = 0; j < 18; ++j) sum += k[i][j] * ddl[j];
            def[i] = sum;
        }
    }
    // ef_i = [ef_temp +] Ti_Tip (def + ef_ip)   (shell.c:2364-2384)
    for (int g = 0; g < 6; ++g) {
        double v3[3];
        for (int c = 0; c < 3; ++c) v3[c] = def[3 * g + c] + efp[3 * g + c];
        for (int r = 0; r < 3; ++r) efi[3 * g + r] = eft[3 * g + r] + dot3(M3[r], v3);
    }
    for (int i = 0; i < 18; ++i) SOA(ef_i, i, e, d.NE_SH) = efi[i];
    for (int g = 0; g < 6; ++g)
        for (int c = 0; c < 3; ++c) {
            double sum = 0;
            sum += Ri[c] * efi[3 * g];
            sum += Ri[3 + c] * efi[3 * g + 1];
            sum += Ri[6 + c] * efi[3 * g + 2];
            CB_FG(d.sh_fg, g / 2, e, d.NE_SH)[(g % 2) * 3 + c] = sum;
        }
    for (int i = 0; i < CB_SH_FRAME; ++i) SOA(frame_i, i, e, d.NE_SH) = Ri[i];
    for (int i = 0; i < 3; ++i) SOA(dsl_i, i, e, d.NE_SH) = dsl[i];
}

// forces_sh, ANAFLAG 1 (shell.c:1695-1727, 2386-2397): ef = k_sh (T D), total displacements
__global__ void __launch_bounds__(CB_TPB)
k_shell_forces_linear(CbDev d, const double *__restrict__ dtot, const double *__restrict__ frame,
                      double *__restrict__ ef_out)
{
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_SH) return;
    double sc[CB_SH_CONST], R[CB_SH_FRAME];
#pragma unroll
    for (int i = 0; i < CB_SH_CONST; ++i) sc[i] = SOA(d.sh_const, i, e, d.NE_SH);
#pragma unroll
    for (int i = 0; i < CB_SH_FRAME; ++i) R[i] = SOA(frame, i, e, d.NE_SH);
    const int4 nd = reinterpret_cast<const int4 *>(d.sh_nodes)[e];
    const int nn[3] = {nd.x, nd.y, nd.z};
    double D[3][6], dl[18];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        gather_node6(d.jc, dtot, nn[a], D[a]);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            dl[6 * a + r] = dot3(R + 3 * r, D[a]);
            dl[6 * a + 3 + r] = dot3(R + 3 * r, D[a] + 3);
        }
    }
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    plane_stress(sc[0], sc[1], C[0][0], C[0][1], C[2][2]);
    C[1][1] = C[0][0]; C[1][0] = C[0][1];
    double Bm[3][6];
    membrane_B(sc, sc[4], Bm);
    const int fm[6] = {0, 1, 6, 7, 12, 13};
    const int fb[9] = {2, 3, 4, 8, 9, 10, 14, 15, 16};
    double ef[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) ef[i] = 0;
    for (int i = 0; i < 6; ++i) {
        double BC[3];
        for (int j = 0; j < 3; ++j) {
            double sum = 0;
            for (int k = 0; k < 3; ++k) sum += Bm[k][i] * C[k][j];
            BC[j] = sum;
        }
        double acc = 0;
        for (int j = 0; j < 6; ++j) {
            double sum = 0;
            for (int k = 0; k < 3; ++k) sum += BC[k] * Bm[k][j];
            acc += (sc[2] * sc[4] * sum) * dl[fm[j]];
        }
        ef[fm[i]] = acc;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double sum = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j) sum += SOA(d.sh_keb, CB_KEB(i, j), e, d.NE_SH) * dl[fb[j]];
        ef[fb[i]] = sum;
    }
    // drilling stiffness ke_b[1][1]/1e4 etc. (shell.c:482-484)
    ef[5] = (SOA(d.sh_keb, CB_KEB(1, 1), e, d.NE_SH) / 10000) * dl[5];
    ef[11] = (SOA(d.sh_keb, CB_KEB(4, 4), e, d.NE_SH) / 10000) * dl[11];
    ef[17] = (SOA(d.sh_keb, CB_KEB(7, 7), e, d.NE_SH) / 10000) * dl[17];
#pragma unroll
    for (int i = 0; i < 18; ++i) SOA(ef_out, i, e, d.NE_SH) = ef[i];
#pragma unroll
    for (int g = 0; g < 6; ++g)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double sum = 0;
            sum += R[c] * ef[3 * g];
            sum += R[3 + c] * ef[3 * g + 1];
            sum += R[6 + c] * ef[3 * g + 2];
            CB_FG(d.sh_fg, g / 2, e, d.NE_SH)[(g % 2) * 3 + c] = sum;
        }
}

// ------------------------------------------------------------------------------------------
// truss: updatc (misc.c:97-108) + forces_tr (truss.c:326-376 nonlinear, 243-325 linear)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CB_TPB)
k_truss_forces(CbDev d, const double *__restrict__ x_temp, double *__restrict__ frame_i,
               double *__restrict__ ef_i)
{
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_TR) return;
    const int j = d.tr_nodes[e * 2], k = d.tr_nodes[e * 2 + 1];
    double el[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) el[m] = x_temp[(long)k * 3 + m] - x_temp[(long)j * 3 + m];
    const double dl = sqrt(dot3(el, el));
    const double c[3] = {el[0] / dl, el[1] / dl, el[2] / dl};
    frame_i[e * 4 + 0] = c[0]; frame_i[e * 4 + 1] = c[1]; frame_i[e * 4 + 2] = c[2];
    frame_i[e * 4 + 3] = dl;
    const double E = d.tr_const[e * 4 + 0], A = d.tr_const[e * 4 + 1], L0 = d.tr_const[e * 4 + 2];
    const double strain = (dl - L0) / L0;
    double ef0 = E * A * (strain + 0.5 * (strain * strain)) * dl / L0;
    double ef1 = -ef0;
    if (d.ANAFLAG == 3) {                   // axial force capped at the squash load (truss.c:335-347)
        const double Py = d.tr_py[e], r = ef0 / Py;
        if (r * r >= 1 - 1e-4) {
            if (ef0 < 0) { ef0 = -Py; ef1 = Py; } else { ef0 = Py; ef1 = -Py; }
        }
    }
    ef_i[e * 2] = ef0; ef_i[e * 2 + 1] = ef1;
#pragma unroll
    for (int m = 0; m < 3; ++m) {           // f_temp -= ef * c  (truss.c:351-375)
        d.tr_fg[e * 6 + m] = -(ef0 * c[m]);
        d.tr_fg[e * 6 + 3 + m] = -(ef1 * c[m]);
    }
}

__global__ void __launch_bounds__(CB_TPB)
k_truss_forces_linear(CbDev d, const double *__restrict__ dtot, const double *__restrict__ frame,
                      double *__restrict__ ef_out)
{
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_TR) return;
    const int nj = d.tr_nodes[e * 2], nk = d.tr_nodes[e * 2 + 1];
    double D[6];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        int q = d.jc[(long)nj * 8 + m]; D[m] = q ? dtot[q - 1] : 0.0;
        q = d.jc[(long)nk * 8 + m];     D[3 + m] = q ? dtot[q - 1] : 0.0;
    }
    const double c[3] = {frame[e * 4], frame[e * 4 + 1], frame[e * 4 + 2]};
    const double d0 = dot3(c, D), d1 = dot3(c, D + 3);
    const double E = d.tr_const[e * 4 + 0], A = d.tr_const[e * 4 + 1], L0 = d.tr_const[e * 4 + 2];
    const double kk = A * E / L0, kn = -(A * E / L0);
    double s0 = 0; s0 += kk * d0; s0 += kn * d1;
    double s1 = 0; s1 += kn * d0; s1 += kk * d1;
    ef_out[e * 2] = s0; ef_out[e * 2 + 1] = s1;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        d.tr_fg[e * 6 + m] = -(s0 * c[m]);
        d.tr_fg[e * 6 + 3 + m] = -(s1 * c[m]);
    }
}

// ------------------------------------------------------------------------------------------
// frames: updatc (misc.c:112-147) + forces_fr (frame.c:902-1312, ANAFLAG 1 / 2, yldflag == 0).
// INPLACE = the linear call of main.c:1782 where ef_ip/ef_i and efFE_ip/efFE_i are the same
// arrays: rows are then updated in place, exactly as the reference does.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void frame_triad(const double *xa, const double *xb, const double *aux,
                                            double *F /*[10]*/)
{
    double el[3], lx[3], ly[3], lz[3], tmp[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) el[m] = xb[m] - xa[m];
    const double L = sqrt(dot3(el, el));
#pragma unroll
    for (int m = 0; m < 3; ++m) { lx[m] = el[m] / L; tmp[m] = aux[m] - xa[m]; }
    cross3(lx, tmp, lz, true);
    cross3(lz, lx, ly, true);
#pragma unroll
    for (int m = 0; m < 3; ++m) { F[m] = lx[m]; F[3 + m] = ly[m]; F[6 + m] = lz[m]; }
    F[9] = L;
}

// y = T x for T = blockdiag(R, R, 1, R, R, 1)  (frame.c:286-295); zeros of T are skipped
__device__ __forceinline__ void frame_T_apply(const double *R, const double *x, double *y)
{
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int o = (g < 2) ? 3 * g : 7 + 3 * (g - 2);
#pragma unroll
        for (int r = 0; r < 3; ++r) y[o + r] = dot3(R + 3 * r, x + o);
    }
    y[6] = 0.0 + 1.0 * x[6]; y[13] = 0.0 + 1.0 * x[13];
}

__device__ __forceinline__ void frame_Tt_apply(const double *R, const double *x, double *y)
{
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int o = (g < 2) ? 3 * g : 7 + 3 * (g - 2);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            s += R[c] * x[o]; s += R[3 + c] * x[o + 1]; s += R[6 + c] * x[o + 2];
            y[o + c] = s;
        }
    }
    y[6] = 0.0 + 1.0 * x[6]; y[13] = 0.0 + 1.0 * x[13];
}

// rigid-link matrix as forces_fr builds it (frame.c:1018-1032)
__device__ __forceinline__ void frame_rigid_link(const double *off, double (*Tr)[14])
{
    for (int i = 0; i < 14; ++i)
        for (int j = 0; j < 14; ++j) Tr[i][j] = (i == j) ? 1.0 : 0.0;
    Tr[3][1] = -off[2]; Tr[3][2] = off[1]; Tr[4][0] = off[2]; Tr[4][2] = -off[0];
    Tr[5][0] = -off[1]; Tr[5][1] = off[0];
    Tr[10][8] = -off[5]; Tr[10][9] = off[4]; Tr[11][7] = off[5]; Tr[11][9] = -off[3];
    Tr[12][7] = -off[4]; Tr[12][8] = off[3];
}

// def = k dl with the local tangent of frame e, generic version (plasticity, end releases, rigid end
// offsets: runtime indices, local memory)
__device__ __noinline__ void frame_def_generic(const CbDev &d, long e, const double *ef_ip,
                                               const double *efFE_ip, const double *Rp, const double *DD12,
                                               int os, double *dl, double *eft, double *def)
{
    double k[14][14];
    frame_local_k(d, e, ef_ip, efFE_ip, Rp[9], k, eft);
    if (os == 0) {
        frame_T_apply(Rp, DD12, dl);
    } else {
        double DDij[14], Tr[14][14];
        frame_rigid_link(d.fr_offset + e * 6, Tr);
        for (int i = 0; i < 14; ++i) {
            double s = 0;
            for (int j = 0; j < 14; ++j) s += Tr[j][i] * DD12[j];
            DDij[i] = s;
        }
        frame_T_apply(Rp, DDij, dl);
    }
    for (int i = 0; i < 14; ++i) {
        double s = 0;
        for (int j = 0; j < 14; ++j) s += k[i][j] * dl[j];
        def[i] = s;
    }
}

// EF <- T_rl EF (frame.c:1290-1309)
__device__ __noinline__ void frame_rigid_link_apply(const double *off, double *EF)
{
    double Tr[14][14], G2[14];
    frame_rigid_link(off, Tr);
    for (int i = 0; i < 14; ++i) {
        double s = 0;
        for (int j = 0; j < 14; ++j) s += Tr[i][j] * EF[j];
        G2[i] = s;
    }
    for (int i = 0; i < 14; ++i) EF[i] = G2[i];
}

// ------------------------------------------------------------------------------------------
// Frames of a model without rigid offsets, end releases or plasticity (CbDev::fr_simple): the same
// updatc + forces_fr arithmetic as k_frame_forces below, statement for statement, but
//   * def = K dl comes from frame_def_direct (cb_frame_def_gen.cuh): every tangent entry lives in a
//     register for the two products it feeds - no 14x14 array in local or shared memory, so the
//     resident warps are limited by registers only (k_frame_forces: 8 warps / SM behind its 54 KB
//     shared-memory columns, 255 registers + spills, profiles/r01e_prof_frame_forces.txt);
//   * the phases are fenced so that each one's inputs are loaded when it starts, not hoisted to the
//     top of the kernel (which is what pushed the single-phase kernel to 236 registers).
// ------------------------------------------------------------------------------------------
#ifndef CB_FR_SIMPLE_CTAS
#define CB_FR_SIMPLE_CTAS 6      // 168 registers: 2.25 ms at 5 M frames (8 CTAs / 128 registers + spills: 2.41 ms)
#endif
#define CB_PHASE_FENCE() asm volatile("" ::: "memory")

template <bool INPLACE>
__global__ void __launch_bounds__(CB_FR_TPB, CB_FR_SIMPLE_CTAS)
k_frame_forces_simple(CbDev d, const double *__restrict__ x_new, const double *__restrict__ dd,
                      const double *frame_ip, double *frame_i, double *xfr_i, const double *ef_ip,
                      double *ef_i, const double *efFE_ip, double *efFE_i, double dlpf, int itecnt)
{
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_FR) return;
    const int2 nn = reinterpret_cast<const int2 *>(d.fr_nodes)[e];
    const int nj = nn.x, nk = nn.y;
    double def[14];
    {   // ---- phase 1: def = (k_e + k_g) T_ip DD (frame.c:1013-1060) ------------------------------
        double DD12[14], dl[14], eft[14], Rp[CB_FR_FRAME], fc[10];
        ldv2<CB_FR_FRAME>(frame_ip + e * CB_FR_FRAME, Rp);
        ldv2<10>(d.fr_const + e * CB_FR_CONST, fc);
        {
            const int4 a0 = reinterpret_cast<const int4 *>(d.jc + (long)nj * 8)[0], a1 = reinterpret_cast<const int4 *>(d.jc + (long)nj * 8)[1];
            const int4 b0 = reinterpret_cast<const int4 *>(d.jc + (long)nk * 8)[0], b1 = reinterpret_cast<const int4 *>(d.jc + (long)nk * 8)[1];
            const int qa[7] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z}, qb[7] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z};
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                DD12[r] = qa[r] ? dd[qa[r] - 1] : 0.0;
                DD12[7 + r] = qb[r] ? dd[qb[r] - 1] : 0.0;
            }
        }
        frame_T_apply(Rp, DD12, dl);
        {   // the geometric tangent reads P = eft[7], M4, M5, M10, M11, M12 only
            double a[14], b[14];
#pragma unroll
            for (int i = 0; i < 14; ++i) { a[i] = 0; b[i] = 0; eft[i] = 0; }
            ldv2<2>(ef_ip + e * 14 + 4, a + 4);   ldv2<2>(efFE_ip + e * 14 + 4, b + 4);
            ldv2<2>(ef_ip + e * 14 + 6, a + 6);   ldv2<2>(efFE_ip + e * 14 + 6, b + 6);
            ldv2<4>(ef_ip + e * 14 + 10, a + 10); ldv2<4>(efFE_ip + e * 14 + 10, b + 10);
#pragma unroll
            for (int i = 4; i < 13; ++i)
                if (i == 4 || i == 5 || i == 7 || i >= 10) eft[i] = a[i] + b[i];
        }
        if (d.ANAFLAG == 2) frame_def_direct<true>(fc, eft, Rp[9], dl, def);
        else frame_def_direct<false>(fc, eft, Rp[9], dl, def);
    }
    CB_PHASE_FENCE();
    double Ri[CB_FR_FRAME], Rp[CB_FR_FRAME];
    ldv2<CB_FR_FRAME>(frame_ip + e * CB_FR_FRAME, Rp);
    if (!INPLACE) {   // ---- phase 2: updatc, frame block (misc.c:108-147) ---------------------------
        double xab[6], aux[4];
#pragma unroll
        for (int m = 0; m < 3; ++m) { xab[m] = x_new[(long)nj * 3 + m]; xab[3 + m] = x_new[(long)nk * 3 + m]; }
        stv2<6>(xfr_i + e * 6, xab);
        ldv2<4>(d.fr_const + e * CB_FR_CONST + 10, aux);
        frame_triad(xab, xab + 3, aux, Ri);
        stv2<CB_FR_FRAME>(frame_i + e * CB_FR_FRAME, Ri);
    } else {
#pragma unroll
        for (int i = 0; i < CB_FR_FRAME; ++i) Ri[i] = Rp[i];
    }
    // M = T_i T_ip^T: four copies of R_i R_ip^T and 1 on the warping DOFs (frame.c:1078-1086)
    double M[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) M[r][s2] = dot3(Ri + 3 * r, Rp + 3 * s2);
    CB_PHASE_FENCE();
    const bool addref = (itecnt == 0);          // yldflag is 0 throughout for ANAFLAG 1 / 2
    // ---- phase 3 (pass 0): ef_i = M (ef_ip + def), f = T_i^T ef_i; phase 4 (pass 1): fixed-end
    // forces (frame.c:1090-1155); with INPLACE later rows see the rows already rewritten
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        double cur[14], out[14], ref[14];
        ldv2<14>((pass == 0 ? ef_ip : efFE_ip) + e * 14, cur);
        if (pass == 1 && addref) ldv2<14>(d.fr_efFE_ref + e * 14, ref);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int o = (g < 2) ? 3 * g : 7 + 3 * (g - 2);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double v[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[c] = pass == 0 ? (cur[o + c] + def[o + c])
                                     : (addref ? (cur[o + c] + dlpf * ref[o + c]) : cur[o + c]);
                out[o + r] = dot3(M[r], v);
                if (INPLACE) cur[o + r] = out[o + r];
            }
        }
#pragma unroll
        for (int w = 6; w < 14; w += 7) {
            const double v = pass == 0 ? (cur[w] + def[w]) : (addref ? (cur[w] + dlpf * ref[w]) : cur[w]);
            out[w] = 0.0 + 1.0 * v;
        }
        stv2<14>((pass == 0 ? ef_i : efFE_i) + e * 14, out);
        if (pass == 0) {
            double EF[14];
            frame_Tt_apply(Ri, out, EF);
            stv2<14>(d.fr_fg + e * 14, EF);
        }
        CB_PHASE_FENCE();
    }
}

template <bool INPLACE>
__global__ void __launch_bounds__(CB_FR_TPB, 4)
k_frame_forces(CbDev d, const double *__restrict__ x_new, const double *__restrict__ dd,
               const double *frame_ip, double *frame_i, double *xfr_i, const double *ef_ip,
               double *ef_i, const double *efFE_ip, double *efFE_i, double dlpf, int itecnt)
{
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_FR) return;
    const double *fc = d.fr_const + e * CB_FR_CONST;
    const int nj = d.fr_nodes[e * 2], nk = d.fr_nodes[e * 2 + 1];
    const int os = d.fr_osflag[e];
    double Rp[CB_FR_FRAME], Ri[CB_FR_FRAME];
#pragma unroll
    for (int i = 0; i < CB_FR_FRAME; ++i) Rp[i] = frame_ip[e * CB_FR_FRAME + i];
    if (!INPLACE) {
        // updatc, frame block
        double xa[3], xb[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            xa[m] = x_new[(long)nj * 3 + m]; xb[m] = x_new[(long)nk * 3 + m];
            if (os != 0) { xa[m] = xa[m] + d.fr_offset[e * 6 + m]; xb[m] = xb[m] + d.fr_offset[e * 6 + 3 + m]; }
            xfr_i[e * 6 + m] = xa[m]; xfr_i[e * 6 + 3 + m] = xb[m];
        }
        frame_triad(xa, xb, fc + 10, Ri);
#pragma unroll
        for (int i = 0; i < CB_FR_FRAME; ++i) frame_i[e * CB_FR_FRAME + i] = Ri[i];
    } else {
#pragma unroll
        for (int i = 0; i < CB_FR_FRAME; ++i) Ri[i] = Rp[i];
    }
    double eft[14], DD12[14], dl[14], def[14];
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        int q = d.jc[(long)nj * 8 + r]; DD12[r] = q ? dd[q - 1] : 0.0;
        q = d.jc[(long)nk * 8 + r];     DD12[7 + r] = q ? dd[q - 1] : 0.0;
    }
    if (os == 0 && d.ANAFLAG != 3 && d.fr_mendrel[e * 5] != 1) {
        // no plasticity, end releases or rigid offsets: the local tangent (upper triangle) is built in
        // this thread's column of shared memory with static indices.  The generic path keeps 14x14
        // doubles per thread in LOCAL memory, which the L2 writes back to HBM (2.7 GB for 0.7 GB of
        // results on a 1.5 M-frame lattice, profiles/r01c_prof_frame_forces.txt).
        extern __shared__ double sk_all[];
        double *sk = sk_all + threadIdx.x;
        const double *fc2 = d.fr_const + e * CB_FR_CONST;
#pragma unroll
        for (int i = 0; i < 105; ++i) sk[i * CB_FR_TPB] = 0;
#pragma unroll
        for (int i = 0; i < 14; ++i) eft[i] = ef_ip[e * 14 + i] + efFE_ip[e * 14 + i];
        frame_elastic_packed(sk, fc2);
        if (d.ANAFLAG == 2) frame_geometric_packed(sk, eft, Rp[9], fc2[2], fc2[8]);
        frame_T_apply(Rp, DD12, dl);
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            double s = 0;
#pragma unroll
            for (int j = 0; j < 14; ++j) s += KP(i, j) * dl[j];
            def[i] = s;
        }
    } else {
        frame_def_generic(d, e, ef_ip, efFE_ip, Rp, DD12, os, dl, eft, def);
    }
    // M = T_i T_ip^T: four copies of R_i R_ip^T and 1 on the warping DOFs (frame.c:1078-1086)
    double M[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) M[r][s2] = dot3(Ri + 3 * r, Rp + 3 * s2);
    double efl[14], fel[14], efn[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) { efl[i] = ef_ip[e * 14 + i]; fel[i] = efFE_ip[e * 14 + i]; }
    // the increment of the fixed-end forces enters on the first iteration, for the ends that have
    // not yielded (frame.c:1100-1155; yldflag is 0 throughout for ANAFLAG 1 / 2)
    int y0 = 0, y1 = 0;
    if (!INPLACE && d.ANAFLAG == 3) { y0 = d.fr_yldflag[e * 2]; y1 = d.fr_yldflag[e * 2 + 1]; }
    const bool addref0 = (itecnt == 0) && y0 == 0, addref1 = (itecnt == 0) && y1 == 0;
    double feo[14];
    // frame.c:1090-1096 then 1100-1155; with INPLACE later rows see the rows already rewritten
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        double *cur = pass == 0 ? efl : fel;
        double out[14];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int o = (g < 2) ? 3 * g : 7 + 3 * (g - 2);
            const bool addref = (g < 2) ? addref0 : addref1;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double v[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[c] = pass == 0 ? (cur[o + c] + def[o + c])
                                     : (addref ? (cur[o + c] + dlpf * d.fr_efFE_ref[e * 14 + o + c]) : cur[o + c]);
                out[o + r] = dot3(M[r], v);
                if (INPLACE) cur[o + r] = out[o + r];
            }
        }
#pragma unroll
        for (int w = 6; w < 14; w += 7) {
            const bool addref = (w == 6) ? addref0 : addref1;
            const double v = pass == 0 ? (cur[w] + def[w])
                                       : (addref ? (cur[w] + dlpf * d.fr_efFE_ref[e * 14 + w]) : cur[w]);
            out[w] = 0.0 + 1.0 * v;
            if (INPLACE) cur[w] = out[w];
        }
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            if (pass == 0) { ef_i[e * 14 + i] = out[i]; efn[i] = out[i]; }
            else { efFE_i[e * 14 + i] = out[i]; feo[i] = out[i]; }
        }
    }
    if (!INPLACE && d.ANAFLAG == 3) {
        // yield check, return to the surface, elastic unloading (frame.c:1157-1268).  The reference
        // returns from the middle of its element loop; here every member reports what it would do
        // and k_frame_trip keeps the flags of the members up to the first one that trips.
        const double *pl = d.fr_plast + e * 3;
        const double Py = pl[0], Mpy = pl[1], Mpz = pl[2];
        double eti[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) eti[i] = efn[i] + feo[i];
        double p[2] = {eti[0] / Py, eti[7] / Py}, my[2] = {eti[4] / Mpy, eti[11] / Mpy};
        double mz[2] = {eti[5] / Mpz, eti[12] / Mpz};
        const double phi[2] = {fr_phi(p[0], my[0], mz[0]), fr_phi(p[1], my[1], mz[1])};
        int code = 0; double tau = 1.0;
        if (phi[0] > phi[1] && phi[0] > 1 + CB_PHITOL && y0 != 2) {
            tau = fr_regula_falsi(eft[0] / Py, (eti[0] - eft[0]) / Py, eft[4] / Mpy, (eti[4] - eft[4]) / Mpy,
                                  eft[5] / Mpz, (eti[5] - eft[5]) / Mpz);
            y0 = 1; code = 1;
        } else if (phi[1] > phi[0] && phi[1] > 1 + CB_PHITOL && y1 != 2) {
            tau = fr_regula_falsi(eft[7] / Py, (eti[7] - eft[7]) / Py, eft[11] / Mpy, (eti[11] - eft[11]) / Mpy,
                                  eft[12] / Mpz, (eti[12] - eft[12]) / Mpz);
            y1 = 1; code = 1;
        } else if ((phi[0] >= 1 - CB_PHITOL && y0 != 2) && (phi[1] >= 1 - CB_PHITOL && y1 != 2)) {
            y0 = y1 = 1;
        } else if (phi[0] >= 1 - CB_PHITOL && y0 != 2) {
            y0 = 1;
        } else if (phi[1] >= 1 - CB_PHITOL && y1 != 2) {
            y1 = 1;
        }
        if (code == 0 && (y0 == 1 || y1 == 1)) {
            double k[14][14];
            for (int i = 0; i < 14; ++i)
                for (int j = 0; j < 14; ++j) k[i][j] = 0;
            frame_elastic(k, fc);
            frame_geometric(k, eft, Rp[9], fc[2], fc[8]);
            const int u = fr_unload(phi, p, my, mz, pl, k, dl);
            if (u == 1) { y0 = y1 = 2; code = 2; }
            else if (u == 2) { y0 = 2; code = 2; }
            else if (u == 3) { y1 = 2; code = 2; }
        }
        d.fr_ynew[e * 2] = y0; d.fr_ynew[e * 2 + 1] = y1;
        d.fr_code[e] = code; d.fr_tau[e] = tau;
        if (code != 0) atomicMin(d.fr_trip, d.fr_gid ? d.fr_gid[e] : (int)e);
    }
    double EF[14];
    frame_Tt_apply(Ri, efn, EF);
    if (os != 0) frame_rigid_link_apply(d.fr_offset + e * 6, EF);
#pragma unroll
    for (int i = 0; i < 14; ++i) d.fr_fg[e * 14 + i] = EF[i];
}

// ------------------------------------------------------------------------------------------
// f_temp: segmented reduction over the node -> corner map (sorted by element type, element), one
// thread per node; reproduces the reference's summation order (all trusses, frames, shells)
// without atomics (replaces the `f_temp[mcode-1] += ...` scatters).
// ------------------------------------------------------------------------------------------
// ANAFLAG 3: forces_fr returns at the first member (lowest index) that overshoots the yield surface
// or unloads, having mutated yldflag of the members before it and of that member (frame.c:1184-
// 1268).  trip[0] = that index (INT_MAX if none).  The flags proposed by the force pass are kept
// for the members up to it; trip[1] / fr_tau of that member go back to the host.
__global__ void __launch_bounds__(256)
k_frame_trip(CbDev d)
{
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_FR) return;
    const int first = d.fr_trip[0];
    const int g = d.fr_gid ? d.fr_gid[e] : (int)e;
    if (g <= first) { d.fr_yldflag[e * 2] = d.fr_ynew[e * 2]; d.fr_yldflag[e * 2 + 1] = d.fr_ynew[e * 2 + 1]; }
    if (g == first) { d.fr_trip[1] = d.fr_code[e]; reinterpret_cast<double *>(d.fr_trip + 2)[0] = d.fr_tau[e]; }
}

__global__ void __launch_bounds__(256)
k_gather_f(CbDev d, long j0, long j1, const int32_t *__restrict__ cstart,
           const CbCorner *__restrict__ corners, double *__restrict__ f)
{
    long n = j0 + blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (n >= j1) return;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const int c0 = cstart[n], c1 = cstart[n + 1];
    // members from the first tripping one onwards never reach the scatter in the reference
    const int fr_end = (d.ANAFLAG == 3 && d.fr_trip) ? d.fr_trip[0] : 0x7fffffff;
    const int sh_end = (d.ANAFLAG == 3 && d.sh_trip) ? d.sh_trip[0] : 0x7fffffff;
    for (int c = c0; c < c1; ++c) {
        const CbCorner cr = corners[c];
        if (cr.type == CB_T_SHELL) {
            if ((d.sh_gid && sh_end != 0x7fffffff ? d.sh_gid[cr.e] : cr.e) >= sh_end) continue;
            const double *p = CB_FG(d.sh_fg, cr.b, cr.e, d.NE_SH);
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[r] += p[r];
        } else if (cr.type == CB_T_FRAME) {
            if ((d.fr_gid && fr_end != 0x7fffffff ? d.fr_gid[cr.e] : cr.e) >= fr_end) continue;
            const double *p = d.fr_fg + (long)cr.e * 14 + cr.b * 7;
#pragma unroll
            for (int r = 0; r < 7; ++r) acc[r] += p[r];
        } else if (cr.type == CB_T_TRUSS) {
            const double *p = d.tr_fg + (long)cr.e * 6 + cr.b * 3;
#pragma unroll
            for (int r = 0; r < 3; ++r) acc[r] += p[r];
        }
    }
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const int q = d.jc[n * 8 + r];
        if (q != 0) f[q - 1] = acc[r];
    }
}

// shell-only models with the nodal update fused into the force kernel: the same segmented reduction over
// the joints [jl0, jl1) this rank touches - f_temp for the ones it owns - plus d_temp += dd (main.c:1949) on
// every touched joint's equations
__global__ void __launch_bounds__(256)
k_gather_f_axpy(CbDev d, long jl0, long jl1, long jo0, long jo1, const int32_t *__restrict__ cstart,
                const CbCorner *__restrict__ corners, double *__restrict__ f, const double *__restrict__ dd,
                double *__restrict__ d_temp)
{
    const long n = jl0 + blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (n >= jl1) return;
    const int4 qa = reinterpret_cast<const int4 *>(d.jc)[n * 2], qb = reinterpret_cast<const int4 *>(d.jc)[n * 2 + 1];
    const int q[6] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y};
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const bool own = n >= jo0 && n < jo1;
    if (own && d.ws_fg) {                 // the joint's slots of warp-level partial sums, in warp order
        const int c0 = d.js_start[n], c1 = d.js_start[n + 1];
        for (int c = c0; c < c1; ++c) {
            const double2 *p = reinterpret_cast<const double2 *>(d.ws_fg + (long)d.js_slots[c] * 6);
            const double2 v0 = p[0], v1 = p[1], v2 = p[2];
            acc[0] += v0.x; acc[1] += v0.y; acc[2] += v1.x; acc[3] += v1.y; acc[4] += v2.x; acc[5] += v2.y;
        }
    } else if (own) {
        const int c0 = cstart[n], c1 = cstart[n + 1];
        for (int c = c0; c < c1; ++c) {
            const CbCorner cr = corners[c];
            const double2 *p = reinterpret_cast<const double2 *>(CB_FG(d.sh_fg, cr.b, cr.e, d.NE_SH));
            const double2 v0 = p[0], v1 = p[1], v2 = p[2];
            acc[0] += v0.x; acc[1] += v0.y; acc[2] += v1.x; acc[3] += v1.y; acc[4] += v2.x; acc[5] += v2.y;
        }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r)
        if (q[r] != 0) {
            if (own) f[q[r] - 1] = acc[r];
            d_temp[q[r] - 1] += dd[q[r] - 1];
        }
}

int cbk_gather_f(const CbForceArgs &a, cudaStream_t s)
{
    if (a.fuse_node) {
        const long nj = a.jl1 - a.jl0;
        if (nj <= 0) return 0;
        k_gather_f_axpy<<<(unsigned)((nj + 255) / 256), 256, 0, s>>>(a.d, a.jl0, a.jl1, a.jo0, a.jo1, a.node_cstart,
                                                                     a.corners, a.f_temp, a.dd, a.d_temp);
        return cudaGetLastError() != cudaSuccess;
    }
    const long nj = a.jo1 - a.jo0;
    if (nj <= 0) return 0;
    unsigned g = (unsigned)((nj + 255) / 256);
    k_gather_f<<<g, 256, 0, s>>>(a.d, a.jo0, a.jo1, a.node_cstart, a.corners, a.f_temp);
    return cudaGetLastError() != cudaSuccess;
}

#define CB_FR_SMEM (105 * CB_FR_TPB * sizeof(double))
static int frame_forces_configure()
{
    static CbPerDevice cfg{};
    int &done = cfg.v[cb_device_slot()];
    if (done) return 0;
    if (cudaFuncSetAttribute(k_frame_forces<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CB_FR_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_frame_forces<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CB_FR_SMEM) != cudaSuccess)
        return 1;
    done = 1;
    return 0;
}

// ANAFLAG 3, after the force kernels (and, on several GPUs, after the lowest tripping member index
// has been agreed on): keep the flags of the members up to it, publish its code and dlpf factor
int cbk_frame_trip(const CbDev &d, cudaStream_t s, long *launches)
{
    if (d.ANAFLAG != 3 || !d.NE_FR) return 0;
    k_frame_trip<<<(unsigned)((d.NE_FR + 255) / 256), 256, 0, s>>>(d);
    ++*launches;
    return cudaGetLastError() != cudaSuccess;
}

int cbk_forces(const CbForceArgs &a, cudaStream_t s, long *launches)
{
    const CbDev &d = a.d;
    if (d.NE_FR && frame_forces_configure()) return 1;
    if (d.NE_TR) {
        unsigned g = (unsigned)((d.NE_TR + CB_TPB - 1) / CB_TPB);
        k_truss_forces<<<g, CB_TPB, 0, s>>>(d, a.x_temp, a.tr_frame_i, a.tr_ef_i);
        ++*launches;
    }
    if (d.NE_FR) {
        unsigned g = (unsigned)((d.NE_FR + 63) / 64);
        if (d.ANAFLAG == 3) {
            static const int init[4] = {0x7fffffff, 0, 0, 0};
            if (cudaMemcpyAsync(d.fr_trip, init, sizeof init, cudaMemcpyHostToDevice, s) != cudaSuccess) return 1;
        }
        if (d.fr_simple)
            k_frame_forces_simple<false><<<g, CB_FR_TPB, 0, s>>>(d, a.x_temp, a.dd, a.fr_frame_ip, a.fr_frame_i,
                                               a.fr_xfr_i, a.fr_ef_ip, a.fr_ef_i, a.fr_efFE_ip,
                                               a.fr_efFE_i, a.dlpf, a.itecnt);
        else
            k_frame_forces<false><<<g, CB_FR_TPB, CB_FR_SMEM, s>>>(d, a.x_temp, a.dd, a.fr_frame_ip, a.fr_frame_i,
                                               a.fr_xfr_i, a.fr_ef_ip, a.fr_ef_i, a.fr_efFE_ip,
                                               a.fr_efFE_i, a.dlpf, a.itecnt);
        ++*launches;
    }
    if (d.NE_SH) {
        unsigned g = (unsigned)((d.NE_SH + CB_SHF_TPB - 1) / CB_SHF_TPB);
        // staged columns per thread (ef_ip, + the DKT matrix when CB_FORCES_STAGE_KEB) or the krec
        // transposition tile, whichever is larger
        const size_t cols = (CB_FORCES_STAGE_KEB && !CB_FORCES_RECOMPUTE_KEB) ? 99 : 18;
        const size_t smem = std::max(cols * CB_SHF_TPB, (size_t)(CB_SHF_TPB / 32) * 32 * (CB_SH_KREC + 1)) * sizeof(double);
        static CbPerDevice cfg{};
        int &configured = cfg.v[cb_device_slot()];
        if (!configured) {
            if (cudaFuncSetAttribute(k_shell_forces<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
                cudaFuncSetAttribute(k_shell_forces<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
                cudaFuncSetAttribute(k_shell_forces<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
                cudaFuncSetAttribute(k_shell_forces<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return 1;
            configured = 1;
        }
        if (d.ANAFLAG == 3) {
            static const int init[4] = {0x7fffffff, 0, 0, 0};
            if (cudaMemcpyAsync(d.sh_trip, init, sizeof init, cudaMemcpyHostToDevice, s) != cudaSuccess) return 1;
            k_shell_forces_pl<<<(unsigned)((d.NE_SH + 63) / 64), 64, 0, s>>>(
                d, a.x_temp, a.x_ip, a.dd, a.sh_frame_ip, a.sh_frame_i, a.sh_dsl_ip, a.sh_dsl_i,
                a.sh_ef_ip, a.sh_ef_i);
        } else if (a.fuse_node) {       // x_temp: coordinates before the update, x_ip's buffer receives the new ones
            if (d.sh_class)
                k_shell_forces<true, true><<<g, CB_SHF_TPB, smem, s>>>(d, a.x_temp, a.dd, a.sh_frame_ip, a.sh_frame_i,
                                                                   a.sh_dsl_i, a.sh_ef_ip, a.sh_ef_i, a.x_ip);
            else
                k_shell_forces<false, true><<<g, CB_SHF_TPB, smem, s>>>(d, a.x_temp, a.dd, a.sh_frame_ip, a.sh_frame_i,
                                                                    a.sh_dsl_i, a.sh_ef_ip, a.sh_ef_i, a.x_ip);
        } else if (d.sh_class)
            k_shell_forces<true, false><<<g, CB_SHF_TPB, smem, s>>>(d, a.x_temp, a.dd, a.sh_frame_ip, a.sh_frame_i,
                                                                a.sh_dsl_i, a.sh_ef_ip, a.sh_ef_i, nullptr);
        else
            k_shell_forces<false, false><<<g, CB_SHF_TPB, smem, s>>>(d, a.x_temp, a.dd, a.sh_frame_ip, a.sh_frame_i,
                                                                 a.sh_dsl_i, a.sh_ef_ip, a.sh_ef_i, nullptr);
        ++*launches;
    }
    return cudaGetLastError() != cudaSuccess;
}

int cbk_forces_linear(const CbForceArgs &a, const double *d_total, cudaStream_t s, long *launches)
{
    const CbDev &d = a.d;
    if (d.NE_FR && frame_forces_configure()) return 1;
    if (d.NE_TR) {
        unsigned g = (unsigned)((d.NE_TR + CB_TPB - 1) / CB_TPB);
        k_truss_forces_linear<<<g, CB_TPB, 0, s>>>(d, d_total, a.tr_frame_i, a.tr_ef_i);
        ++*launches;
    }
    if (d.NE_FR) {
        unsigned g = (unsigned)((d.NE_FR + 63) / 64);
        if (d.fr_simple)
            k_frame_forces_simple<true><<<g, CB_FR_TPB, 0, s>>>(d, a.x_temp, d_total, a.fr_frame_i, a.fr_frame_i,
                                              a.fr_xfr_i, a.fr_ef_i, a.fr_ef_i, a.fr_efFE_i,
                                              a.fr_efFE_i, 0.0, 0);
        else
            k_frame_forces<true><<<g, CB_FR_TPB, CB_FR_SMEM, s>>>(d, a.x_temp, d_total, a.fr_frame_i, a.fr_frame_i,
                                              a.fr_xfr_i, a.fr_ef_i, a.fr_ef_i, a.fr_efFE_i,
                                              a.fr_efFE_i, 0.0, 0);
        ++*launches;
    }
    if (d.NE_SH) {
        unsigned g = (unsigned)((d.NE_SH + CB_TPB - 1) / CB_TPB);
        k_shell_forces_linear<<<g, CB_TPB, 0, s>>>(d, d_total, a.sh_frame_i, a.sh_ef_i);
        ++*launches;
    }
    return cudaGetLastError() != cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// lumped / row-summed diagonal mass (SLVFLAG==0 layout): mass_tr truss.c:381-441, mass_sh
// shell.c:1505-1591.  The reference routines first overwrite llength / slength / farea with
// values recomputed from the committed coordinates; k_mass_refresh does the
// same to the device constants, then one thread per node sums its corners.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CB_TPB)
k_mass_refresh_shell(CbDev d, const double *__restrict__ x, double *__restrict__ sh_const)
{
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_SH) return;
    const int4 nd = reinterpret_cast<const int4 *>(d.sh_nodes)[e];
    double xj[3], xk[3], xl[3], el12[3], el23[3], el31[3], normal[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        xj[m] = x[(long)nd.x * 3 + m]; xk[m] = x[(long)nd.y * 3 + m]; xl[m] = x[(long)nd.z * 3 + m];
        el23[m] = xl[m] - xk[m]; el31[m] = xl[m] - xj[m]; el12[m] = xk[m] - xj[m];
    }
    SOA(sh_const, 9, e, d.NE_SH) = sqrt(dot3(el23, el23));
    SOA(sh_const, 10, e, d.NE_SH) = sqrt(dot3(el31, el31));
    SOA(sh_const, 8, e, d.NE_SH) = sqrt(dot3(el12, el12));
    cross3(el12, el31, normal, false);
    SOA(sh_const, 4, e, d.NE_SH) = 0.5 * sqrt(dot3(normal, normal));
}

__global__ void __launch_bounds__(CB_TPB)
k_mass_refresh_truss(CbDev d, const double *__restrict__ x, double *__restrict__ tr_const)
{
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_TR) return;
    const int j = d.tr_nodes[e * 2], k = d.tr_nodes[e * 2 + 1];
    double el[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) el[m] = x[(long)k * 3 + m] - x[(long)j * 3 + m];
    const double L = sqrt(dot3(el, el));
    tr_const[e * CB_TR_CONST + 2] = L;
    tr_const[e * CB_TR_CONST + 3] = cube_rn(L);
}

// mass_fr part 1 (frame.c:1326-1343): end coordinates and length from the committed x
__global__ void __launch_bounds__(CB_TPB)
k_mass_refresh_frame(CbDev d, const double *__restrict__ x, double *__restrict__ fr_const,
                     double *__restrict__ xfr)
{
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= d.NE_FR) return;
    const int j = d.fr_nodes[e * 2], k = d.fr_nodes[e * 2 + 1];
    double el[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        double xa = x[(long)j * 3 + m], xb = x[(long)k * 3 + m];
        if (d.fr_osflag[e] != 0) { xa = xa + d.fr_offset[e * 6 + m]; xb = xb + d.fr_offset[e * 6 + 3 + m]; }
        xfr[e * 6 + m] = xa; xfr[e * 6 + 3 + m] = xb;
        el[m] = xb - xa;
    }
    const double L = sqrt(dot3(el, el));
    fr_const[e * CB_FR_CONST + 3] = L;
    fr_const[e * CB_FR_CONST + 4] = L * L;
    fr_const[e * CB_FR_CONST + 5] = cube_rn(L);
}

__global__ void __launch_bounds__(256)
k_mass_gather(CbDev d, const int32_t *__restrict__ cstart, const CbCorner *__restrict__ corners,
              const double *__restrict__ dens_tr, const double *__restrict__ dens_fr,
              const double *__restrict__ dens_sh, double *__restrict__ sm)
{
    long n = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (n >= d.NJ) return;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const int c0 = cstart[n], c1 = cstart[n + 1];
    for (int c = c0; c < c1; ++c) {
        const CbCorner cr = corners[c];
        if (cr.type == CB_T_SHELL) {
            const double A0 = SOA(d.sh_const, 4, cr.e, d.NE_SH), th = SOA(d.sh_const, 2, cr.e, d.NE_SH);
            const double Mtot = dens_sh[cr.e] * A0 * th;
            const double mt = Mtot / 3;
            const double mr = Mtot / 3 * (th * th) / 12;
#pragma unroll
            for (int r = 0; r < 3; ++r) { acc[r] += mt; acc[3 + r] += mr; }
        } else if (cr.type == CB_T_FRAME) {
            // lumped (frame.c:1355-1360): rho A L / 2 on translations, rho A L^3 / 24 on rotations
            const double *fc = d.fr_const + (long)cr.e * CB_FR_CONST;
            const double mt = (dens_fr[cr.e] * fc[2] * fc[3]) / 24 * 12;
            const double mr = (dens_fr[cr.e] * fc[2] * fc[3]) / 24 * (fc[3] * fc[3]);
#pragma unroll
            for (int r = 0; r < 3; ++r) { acc[r] += mt; acc[3 + r] += mr; }
        } else if (cr.type == CB_T_TRUSS) {
            // consistent mass row-summed onto the diagonal, only over free partner DOFs
            // (truss.c:416-425): row ie gets m[ie][ie] + m[ie][ie+-3] if that DOF is free
            const double *tc = d.tr_const + (long)cr.e * CB_TR_CONST;
            const double m3 = (dens_tr[cr.e] * tc[1] * tc[2]) / 3;
            const double m6 = (dens_tr[cr.e] * tc[1] * tc[2]) / 6;
            const int other = d.tr_nodes[(long)cr.e * 2 + (1 - cr.b)];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int qo = d.jc[(long)other * 8 + r];
                // the reference loops je ascending: end 0 adds diag then partner, end 1 partner
                // then diag
                if (cr.b == 0) { acc[r] += m3; if (qo) acc[r] += m6; }
                else { if (qo) acc[r] += m6; acc[r] += m3; }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const int q = d.jc[n * 8 + r];
        if (q != 0) sm[q - 1] = acc[r];
    }
}

int cbk_mass(const CbDev &d, const double *x, double *sh_const_mut, double *tr_const_mut,
             double *fr_const_mut, double *fr_xfr, const double *dens_tr, const double *dens_fr,
             const double *dens_sh, const int32_t *node_cstart, const CbCorner *corners,
             double *sm, cudaStream_t s, long *launches)
{
    if (d.NE_FR) {
        unsigned g = (unsigned)((d.NE_FR + CB_TPB - 1) / CB_TPB);
        k_mass_refresh_frame<<<g, CB_TPB, 0, s>>>(d, x, fr_const_mut, fr_xfr); ++*launches;
    }
    if (d.NE_TR) {
        unsigned g = (unsigned)((d.NE_TR + CB_TPB - 1) / CB_TPB);
        k_mass_refresh_truss<<<g, CB_TPB, 0, s>>>(d, x, tr_const_mut); ++*launches;
    }
    if (d.NE_SH) {
        unsigned g = (unsigned)((d.NE_SH + CB_TPB - 1) / CB_TPB);
        k_mass_refresh_shell<<<g, CB_TPB, 0, s>>>(d, x, sh_const_mut); ++*launches;
    }
    unsigned g = (unsigned)((d.NJ + 255) / 256);
    k_mass_gather<<<g, 256, 0, s>>>(d, node_cstart, corners, dens_tr, dens_fr, dens_sh, sm); ++*launches;
    return cudaGetLastError() != cudaSuccess;
}
