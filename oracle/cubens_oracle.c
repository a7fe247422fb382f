/* cubens_oracle.c - TEST INFRASTRUCTURE ONLY (see cubens_oracle.h for the parity status).
 *
 * CPU restatement of the reference's per-iteration element + assembly path, dense and literal:
 * every element matrix is formed in full and rotated with the generic n^3 triple product, and
 * contributions are scattered in element order, exactly as the reference does.  Floating-point
 * operation order follows the cited reference lines (compile with -ffp-contract=off, which is
 * what gcc does for the reference on baseline x86-64), so results are bit-comparable with
 * oracle/_ref.  Nothing in the product links or includes this file.
 */
#include "cubens_oracle.h"
#include <math.h>
#include <string.h>

/* ---------------------------------------------------------------- small helpers (misc.c) */
static double dotn(const double *a, const double *b, int n)       /* misc.c:252-262 */
{
    double dp = 0;
    for (int i = 0; i < n; ++i) dp += a[i] * b[i];
    return dp;
}

static void crossv(const double *a, const double *b, double *c, int unit)   /* misc.c:264-282 */
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
    if (unit) {
        double len = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        for (int i = 0; i < 3; ++i) c[i] /= len;
    }
}

/* K = T^T k T, dense (misc.c:41-69) */
static void triple(const double *k, const double *T, double *K, int n)
{
    double tmp[24 * 24];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int m = 0; m < n; ++m) s += T[m * n + i] * k[m * n + j];
            tmp[i * n + j] = s;
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int m = 0; m < n; ++m) s += tmp[i * n + m] * T[m * n + j];
            K[i * n + j] = s;
        }
}

/* Gauss-Jordan inverse without pivoting (misc.c:284-343; the row swap there only triggers on an
 * exactly-zero leading entry, which SPD / well-shaped inputs never produce) */
static void gj_inverse(double *A, int n)
{
    double aug[4][8];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 2 * n; ++j)
            aug[i][j] = (j < n) ? A[n * i + j] : ((j - n == i) ? 1.0 : 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            if (j != i) {
                double m = aug[j][i] / aug[i][i];
                for (int k = 0; k < 2 * n; ++k) aug[j][k] -= m * aug[i][k];
            }
    for (int i = 0; i < n; ++i)
        for (int j = n; j < 2 * n; ++j) A[n * i + j - n] = aug[i][j] / aug[i][i];
}

/* scatter of one element matrix (shell.c:307-342, truss.c:167-202, frame.c:326-360):
 * skyline takes the local upper triangle, the dense layout stores K[je][ie] at row i, column j */
static void scatter(const orc_dims *D, double *ss, const double *K, int n, const long *mc,
                    const long *maxa)
{
    if (D->SLVFLAG == 0) {
        for (int je = 0; je < n; ++je) {
            long j = mc[je];
            if (j == 0) continue;
            for (int ie = 0; ie <= je; ++ie) {
                long i = mc[ie], k;
                if (i == 0) continue;
                k = (i > j) ? maxa[i - 1] + (i - j) : maxa[j - 1] + (j - i);
                ss[k - 1] += K[ie * n + je];
            }
        }
    } else {
        for (int ie = 0; ie < n; ++ie)
            for (int je = 0; je < n; ++je) {
                long i = mc[ie], j = mc[je];
                if (i != 0 && j != 0) ss[(i - 1) * D->NEQ + j - 1] += K[je * n + ie];
            }
    }
}

/* ------------------------------------------------------------------ model.c: codes, skylin */
long orc_codes(const orc_dims *D, long *mcode, long *jcode, const long *minc)
{   /* model.c:941-960 (no released warping joints), 992-1085 */
    long neq = 0;
    for (long i = 0; i < D->NJ; ++i)
        for (int j = 0; j < 7; ++j)
            if (jcode[i * 7 + j] != 0) jcode[i * 7 + j] = ++neq;
    long pm = 0, pc = 0;
    for (long e = 0; e < D->NE_TR; ++e, pm += 2, pc += 6)
        for (int a = 0; a < 2; ++a)
            for (int m = 0; m < 3; ++m) mcode[pc + a * 3 + m] = jcode[(minc[pm + a] - 1) * 7 + m];
    for (long e = 0; e < D->NE_FR; ++e, pm += 2, pc += 14)
        for (int a = 0; a < 2; ++a)
            for (int m = 0; m < 7; ++m) mcode[pc + a * 7 + m] = jcode[(minc[pm + a] - 1) * 7 + m];
    for (long e = 0; e < D->NE_SH; ++e, pm += 3, pc += 18)
        for (int a = 0; a < 3; ++a)
            for (int m = 0; m < 6; ++m) mcode[pc + a * 6 + m] = jcode[(minc[pm + a] - 1) * 7 + m];
    for (long e = 0; e < D->NE_BR; ++e, pm += 8, pc += 24)
        for (int a = 0; a < 8; ++a)
            for (int m = 0; m < 3; ++m) mcode[pc + a * 3 + m] = jcode[(minc[pm + a] - 1) * 7 + m];
    return neq;
}

long orc_skylin(const orc_dims *D, long *maxa, long *kht, const long *mcode)
{   /* model.c:1204-1281; bricks are not visited by the reference */
    const long ne[3] = {D->NE_TR, D->NE_FR, D->NE_SH};
    const int nd[3] = {6, 14, 18};
    long p = 0;
    for (long i = 0; i < D->NEQ; ++i) kht[i] = 0;
    for (int t = 0; t < 3; ++t)
        for (long e = 0; e < ne[t]; ++e, p += nd[t]) {
            long mn = D->NEQ;
            for (int j = 0; j < nd[t]; ++j)
                if (mcode[p + j] > 0 && mcode[p + j] < mn) mn = mcode[p + j];
            for (int j = 0; j < nd[t]; ++j) {
                long k = mcode[p + j];
                if (k != 0 && k - mn > kht[k - 1]) kht[k - 1] = k - mn;
            }
        }
    maxa[0] = 1;
    for (long i = 0; i < D->NEQ; ++i) maxa[i + 1] = maxa[i] + kht[i] + 1;
    return maxa[D->NEQ] - 1;
}

/* ----------------------------------------------------------------------- misc.c: updatc */
static void shell_frame(const double *x, long j, long k, long l, double *dsl, double *area,
                        double *lx, double *ly, double *lz)
{   /* misc.c:160-178 == shell.c:69-87 */
    double el12[3], el23[3], el31[3], nrm[3];
    for (int m = 0; m < 3; ++m) {
        el23[m] = x[l * 3 + m] - x[k * 3 + m];
        el31[m] = x[l * 3 + m] - x[j * 3 + m];
        el12[m] = x[k * 3 + m] - x[j * 3 + m];
    }
    dsl[1] = sqrt(dotn(el23, el23, 3));
    dsl[2] = sqrt(dotn(el31, el31, 3));
    dsl[0] = sqrt(dotn(el12, el12, 3));
    crossv(el12, el31, nrm, 0);
    *area = 0.5 * sqrt(dotn(nrm, nrm, 3));
    if (lx) {
        for (int m = 0; m < 3; ++m) { lx[m] = el12[m] / dsl[0]; lz[m] = nrm[m] / (2 * (*area)); }
        crossv(lz, lx, ly, 1);
    }
}

void orc_updatc(const orc_dims *D, double *x_temp, double *x_ip, double *xfr_temp, const double *dd,
                double *defllen_i, double *deffarea_i, double *defslen_i, const double *offset,
                const int *osflag, const double *auxpt, double *c1_i, double *c2_i, double *c3_i,
                const long *minc, const long *jcode)
{
    for (long i = 0; i < D->NJ; ++i)                               /* misc.c:83-93 */
        for (int j = 0; j < 3; ++j) {
            long k = jcode[i * 7 + j];
            x_ip[i * 3 + j] = x_temp[i * 3 + j];
            if (k != 0) x_temp[i * 3 + j] += dd[k - 1];
        }
    for (long e = 0; e < D->NE_TR; ++e) {                          /* misc.c:97-108 */
        long j = minc[e * 2] - 1, k = minc[e * 2 + 1] - 1;
        double el[3];
        for (int m = 0; m < 3; ++m) el[m] = x_temp[k * 3 + m] - x_temp[j * 3 + m];
        defllen_i[e] = sqrt(dotn(el, el, 3));
        c1_i[e] = el[0] / defllen_i[e]; c2_i[e] = el[1] / defllen_i[e]; c3_i[e] = el[2] / defllen_i[e];
    }
    long pm = D->NE_TR * 2;
    for (long e = 0; e < D->NE_FR; ++e) {                          /* misc.c:112-147 */
        long j = minc[pm + e * 2] - 1, k = minc[pm + e * 2 + 1] - 1;
        double el[3], lx[3], ly[3], lz[3], tmp[3];
        for (int m = 0; m < 3; ++m) {
            xfr_temp[e * 6 + m] = x_temp[j * 3 + m];
            xfr_temp[e * 6 + 3 + m] = x_temp[k * 3 + m];
            if (osflag[e] != 0) {
                xfr_temp[e * 6 + m] = x_temp[j * 3 + m] + offset[e * 6 + m];
                xfr_temp[e * 6 + 3 + m] = x_temp[k * 3 + m] + offset[e * 6 + 3 + m];
            }
        }
        for (int m = 0; m < 3; ++m) el[m] = xfr_temp[e * 6 + 3 + m] - xfr_temp[e * 6 + m];
        double L = sqrt(dotn(el, el, 3));
        defllen_i[D->NE_TR + e] = L;
        for (int m = 0; m < 3; ++m) { lx[m] = el[m] / L; tmp[m] = auxpt[e * 3 + m] - xfr_temp[e * 6 + m]; }
        crossv(lx, tmp, lz, 1);
        crossv(lz, lx, ly, 1);
        for (int m = 0; m < 3; ++m) {
            c1_i[D->NE_TR + e * 3 + m] = lx[m]; c2_i[D->NE_TR + e * 3 + m] = ly[m];
            c3_i[D->NE_TR + e * 3 + m] = lz[m];
        }
    }
    pm = D->NE_TR * 2 + D->NE_FR * 2;
    long pc = D->NE_TR + D->NE_FR * 3;
    for (long e = 0; e < D->NE_SH; ++e) {                          /* misc.c:153-184 */
        double lx[3], ly[3], lz[3];
        shell_frame(x_temp, minc[pm + e * 3] - 1, minc[pm + e * 3 + 1] - 1, minc[pm + e * 3 + 2] - 1,
                    defslen_i + e * 3, deffarea_i + e, lx, ly, lz);
        for (int m = 0; m < 3; ++m) {
            c1_i[pc + e * 3 + m] = lx[m]; c2_i[pc + e * 3 + m] = ly[m]; c3_i[pc + e * 3 + m] = lz[m];
        }
    }
}

/* ---------------------------------------------------------- ANAFLAG 3 (material nonlinear) */
#define PHITOL 1e-4                                    /* frame.c:37, truss.c:100, shell.c:37 */
static struct {
    const double *yield, *zstrong, *zweak; int *yldflag;             /* trusses, frames */
    double *chi_temp, *efN_temp, *efM_temp;                         /* shells: main.c:889-913 */
    const double *x_ip, *deffarea_ip, *defslen_ip;
} g_pl;
void orc_set_plastic(const double *yield, const double *zstrong, const double *zweak, int *yldflag)
{   /* the arrays main.c owns for ANAFLAG 3 (main.c:559-571, 802) */
    g_pl.yield = yield; g_pl.zstrong = zstrong; g_pl.zweak = zweak; g_pl.yldflag = yldflag;
}
void orc_set_plastic_sh(double *chi_temp, double *efN_temp, double *efM_temp, const double *x_ip,
                        const double *deffarea_ip, const double *defslen_ip)
{   /* the extra arguments stiff_sh / forces_sh take for ANAFLAG 3 (prototypes.h:187-191, 246-251) */
    g_pl.chi_temp = chi_temp; g_pl.efN_temp = efN_temp; g_pl.efM_temp = efM_temp;
    g_pl.x_ip = x_ip; g_pl.deffarea_ip = deffarea_ip; g_pl.defslen_ip = defslen_ip;
}

/* --------------------------------------------------------------------------------- shell.c */
static const int FM[6] = {0, 1, 6, 7, 12, 13};              /* membrane DOFs  (shell.c:1603) */
static const int FB[9] = {2, 3, 4, 8, 9, 10, 14, 15, 16};   /* bending DOFs   (shell.c:1604) */
static const int FG[9] = {0, 1, 2, 6, 7, 8, 12, 13, 14};    /* translations   (shell.c:759-839) */

static void plane_stress(double E, double nu, double C[3][3])   /* shell.c:497-501 */
{
    memset(C, 0, 9 * sizeof(double));
    C[0][0] = C[1][1] = E / (1 - nu * nu);
    C[0][1] = C[1][0] = E / (1 - nu * nu) * nu;
    C[2][2] = E / (1 - nu * nu) * (1 - nu) / 2;
}

static void membrane_B(const double *xl, double A, double Bm[3][6])   /* shell.c:503-510 */
{
    memset(Bm, 0, 18 * sizeof(double));
    Bm[0][0] = Bm[2][1] = -(xl[2] / (2 * A));
    Bm[0][2] = Bm[2][3] = xl[2] / (2 * A);
    Bm[1][1] = Bm[2][0] = (xl[1] - xl[0]) / (2 * A);
    Bm[1][3] = Bm[2][2] = -(xl[1] / (2 * A));
    Bm[1][5] = Bm[2][4] = xl[0] / (2 * A);
}

static void cst_membrane(double ke[6][6], double E, double nu, const double *xl, double t, double A)
{   /* stiffe_m_sh, shell.c:487-531 */
    double C[3][3], Bm[3][6], BC[6][3];
    plane_stress(E, nu, C);
    membrane_B(xl, A, Bm);
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += Bm[k][i] * C[k][j];
            BC[i][j] = s;
        }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += BC[i][k] * Bm[k][j];
            ke[i][j] = t * A * s;
        }
}

static void dkt_alpha_T(double aT[9][9], const double *xl, const double *sl);

static void dkt_bending(double kb[9][9], double E, double nu, const double *xl, double t, double A,
                        const double *sl)
{   /* stiffe_b_sh, shell.c:533-658: Batoz' explicit DKT matrix, kb = Q alpha^T / (2A) */
    const double E1 = E * pow(t, 3) / (12 * (1 - nu * nu)), E3 = E1;
    const double E2 = E * pow(t, 3) / (12 * (1 - nu * nu)) * nu;
    const double E4 = E * pow(t, 3) / (12 * (1 - nu * nu)) * (1 - nu) / 2;
    double aT[9][9];
    dkt_alpha_T(aT, xl, sl);
    double Q[9][9];
    for (int i = 0; i < 9; ++i) {        /* shell.c:610-648: the three row blocks are alike */
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
            double s = 0;
            for (int k = 0; k < 9; ++k) s += Q[i][k] * aT[j][k];
            kb[i][j] = s / (2 * A);
        }
}

static void dkt_alpha_T(double aT_out[9][9], const double *xl, const double *sl)
{   /* transpose of Batoz' alpha matrix, shell.c:545-607 (= 1249-1311, 1424-1486); the rows of
     * strn_curv's LL_alpha (shell.c:2489-2540) are its columns: LL_alpha[v][j][k] = aT[k][3j+v] */
    const double X2 = xl[0], X3 = xl[1], Y3 = xl[2];
    const double x23 = X2 - X3, l12 = sl[0] * sl[0], l23 = sl[1] * sl[1], l31 = sl[2] * sl[2];
    const double p4 = -6 * x23 / l23, p5 = -6 * X3 / l31, p6 = 6 * X2 / l12;
    const double t4 = 6 * Y3 / l23, t5 = -6 * Y3 / l31;
    const double q4 = -3 * x23 * Y3 / l23, q5 = 3 * X3 * Y3 / l31;
    const double r4 = 3 * (Y3 * Y3) / l23, r5 = 3 * (Y3 * Y3) / l31;
    const double aT[9][9] = {
        {Y3 * p6, -(Y3 * p6), Y3 * p5, -(X2 * t5), 0, x23 * t5, -(X3 * p6) - X2 * p5, -x23 * p6,
         x23 * p5 + Y3 * t5},
        {0, 0, -(Y3 * q5), x23 + X2 * r5, x23, x23 * (1 - r5), X2 * q5 + Y3, Y3,
         -x23 * q5 + Y3 * (1 - r5)},
        {-4 * Y3, 2 * Y3, Y3 * (2 - r5), -(X2 * q5), 0, x23 * q5, -4 * x23 + X2 * r5, 2 * x23,
         x23 * (2 - r5) + Y3 * q5},
        {-(Y3 * p6), Y3 * p6, Y3 * p4, 0, X2 * t4, -(X3 * t4), X3 * p6, x23 * p6 + X2 * p4,
         -(X3 * p4) + Y3 * t4},
        {0, 0, Y3 * q4, X3, X3 + X2 * r4, X3 * (1 - r4), -Y3, -Y3 + X2 * q4,
         Y3 * (r4 - 1) - X3 * q4},
        {-2 * Y3, 4 * Y3, Y3 * (r4 - 2), 0, -(X2 * q4), X3 * q4, 2 * X3, -4 * X3 + X2 * r4,
         X3 * (2 - r4) - Y3 * q4},
        {0, 0, -(Y3 * (p4 + p5)), X2 * t5, -(X2 * t4), -x23 * t5 + X3 * t4, X2 * p5, -(X2 * p4),
         -x23 * p5 + X3 * p4 - Y3 * (t4 + t5)},
        {0, 0, Y3 * (q4 - q5), X2 * (r5 - 1), X2 * (r4 - 1), -x23 * r5 - X3 * r4 - X2, X2 * q5,
         X2 * q4, -x23 * q5 - X3 * q4 + Y3 * (r4 - r5)},
        {0, 0, Y3 * (r4 - r5), -(X2 * q5), -(X2 * q4), X3 * q4 + x23 * q5, X2 * (r5 - 2),
         X2 * (r4 - 2), -x23 * r5 - X3 * r4 + 4 * X2 + Y3 * (q5 - q4)}};
    memcpy(aT_out, aT, sizeof aT);
}

static void shell_elastic(double k[18][18], double E, double nu, const double *xl, double t, double A,
                          const double *sl)
{   /* stiffe_sh, shell.c:346-485 */
    double km[6][6], kb[9][9];
    cst_membrane(km, E, nu, xl, t, A);
    dkt_bending(kb, E, nu, xl, t, A, sl);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) k[FM[i]][FM[j]] = km[i][j];
    for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) k[FB[i]][FB[j]] = kb[i][j];
    k[5][5] = kb[1][1] / 10000; k[11][11] = kb[4][4] / 10000; k[17][17] = kb[7][7] / 10000;
}

static void local_membrane_coords(double *out3, const double *x, long j, long k, long l,
                                  const double *c1, const double *c2)
{   /* mem_coord, shell.c:2402-2446: x2, x3, y3 of the triangle in the element frame */
    double Xk[3], Xl[3];
    for (int m = 0; m < 3; ++m) { Xk[m] = x[k * 3 + m] - x[j * 3 + m]; Xl[m] = x[l * 3 + m] - x[j * 3 + m]; }
    out3[0] = dotn(c1, Xk, 3);
    out3[1] = dotn(c1, Xl, 3);
    out3[2] = dotn(c2, Xl, 3);
}

static void shell_geometric(double k[18][18], double E, double nu, const double *xl, double t,
                            double Adef, const double *dm)
{   /* stiffg_sh, shell.c:660-840 */
    double C[3][3], Bm[3][6], CB[3][6], Nm[3];
    plane_stress(E, nu, C);
    membrane_B(xl, Adef, Bm);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0;
            for (int q = 0; q < 3; ++q) s += C[i][q] * Bm[q][j];
            CB[i][j] = s;
        }
    for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int j = 0; j < 6; ++j) s += CB[i][j] * dm[j];
        Nm[i] = t * s;
    }
    double N[6][6], Bnl[6][9], BN[9][6];
    memset(N, 0, sizeof N); memset(Bnl, 0, sizeof Bnl);
    for (int c = 0; c < 3; ++c) {
        N[2 * c][2 * c] = Nm[0]; N[2 * c][2 * c + 1] = Nm[2];
        N[2 * c + 1][2 * c] = Nm[2]; N[2 * c + 1][2 * c + 1] = Nm[1];
        /* rows 2c, 2c+1: d/dx and d/dy of displacement component c (shell.c:718-735) */
        Bnl[2 * c][c] = -(xl[2] / (2 * Adef));         Bnl[2 * c][3 + c] = xl[2] / (2 * Adef);
        Bnl[2 * c + 1][c] = (xl[1] - xl[0]) / (2 * Adef);
        Bnl[2 * c + 1][3 + c] = -(xl[1] / (2 * Adef)); Bnl[2 * c + 1][6 + c] = xl[0] / (2 * Adef);
    }
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0;
            for (int q = 0; q < 6; ++q) s += Bnl[q][i] * N[q][j];
            BN[i][j] = s;
        }
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) {
            double s = 0;
            for (int q = 0; q < 6; ++q) s += BN[i][q] * Bnl[q][j];
            k[FG[i]][FG[j]] += Adef * s;
        }
}

/* ---- Ivanov yield criterion in stress resultants, ANAFLAG 3 ---------------------------- */
typedef struct { double alpha, Me, Nbar, Mbar, MNbar, q, r, s, phi; int h; } ivanov;

static void ivanov_eval(ivanov *v, const double *N, const double *M, double chi, double fy, double t,
                        double No)
{   /* shell.c:178-234 (the same block at 1833-1876, 1996-2041 and 2145-2188).  r, s, h keep their
     * previous values when q < 1e-4, as the reference's per-vertex arrays do. */
    v->alpha = 1.0 - 0.4 * exp(-2.6 * sqrt(chi));
    v->Me = v->alpha * 0.25 * fy * pow(t, 2);
    v->Nbar = pow(N[0], 2) + pow(N[1], 2) - N[0] * N[1] + 3 * pow(N[2], 2);
    v->Mbar = pow(M[0], 2) + pow(M[1], 2) - M[0] * M[1] + 3 * pow(M[2], 2);
    v->MNbar = M[0] * N[0] + M[1] * N[1] - 0.5 * M[0] * N[1] - 0.5 * M[1] * N[0] + 3 * M[2] * N[2];
    v->q = v->Nbar * pow(v->Me, 2) + 0.48 * v->Mbar * pow(No, 2);
    if (v->q >= 1e-4) {
        v->r = sqrt(pow(No, 2) * pow(v->Mbar, 2) + 4 * pow(v->Me, 2) * pow(v->MNbar, 2));
        v->h = (v->r / (2 * pow(v->Me, 2) * No) >= 1e-4) ? 1 : 0;
        v->s = v->Nbar * v->Mbar - pow(v->MNbar, 2);
        if (v->h == 1)
            v->phi = v->Nbar / pow(No, 2) + 0.5 * v->Mbar / pow(v->Me, 2) - 0.25 * v->s / v->q +
                     v->r / (2 * pow(v->Me, 2) * No);
        else
            v->phi = v->Nbar / pow(No, 2) + 0.5 * v->Mbar / pow(v->Me, 2) - 0.25 * v->s / v->q;
    }
}

typedef struct { double fn[3], fm[3], fnC[3], fmC[3], jf, kf, Bf, df_da, da_dchi; } ivflow;

static void ivanov_flow(ivflow *f, const ivanov *v, const double *N, const double *M, double C[3][3],
                        double E, double t, double fy, double chi, double No)
{   /* plastic flow directions and the factors of the elasto-plastic moduli, shell.c:868-937
     * (the same block at 1880-1960 and 2052-2122) */
    const double c_fact = 1 / pow(No, 2) - v->Mbar / (4 * v->q) + v->s * pow(v->Me, 2) / (4 * pow(v->q, 2));
    double g_fact, d_fact;
    if (v->h == 1) {
        g_fact = v->MNbar * (1 / (4 * v->q) + 1 / (No * v->r));
        d_fact = 1 / (2 * pow(v->Me, 2)) - v->Nbar / (4 * v->q) + 0.12 * pow(No, 2) * v->s / pow(v->q, 2) +
                 v->Mbar * No / (2 * pow(v->Me, 2) * v->r);
    } else {
        g_fact = v->MNbar / (4 * v->q);
        d_fact = 1 / (2 * pow(v->Me, 2)) - v->Nbar / (4 * v->q) + 0.12 * pow(No, 2) * v->s / pow(v->q, 2);
    }
    const double gN[3] = {2 * N[0] - N[1], 2 * N[1] - N[0], 6 * N[2]};
    const double gM[3] = {2 * M[0] - M[1], 2 * M[1] - M[0], 6 * M[2]};
    for (int j = 0; j < 3; ++j) {
        f->fn[j] = c_fact * gN[j] + g_fact * gM[j];
        f->fm[j] = g_fact * gN[j] + d_fact * gM[j];
    }
    for (int j = 0; j < 3; ++j) {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 3; ++k) { s1 += f->fn[k] * C[k][j]; s2 += f->fm[k] * C[k][j]; }
        f->fnC[j] = s1; f->fmC[j] = s2;
    }
    double s1 = 0, s2 = 0;
    for (int j = 0; j < 3; ++j) { s1 += f->fnC[j] * f->fn[j]; s2 += f->fmC[j] * f->fm[j]; }
    f->jf = t * s1;
    f->kf = pow(t, 3) * s2 / 12;
    f->Bf = 2 * sqrt(pow(g_fact, 2) * v->Nbar + pow(d_fact, 2) * v->Mbar + 2 * d_fact * g_fact * v->MNbar);
    if (v->h == 1)
        f->df_da = -(v->Mbar / (v->alpha * pow(v->Me, 2))) +
                   v->s * v->Nbar * pow(v->Me, 2) / (2 * pow(v->q, 2) * v->alpha) -
                   v->r / (v->alpha * pow(v->Me, 2) * No) + 2 * pow(v->MNbar, 2) / (v->alpha * No * v->r);
    else
        f->df_da = -(v->Mbar / (v->alpha * pow(v->Me, 2))) +
                   v->s * v->Nbar * pow(v->Me, 2) / (2 * pow(v->q, 2) * v->alpha);
    if (chi >= 1e-6) f->da_dchi = 0.52 * E * t * exp(-2.6 * sqrt(chi)) / (3 * fy * sqrt(chi));
    else f->da_dchi = 0;
}

static void shell_plastic(double k[18][18], const ivflow *f, double C[3][3], const double *xl, double t,
                          double Adef, const double *dsl)
{   /* stiffm_sh + stiffm_m_sh + stiffm_b_sh + stiffm_mb_sh, shell.c:842-1503: elasto-plastic
     * membrane, bending and coupling blocks at the controlling yielded vertex (entries are
     * assigned, the geometric part is added afterwards by the caller) */
    const double den = f->jf + f->kf - f->Bf * f->df_da * f->da_dchi;
    double Bm[3][6], aT[9][9];
    membrane_B(xl, Adef, Bm);
    dkt_alpha_T(aT, xl, dsl);
    {   /* membrane, shell.c:1134-1199 */
        const double xi = t / den;
        double xNC[3][3], Cs[3][3], BC[6][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += (f->fn[i] * f->fn[q]) * C[q][j];
                xNC[i][j] = xi * s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += C[i][q] * xNC[q][j];
                Cs[i][j] = t * (C[i][j] - s);
            }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += Bm[q][i] * Cs[q][j];
                BC[i][j] = s;
            }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += BC[i][q] * Bm[q][j];
                k[FM[i]][FM[j]] = Adef * s;
            }
    }
    {   /* bending, shell.c:1201-1368 */
        const double xi = pow(t, 3) / (12 * den);
        double xMC[3][3], Ds[3][3], Q[9][9], kb[9][9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += (f->fm[i] * f->fm[q]) * C[q][j];
                xMC[i][j] = xi * s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += C[i][q] * xMC[q][j];
                Ds[i][j] = pow(t, 3) * (C[i][j] - s) / 12;
            }
        for (int i = 0; i < 9; ++i) {    /* shell.c:1314-1357: the three row blocks are alike */
            double b[3];
            for (int j = 0; j < 3; ++j) {
                b[j] = 0;
                for (int q = 0; q < 3; ++q)
                    b[j] += Ds[j][0] * aT[i][q] + Ds[j][1] * aT[i][q + 3] + Ds[j][2] * aT[i][q + 6];
            }
            for (int j = 0; j < 3; ++j) {
                Q[i][j] = (Ds[0][0] * aT[i][j] + Ds[1][0] * aT[i][j + 3] + Ds[2][0] * aT[i][j + 6] + b[0]) / 24;
                Q[i][j + 3] = (Ds[0][1] * aT[i][j] + Ds[1][1] * aT[i][j + 3] + Ds[2][1] * aT[i][j + 6] + b[1]) / 24;
                Q[i][j + 6] = (Ds[0][2] * aT[i][j] + Ds[1][2] * aT[i][j + 3] + Ds[2][2] * aT[i][j + 6] + b[2]) / 24;
            }
        }
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 9; ++j) {
                double s = 0;
                for (int q = 0; q < 9; ++q) s += Q[i][q] * aT[j][q];
                kb[i][j] = s / (2 * Adef);
            }
        for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) k[FB[i]][FB[j]] = kb[i][j];
        k[5][5] = kb[1][1] / 10000; k[11][11] = kb[4][4] / 10000; k[17][17] = kb[7][7] / 10000;
    }
    {   /* membrane-bending coupling, shell.c:1370-1503 */
        const double xi = -pow(t, 4) / (12 * den);
        double xNMC[3][3], cd[3][3], cdL[3][9], BcL[6][9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += (f->fn[i] * f->fm[q]) * C[q][j];
                xNMC[i][j] = xi * s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += C[i][q] * xNMC[q][j];
                cd[i][j] = s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) cdL[i][j * 3] = cdL[i][j * 3 + 1] = cdL[i][j * 3 + 2] = cd[i][j] / 6;
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 9; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += Bm[q][i] * cdL[q][j];
                BcL[i][j] = s;
            }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 9; ++j) {
                double s = 0;
                for (int q = 0; q < 9; ++q) s += BcL[i][q] * aT[j][q];
                k[FM[i]][FB[j]] = k[FB[j]][FM[i]] = s;
            }
    }
}

static void strain_curvature(double *strn, double curv[3][3], const double *ddm, const double *ddb,
                             const double *xl, double Adef, const double *dsl)
{   /* strn_curv, shell.c:2448-2553: membrane strain increment and the curvature increments at
     * the three vertices */
    double Bm[3][6], aT[9][9];
    membrane_B(xl, Adef, Bm);
    for (int i = 0; i < 3; ++i) strn[i] = dotn(Bm[i], ddm, 6);
    dkt_alpha_T(aT, xl, dsl);
    for (int v = 0; v < 3; ++v)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int q = 0; q < 9; ++q) s += aT[q][3 * j + v] * ddb[q];
            curv[v][j] = s / (2 * Adef);
        }
}

/* the yielded vertex that controls the element: the one with the smallest phi among those on
 * the surface (shell.c:214-222) */
static void pick_vertex(int *yv, const double *phi, int i)
{
    if (*yv == 0) *yv = i + 1;
    else if (phi[i] < phi[*yv - 1]) *yv = i + 1;
}

static void triad18(double T[18][18], const double *c1, const double *c2, const double *c3)
{   /* shell.c:285-302 */
    memset(T, 0, 18 * 18 * sizeof(double));
    for (int g = 0; g < 6; ++g)
        for (int m = 0; m < 3; ++m) {
            T[3 * g][3 * g + m] = c1[m]; T[3 * g + 1][3 * g + m] = c2[m]; T[3 * g + 2][3 * g + m] = c3[m];
        }
}

void orc_shell_element_K(const orc_dims *D, long n, double *K18, const double *emod, const double *nu,
                         const double *x_temp, const double *xlocal, const double *thick,
                         const double *farea, const double *deffarea_ip, const double *slength,
                         const double *c1_ip, const double *c2_ip, const double *c3_ip,
                         const long *minc)
{   /* body of the element loop of stiff_sh, shell.c:140-305, ANAFLAG 1 / 2 */
    const long pe = D->NE_TR + D->NE_FR, pm = 2 * D->NE_TR + 2 * D->NE_FR, pc = D->NE_TR + 3 * D->NE_FR;
    double k[18][18], T[18][18];
    memset(k, 0, sizeof k);
    int yv = 0;
    if (D->ANAFLAG == 3) {                             /* shell.c:171-235 */
        const double fy = g_pl.yield[pe + n], No = fy * thick[n];
        ivanov iv[3]; double phi[3];
        memset(iv, 0, sizeof iv);
        for (int i = 0; i < 3; ++i) {
            const double *N = g_pl.efN_temp + n * 9 + i * 3, *M = g_pl.efM_temp + n * 9 + i * 3;
            ivanov_eval(&iv[i], N, M, g_pl.chi_temp[n * 3 + i], fy, thick[n], No);
            phi[i] = (iv[i].q >= 1e-4) ? iv[i].phi : 0;
            if (iv[i].q >= 1e-4 && phi[i] >= 1 - PHITOL) pick_vertex(&yv, phi, i);
        }
        if (yv != 0) {                                 /* shell.c:255-262 */
            const int v = yv - 1;
            double C[3][3]; ivflow fl;
            plane_stress(emod[pe + n], nu[n], C);
            ivanov_flow(&fl, &iv[v], g_pl.efN_temp + n * 9 + v * 3, g_pl.efM_temp + n * 9 + v * 3, C,
                        emod[pe + n], thick[n], fy, g_pl.chi_temp[n * 3 + v], No);
            shell_plastic(k, &fl, C, xlocal + n * 3, thick[n], deffarea_ip[n], g_pl.defslen_ip + n * 3);
        }
    }
    if (yv == 0)
        shell_elastic(k, emod[pe + n], nu[n], xlocal + n * 3, thick[n], farea[n], slength + n * 3);
    if (D->ANAFLAG == 2 || D->ANAFLAG == 3) {
        double cur[3], dm[6] = {0, 0, 0, 0, 0, 0};
        local_membrane_coords(cur, x_temp, minc[pm + n * 3] - 1, minc[pm + n * 3 + 1] - 1,
                              minc[pm + n * 3 + 2] - 1, c1_ip + pc + n * 3, c2_ip + pc + n * 3);
        dm[5] = cur[2] - xlocal[n * 3 + 2];      /* shell.c:164-167 */
        dm[4] = cur[1] - xlocal[n * 3 + 1];
        dm[2] = cur[0] - xlocal[n * 3];
        shell_geometric(k, emod[pe + n], nu[n], xlocal + n * 3, thick[n], deffarea_ip[n], dm);
    }
    triad18(T, c1_ip + pc + n * 3, c2_ip + pc + n * 3, c3_ip + pc + n * 3);
    triple(&k[0][0], &T[0][0], K18, 18);
}

void orc_stiff_sh(const orc_dims *D, double *ss, const double *emod, const double *nu,
                  const double *x_temp, const double *xlocal, const double *thick,
                  const double *farea, const double *deffarea_ip, const double *slength,
                  const double *c1_ip, const double *c2_ip, const double *c3_ip, const long *maxa,
                  const long *minc, const long *mcode)
{
    const long pmc = 6 * D->NE_TR + 14 * D->NE_FR;
    double K[18 * 18];
    for (long n = 0; n < D->NE_SH; ++n) {
        orc_shell_element_K(D, n, K, emod, nu, x_temp, xlocal, thick, farea, deffarea_ip, slength,
                            c1_ip, c2_ip, c3_ip, minc);
        scatter(D, ss, K, 18, mcode + pmc + n * 18, maxa);
    }
}

int orc_forces_sh(const orc_dims *D, double *f_temp, double *ef_ip, double *ef_i, const double *dd,
                   const double *d_temp, const double *x_temp, const double *emod, const double *nu,
                   const double *xlocal, const double *thick, const double *farea,
                   const double *slength, const double *c1_ip, const double *c2_ip,
                   const double *c3_ip, const double *c1_i, const double *c2_i, const double *c3_i,
                   const long *minc, const long *mcode)
{   /* forces_sh, shell.c:1593-2400 (ANAFLAG 3: extra arrays from orc_set_plastic[_sh]) */
    const long pe = D->NE_TR + D->NE_FR, pm = 2 * D->NE_TR + 2 * D->NE_FR, pc = D->NE_TR + 3 * D->NE_FR;
    const long pmc = 6 * D->NE_TR + 14 * D->NE_FR, pef = 2 * D->NE_TR + 14 * D->NE_FR;
    for (long n = 0; n < D->NE_SH; ++n) {
        double Tp[18][18], Ti[18][18], eft[18], def[18], V[18];
        const long *mc = mcode + pmc + n * 18;
        double *efi = ef_i + pef + n * 18, *efp = ef_ip + pef + n * 18;
        memset(eft, 0, sizeof eft); memset(def, 0, sizeof def);
        triad18(Tp, c1_ip + pc + n * 3, c2_ip + pc + n * 3, c3_ip + pc + n * 3);
        triad18(Ti, c1_i + pc + n * 3, c2_i + pc + n * 3, c3_i + pc + n * 3);
        if (D->ANAFLAG == 1) {                         /* shell.c:1695-1727 */
            double k[18][18], dl[18];
            memset(k, 0, sizeof k);
            shell_elastic(k, emod[pe + n], nu[n], xlocal + n * 3, thick[n], farea[n], slength + n * 3);
            for (int i = 0; i < 18; ++i) V[i] = mc[i] ? d_temp[mc[i] - 1] : 0.0;
            for (int i = 0; i < 18; ++i) dl[i] = dotn(Tp[i], V, 18);
            for (int i = 0; i < 18; ++i) efi[i] = dotn(k[i], dl, 18);
        } else {                                       /* shell.c:1728-1785, 2326-2348 */
            double km[6][6], kb[9][9], cur[3], dm[6] = {0, 0, 0, 0, 0, 0}, ddb[9];
            int yv = 0;
            if (D->ANAFLAG == 3) {                     /* shell.c:1786-2213 */
                const double E = emod[pe + n], t = thick[n], fy = g_pl.yield[pe + n], No = fy * t;
                const long j0 = minc[pm + n * 3] - 1, k0 = minc[pm + n * 3 + 1] - 1, l0 = minc[pm + n * 3 + 2] - 1;
                double xi_[3], xp_[3], ddm[6] = {0, 0, 0, 0, 0, 0}, strn[3], curv[3][3], C[3][3], phi[3];
                local_membrane_coords(xi_, x_temp, j0, k0, l0, c1_i + pc + n * 3, c2_i + pc + n * 3);
                local_membrane_coords(xp_, g_pl.x_ip, j0, k0, l0, c1_ip + pc + n * 3, c2_ip + pc + n * 3);
                ddm[2] = xi_[0] - xp_[0]; ddm[4] = xi_[1] - xp_[1]; ddm[5] = xi_[2] - xp_[2];
                for (int i = 0; i < 18; ++i) V[i] = mc[i] ? dd[mc[i] - 1] : 0.0;
                for (int i = 0; i < 9; ++i) ddb[i] = dotn(Tp[FB[i]], V, 18);
                strain_curvature(strn, curv, ddm, ddb, xlocal + n * 3, g_pl.deffarea_ip[n],
                                 g_pl.defslen_ip + n * 3);
                plane_stress(E, nu[n], C);
                ivanov iv[3]; memset(iv, 0, sizeof iv);
                for (int i = 0; i < 3; ++i) {
                    double *N = g_pl.efN_temp + n * 9 + i * 3, *M = g_pl.efM_temp + n * 9 + i * 3;
                    double *chi = g_pl.chi_temp + n * 3 + i;
                    ivanov_eval(&iv[i], N, M, *chi, fy, t, No);
                    if (iv[i].q >= 1e-4 && iv[i].phi >= 1 - PHITOL) {          /* shell.c:1877-1992 */
                        ivflow fl;
                        ivanov_flow(&fl, &iv[i], N, M, C, E, t, fy, *chi, No);
                        double fs = 0, fc = 0;
                        for (int j = 0; j < 3; ++j) { fs += fl.fnC[j] * strn[j]; fc += fl.fmC[j] * curv[i][j]; }
                        fs *= t; fc *= pow(t, 3) / 12;
                        const double lambda = (fs + fc) / (fl.jf + fl.kf - fl.Bf * fl.df_da * fl.da_dchi);
                        *chi += sqrt(pow((E * t) / (3 * fy), 2) * pow(fl.Bf * lambda, 2));
                        for (int j = 0; j < 3; ++j) {
                            double s1 = 0, s2 = 0;
                            for (int q = 0; q < 3; ++q) {
                                s1 += C[j][q] * (strn[q] - lambda * fl.fn[q]);
                                s2 += C[j][q] * (curv[i][q] - lambda * fl.fm[q]);
                            }
                            N[j] += t * s1; M[j] += pow(t, 3) * s2 / 12;
                        }
                    } else {                                                 /* elastic increment */
                        for (int j = 0; j < 3; ++j) {
                            double s1 = 0, s2 = 0;
                            for (int q = 0; q < 3; ++q) { s1 += C[j][q] * strn[q]; s2 += C[j][q] * curv[i][q]; }
                            N[j] += t * s1; M[j] += pow(t, 3) * s2 / 12;
                        }
                    }
                    ivanov_eval(&iv[i], N, M, *chi, fy, t, No);                /* shell.c:1994-2041 */
                    if (iv[i].q >= 1e-4) {
                        phi[i] = iv[i].phi;
                        if (phi[i] > 1 + 10 * PHITOL) return 1;
                        else if (phi[i] > 1 + PHITOL) {                       /* return to the surface */
                            do {
                                ivflow fl;
                                ivanov_flow(&fl, &iv[i], N, M, C, E, t, fy, *chi, No);
                                const double lambda = (phi[i] - 1) / (fl.jf + fl.kf - fl.Bf * fl.df_da * fl.da_dchi);
                                for (int j = 0; j < 3; ++j) {
                                    double s1 = 0, s2 = 0;
                                    for (int q = 0; q < 3; ++q) {
                                        s1 += C[j][q] * (-lambda * fl.fn[q]);
                                        s2 += C[j][q] * (-lambda * fl.fm[q]);
                                    }
                                    N[j] += t * s1; M[j] += pow(t, 3) * s2 / 12;
                                }
                                ivanov_eval(&iv[i], N, M, *chi, fy, t, No);
                                if (iv[i].q >= 1e-4) phi[i] = iv[i].phi;
                            } while (phi[i] > 1 + PHITOL);
                            pick_vertex(&yv, phi, i);
                        } else if (phi[i] >= 1 - PHITOL) pick_vertex(&yv, phi, i);
                    } else phi[i] = 0;
                }
                if (yv != 0) {                         /* shell.c:2279-2306, 2371-2384 */
                    const int v = yv - 1;
                    double k[18][18], dl[18], M18[18][18];
                    ivflow fl;
                    memset(k, 0, sizeof k);
                    ivanov_flow(&fl, &iv[v], g_pl.efN_temp + n * 9 + v * 3, g_pl.efM_temp + n * 9 + v * 3, C,
                                E, t, fy, g_pl.chi_temp[n * 3 + v], No);
                    shell_plastic(k, &fl, C, xlocal + n * 3, t, g_pl.deffarea_ip[n], g_pl.defslen_ip + n * 3);
                    for (int i = 0; i < 18; ++i) dl[i] = dotn(Tp[i], V, 18);
                    for (int i = 0; i < 18; ++i) def[i] = dotn(k[i], dl, 18);
                    for (int i = 0; i < 18; ++i)
                        for (int j = 0; j < 18; ++j) M18[i][j] = dotn(Ti[i], Tp[j], 18);
                    for (int i = 0; i < 18; ++i) {
                        double s = 0;
                        for (int j = 0; j < 18; ++j) s += M18[i][j] * (def[j] + efp[j]);
                        efi[i] = s;
                    }
                }
            }
            if (yv == 0) {
            cst_membrane(km, emod[pe + n], nu[n], xlocal + n * 3, thick[n], farea[n]);
            dkt_bending(kb, emod[pe + n], nu[n], xlocal + n * 3, thick[n], farea[n], slength + n * 3);
            local_membrane_coords(cur, x_temp, minc[pm + n * 3] - 1, minc[pm + n * 3 + 1] - 1,
                                  minc[pm + n * 3 + 2] - 1, c1_i + pc + n * 3, c2_i + pc + n * 3);
            dm[5] = cur[2] - xlocal[n * 3 + 2];
            dm[4] = cur[1] - xlocal[n * 3 + 1];
            dm[2] = cur[0] - xlocal[n * 3];
            for (int i = 0; i < 18; ++i) V[i] = mc[i] ? dd[mc[i] - 1] : 0.0;
            for (int i = 0; i < 9; ++i) ddb[i] = dotn(Tp[FB[i]], V, 18);
            for (int i = 0; i < 6; ++i) {
                efp[FM[i]] = 0;                        /* side effect on the input, shell.c:1773 */
                eft[FM[i]] = dotn(km[i], dm, 6);
            }
            for (int i = 0; i < 9; ++i) def[FB[i]] = dotn(kb[i], ddb, 9);
            double M[18][18];
            for (int i = 0; i < 18; ++i)
                for (int j = 0; j < 18; ++j) M[i][j] = dotn(Ti[i], Tp[j], 18);
            for (int i = 0; i < 18; ++i) {
                double s = 0;
                for (int j = 0; j < 18; ++j) s += M[i][j] * (def[j] + efp[j]);
                efi[i] = eft[i] + s;
            }
            }
        }
        for (int i = 0; i < 18; ++i) {                 /* shell.c:2388-2397 */
            double s = 0;
            for (int j = 0; j < 18; ++j) s += Ti[j][i] * efi[j];
            if (mc[i] != 0) f_temp[mc[i] - 1] += s;
        }
    }
    return 0;
}

void orc_mass_sh(const orc_dims *D, double *sm, const double *dens, const double *thick, double *farea,
                 double *slength, const double *x, const long *minc, const long *mcode)
{   /* mass_sh, shell.c:1505-1591, SLVFLAG 0 (diagonal) layout.  The reference indexes mcode and
     * dens without the element-type offsets; kept as written. */
    const long pm = 2 * D->NE_TR + 2 * D->NE_FR;
    for (long i = 0; i < D->NE_SH; ++i) {
        shell_frame(x, minc[pm + i * 3] - 1, minc[pm + i * 3 + 1] - 1, minc[pm + i * 3 + 2] - 1,
                    slength + i * 3, farea + i, 0, 0, 0);
        const double Mtot = dens[i] * farea[i] * thick[i];
        for (int ie = 0; ie < 18; ++ie) {
            long j = mcode[i * 18 + ie];
            double mv = (ie % 6 < 3) ? Mtot / 3 : Mtot / 3 * (thick[i] * thick[i]) / 12;
            if (j == 0) continue;
            if (D->SLVFLAG == 0) sm[j - 1] += mv;
            else sm[(j - 1) * D->NEQ + j - 1] += mv;   /* full-order layout, shell.c:1576-1588 */
        }
    }
}

/* --------------------------------------------------------------------------------- truss.c */
void orc_stiff_tr(const orc_dims *D, double *ss, const double *emod, const double *carea,
                  const double *llength, const double *defllen_ip, const double *c1_ip,
                  const double *c2_ip, const double *c3_ip, const double *ef_ip, const long *maxa,
                  const long *mcode)
{   /* stiff_tr, truss.c:82-204 */
    for (long n = 0; n < D->NE_TR; ++n) {
        double k2[2][2], T[2][6], Tk[6][2], K[36], Py = 1;
        if (D->ANAFLAG == 1) {
            k2[0][0] = k2[1][1] = emod[n] * carea[n] / llength[n];
            k2[0][1] = k2[1][0] = -(emod[n] * carea[n] / llength[n]);
        } else {
            k2[0][0] = k2[1][1] = emod[n] * carea[n] * (defllen_ip[n] * defllen_ip[n]) / pow(llength[n], 3);
            k2[0][1] = k2[1][0] = -(emod[n] * carea[n] * (defllen_ip[n] * defllen_ip[n]) / pow(llength[n], 3));
            if (D->ANAFLAG == 3) {                      /* stiffm_tr, truss.c:206-229 */
                Py = carea[n] * g_pl.yield[n];
                if (pow(ef_ip[n * 2] / Py, 2) > 1 + PHITOL) {
                    const double G[2] = {2 * ef_ip[n * 2] / pow(Py, 2), 2 * ef_ip[n * 2 + 1] / pow(Py, 2)};
                    const double kG[2] = {k2[0][0] * G[0] + k2[0][1] * G[1], k2[1][0] * G[0] + k2[1][1] * G[1]};
                    const double GkG = kG[0] * G[0] + kG[1] * G[1];
                    k2[0][0] -= pow(kG[0], 2) / GkG; k2[0][1] -= kG[0] * kG[1] / GkG;
                    k2[1][0] -= kG[1] * kG[0] / GkG; k2[1][1] -= pow(kG[1], 2) / GkG;
                }
            }
        }
        memset(T, 0, sizeof T);
        T[0][0] = T[1][3] = c1_ip[n]; T[0][1] = T[1][4] = c2_ip[n]; T[0][2] = T[1][5] = c3_ip[n];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 2; ++j) {
                double s = 0;
                for (int q = 0; q < 2; ++q) s += T[q][i] * k2[q][j];
                Tk[i][j] = s;
            }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                double s = 0;
                for (int q = 0; q < 2; ++q) s += Tk[i][q] * T[q][j];
                K[i * 6 + j] = s;
            }
        if (D->ANAFLAG == 2 ||                         /* truss.c:155-166 */
            (D->ANAFLAG == 3 && pow(ef_ip[n * 2] / Py, 2) >= 1 - PHITOL)) {
            const double g = ef_ip[n * 2] / defllen_ip[n];
            for (int i = 0; i < 6; ++i) K[i * 6 + i] += g;
            for (int i = 0; i < 3; ++i) { K[i * 6 + i + 3] -= g; K[(i + 3) * 6 + i] -= g; }
        }
        scatter(D, ss, K, 6, mcode + n * 6, maxa);
    }
}

void orc_forces_tr(const orc_dims *D, double *f_temp, double *ef_i, const double *d, const double *emod,
                   const double *carea, const double *llength, const double *defllen_i,
                   const double *c1_i, const double *c2_i, const double *c3_i, const long *mcode)
{   /* forces_tr, truss.c:231-378 */
    for (long n = 0; n < D->NE_TR; ++n) {
        const long *mc = mcode + n * 6;
        const double c[3] = {c1_i[n], c2_i[n], c3_i[n]};
        if (D->ANAFLAG == 1) {
            double Dg[6], dl[2], T[2][6], kk[2][2];
            for (int i = 0; i < 6; ++i) Dg[i] = mc[i] ? d[mc[i] - 1] : 0.0;
            memset(T, 0, sizeof T);
            T[0][0] = T[1][3] = c[0]; T[0][1] = T[1][4] = c[1]; T[0][2] = T[1][5] = c[2];
            for (int i = 0; i < 2; ++i) dl[i] = dotn(T[i], Dg, 6);
            kk[0][0] = kk[1][1] = carea[n] * emod[n] / llength[n];
            kk[0][1] = kk[1][0] = -(carea[n] * emod[n] / llength[n]);
            for (int i = 0; i < 2; ++i) ef_i[n * 2 + i] = dotn(kk[i], dl, 2);
        } else {
            const double strain = (defllen_i[n] - llength[n]) / llength[n];
            ef_i[n * 2] = emod[n] * carea[n] * (strain + 0.5 * (strain * strain)) * defllen_i[n] / llength[n];
            ef_i[n * 2 + 1] = -ef_i[n * 2];
            if (D->ANAFLAG == 3) {                     /* truss.c:335-347: capped at the squash load */
                const double Py = g_pl.yield[n] * carea[n];
                if (pow(ef_i[n * 2] / Py, 2) >= 1 - PHITOL) {
                    if (ef_i[n * 2] < 0) { ef_i[n * 2] = -Py; ef_i[n * 2 + 1] = Py; }
                    else { ef_i[n * 2] = Py; ef_i[n * 2 + 1] = -Py; }
                }
            }
        }
        for (int j = 0; j < 6; ++j)
            if (mc[j] != 0) f_temp[mc[j] - 1] -= ef_i[n * 2 + j / 3] * c[j % 3];
    }
}

void orc_mass_tr(const orc_dims *D, double *sm, const double *carea, double *llength, const double *dens,
                 const double *x, const long *minc, const long *mcode)
{   /* mass_tr, truss.c:381-441, SLVFLAG 0: consistent mass row-summed onto the diagonal */
    for (long i = 0; i < D->NE_TR; ++i) {
        long j = minc[i * 2] - 1, k = minc[i * 2 + 1] - 1;
        double el[3], m[6][6];
        for (int l = 0; l < 3; ++l) el[l] = x[k * 3 + l] - x[j * 3 + l];
        llength[i] = sqrt(dotn(el, el, 3));
        memset(m, 0, sizeof m);
        for (int a = 0; a < 6; ++a) m[a][a] = (dens[i] * carea[i] * llength[i]) / 3;
        for (int a = 0; a < 3; ++a) m[a + 3][a] = m[a][a + 3] = (dens[i] * carea[i] * llength[i]) / 6;
        for (int ie = 0; ie < 6; ++ie)
            for (int je = 0; je < 6; ++je) {
                long p = mcode[i * 6 + ie], q = mcode[i * 6 + je];
                if (p != 0 && q != 0) sm[p - 1] += m[ie][je];
            }
    }
}

/* --------------------------------------------------------------------------------- frame.c */
static void frame_elastic(double k[14][14], double E, double G, double A, double L, double Iz,
                          double Iy, double J, double Cw)
{   /* stiffe_fr, frame.c:364-408 */
#define SYM(i, j, v) k[i][j] = k[j][i] = (v)
    k[0][0] = k[7][7] = E * A / L;                 SYM(7, 0, -k[0][0]);
    k[1][1] = k[8][8] = 12 * E * Iz / pow(L, 3);   SYM(8, 1, -k[1][1]);
    k[2][2] = k[9][9] = 12 * E * Iy / pow(L, 3);   SYM(9, 2, -k[2][2]);
    k[3][3] = k[10][10] = 6 * G * J / (5 * L) + 12 * E * Cw / pow(L, 3);   SYM(10, 3, -k[3][3]);
    k[5][5] = k[12][12] = 4 * E * Iz / L;
    k[4][4] = k[11][11] = 4 * E * Iy / L;
    k[6][6] = k[13][13] = 2 * G * J * L / 15 + 4 * E * Cw / L;
    SYM(5, 1, 6 * E * Iz / (L * L)); SYM(12, 1, k[5][1]);
    SYM(8, 5, -k[5][1]);             SYM(12, 8, -k[5][1]);
    SYM(9, 4, 6 * E * Iy / (L * L)); SYM(11, 9, k[9][4]);
    SYM(4, 2, -k[9][4]);             SYM(11, 2, -k[9][4]);
    SYM(6, 3, G * J / 10 + 6 * E * Cw / (L * L)); SYM(13, 3, k[6][3]);
    SYM(10, 6, -k[6][3]);            SYM(13, 10, -k[6][3]);
    SYM(12, 5, 2 * E * Iz / L);
    SYM(11, 4, 2 * E * Iy / L);
    SYM(13, 6, -(G * J * L / 30 - 2 * E * Cw / L));
#undef SYM
}

static void frame_geometric(double k[14][14], const double *ef, double L, double A, double J)
{   /* stiffg_fr, frame.c:410-579; ef = total end forces in the previous local frame */
    const double P = ef[7], M4 = ef[4], M5 = ef[5], M10 = ef[10], M11 = ef[11], M12 = ef[12];
#define ADD(i, j, v) do { k[i][j] += (v); k[j][i] += (v); } while (0)
#define SUB(i, j, v) do { k[i][j] -= (v); k[j][i] -= (v); } while (0)
    k[0][0] += P / L; k[7][7] += P / L; SUB(7, 0, P / L);
    k[1][1] += 6 * P / (5 * L); k[8][8] += 6 * P / (5 * L);
    k[2][2] += 6 * P / (5 * L); k[9][9] += 6 * P / (5 * L);
    SUB(8, 1, 6 * P / (5 * L)); SUB(9, 2, 6 * P / (5 * L));
    k[3][3] += 6 * P * J / (5 * A * L); k[10][10] += 6 * P * J / (5 * A * L);
    SUB(10, 3, 6 * P * J / (5 * A * L));
    k[4][4] += 2 * P * L / 15; k[11][11] += 2 * P * L / 15;
    k[5][5] += 2 * P * L / 15; k[12][12] += 2 * P * L / 15;
    k[6][6] += 2 * P * J / (15 * A); k[13][13] += 2 * P * J / (15 * A);
    ADD(3, 1, (11 * M4 - M11) / (10 * L)); SUB(8, 3, (11 * M4 - M11) / (10 * L));
    ADD(4, 1, M10 / L); ADD(5, 2, M10 / L); ADD(11, 8, M10 / L); ADD(12, 9, M10 / L);
    SUB(11, 1, M10 / L); SUB(12, 2, M10 / L); SUB(8, 4, M10 / L); SUB(9, 5, M10 / L);
    ADD(5, 1, P / 10); ADD(12, 1, P / 10); ADD(9, 4, P / 10); ADD(11, 9, P / 10);
    SUB(4, 2, P / 10); SUB(11, 2, P / 10); SUB(8, 5, P / 10); SUB(12, 8, P / 10);
    ADD(6, 1, M4 / 10); SUB(8, 6, M4 / 10);
    ADD(10, 8, (M4 - 11 * M11) / (10 * L)); SUB(10, 1, (M4 - 11 * M11) / (10 * L));
    ADD(13, 8, M11 / 10); SUB(13, 1, M11 / 10);
    ADD(3, 2, (11 * M5 - M12) / (10 * L)); SUB(9, 3, (11 * M5 - M12) / (10 * L));
    ADD(6, 2, M5 / 10); SUB(9, 6, M5 / 10);
    ADD(10, 9, (M5 - 11 * M12) / (10 * L)); SUB(10, 2, (M5 - 11 * M12) / (10 * L));
    ADD(13, 9, M12 / 10); SUB(13, 2, M12 / 10);
    SUB(4, 3, (2 * M5 - M12) / 5); ADD(5, 3, (2 * M4 - M11) / 5);
    ADD(6, 3, P * J / (10 * A)); ADD(13, 3, P * J / (10 * A));
    SUB(10, 6, P * J / (10 * A)); SUB(13, 10, P * J / (10 * A));
    SUB(11, 3, (2 * M5 + M12) / 10); ADD(12, 3, (2 * M4 + M11) / 10);
    SUB(6, 4, (3 * M5 - M12) * L / 30); SUB(10, 4, (M5 + 2 * M12) / 10);
    SUB(11, 4, P * L / 30); SUB(12, 5, P * L / 30);
    ADD(12, 4, M10 / 2); SUB(11, 5, M10 / 2);
    ADD(13, 4, M5 * L / 30);
    ADD(6, 5, (3 * M4 - M11) * L / 30); ADD(10, 5, (M4 + 2 * M11) / 10);
    SUB(13, 5, M4 * L / 30);
    SUB(11, 6, M12 * L / 30); ADD(12, 6, M11 * L / 30);
    SUB(13, 6, P * J / (30 * A));
    ADD(11, 10, (M5 - 2 * M12) / 5); SUB(12, 10, (M4 - 2 * M11) / 5);
    SUB(13, 11, (M5 - 3 * M12) * L / 30); ADD(13, 12, (M4 - 3 * M11) * L / 30);
#undef ADD
#undef SUB
}

static void frame_release(double k[14][14], const int *rel4)
{   /* release, frame.c:798-900: static condensation k -= kG (G^T k G)^-1 G^T k */
    static const int dof[4] = {5, 4, 12, 11};
    int idx[4], r = 0;
    for (int i = 0; i < 4; ++i) if (rel4[i] == 1) idx[r++] = dof[i];
    if (r == 0) return;
    double kG[14][4], GkG[16], W[14][4];
    for (int i = 0; i < 14; ++i)
        for (int j = 0; j < r; ++j) kG[i][j] = 0 + k[i][idx[j]];   /* sum over k of k[i][k]*G[k][j] */
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < r; ++j) GkG[i * r + j] = kG[idx[j]][i];
    if (r == 1) GkG[0] = 1 / GkG[0];
    else if (r == 2) {
        double det = GkG[0] * GkG[3] - GkG[1] * GkG[2], t = GkG[0];
        GkG[0] = GkG[3] / det; GkG[3] = t / det; GkG[1] *= -1 / det; GkG[2] *= -1 / det;
    } else gj_inverse(GkG, r);
    for (int i = 0; i < 14; ++i)
        for (int j = 0; j < r; ++j) {
            double s = 0;
            for (int q = 0; q < r; ++q) s += kG[i][q] * GkG[q * r + j];
            W[i][j] = s;
        }
    for (int i = 0; i < 14; ++i)
        for (int j = 0; j < 14; ++j) {
            double s = 0;
            for (int q = 0; q < r; ++q) s += W[i][q] * kG[j][q];
            k[i][j] -= s;
        }
}

static double fr_phi(double p, double my, double mz)
{   /* yield function, frame.c:617-621 */
    return pow(p, 2) + pow(mz, 2) + pow(my, 4) + 3.5 * pow(p, 2) * pow(mz, 2) +
           3 * pow(p, 6) * pow(my, 2) + 4.5 * pow(mz, 4) * pow(my, 2);
}

static void fr_grad(double p, double my, double mz, double Py, double Mpy, double Mpz, double *g)
{   /* yield-surface gradients w.r.t. axial force, weak- and strong-axis moment, frame.c:625-648 */
    g[0] = 2 * p / Py + 7 * p * pow(mz, 2) / Py + 18 * pow(p, 5) * pow(my, 2) / Py;
    g[1] = 4 * pow(my, 3) / Mpy + 6 * pow(p, 6) * my / Mpy + 9 * pow(mz, 4) * my / Mpy;
    g[2] = 2 * mz / Mpz + 7 * pow(p, 2) * mz / Mpz + 18 * pow(mz, 3) * pow(my, 2) / Mpz;
}

static void frame_plastic(double k[14][14], const double *eft, const int *yld, double Py, double Mpy,
                          double Mpz)
{   /* stiffm_fr, frame.c:581-796: k -= kG (G^T k G)^-1 G^T k */
    const double p[2] = {eft[0] / Py, eft[7] / Py}, my[2] = {eft[4] / Mpy, eft[11] / Mpy};
    const double mz[2] = {eft[5] / Mpz, eft[12] / Mpz};
    const double phi[2] = {fr_phi(p[0], my[0], mz[0]), fr_phi(p[1], my[1], mz[1])};
    double G[14][2], kG[14][2], A[2][2], W[14][2];
    memset(G, 0, sizeof G);
    int nc;
    if (phi[0] >= 1 - PHITOL && phi[1] >= 1 - PHITOL) {
        double g[3];
        if (yld[0] != 2) { fr_grad(p[0], my[0], mz[0], Py, Mpy, Mpz, g); G[0][0] = g[0]; G[4][0] = g[1]; G[5][0] = g[2]; }
        if (yld[1] != 2) { fr_grad(p[1], my[1], mz[1], Py, Mpy, Mpz, g); G[7][1] = g[0]; G[11][1] = g[1]; G[12][1] = g[2]; }
        nc = 2;
    } else if (phi[0] >= 1 - PHITOL && yld[0] != 2) {
        double g[3];
        fr_grad(p[0], my[0], mz[0], Py, Mpy, Mpz, g); G[0][0] = g[0]; G[4][0] = g[1]; G[5][0] = g[2];
        nc = 1;
    } else if (phi[1] >= 1 - PHITOL && yld[1] != 2) {
        double g[3];
        fr_grad(p[1], my[1], mz[1], Py, Mpy, Mpz, g); G[7][0] = g[0]; G[11][0] = g[1]; G[12][0] = g[2];
        nc = 1;
    } else return;
    for (int i = 0; i < 14; ++i)
        for (int j = 0; j < nc; ++j) {
            double s = 0;
            for (int q = 0; q < 14; ++q) s += k[i][q] * G[q][j];
            kG[i][j] = s;
        }
    for (int i = 0; i < nc; ++i)
        for (int j = 0; j < nc; ++j) {
            double s = 0;
            for (int q = 0; q < 14; ++q) s += kG[q][i] * G[q][j];
            A[i][j] = s;
        }
    if (nc == 2) {
        const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0], t = A[0][0];
        A[0][0] = A[1][1] / det; A[1][1] = t / det; A[0][1] *= -1 / det; A[1][0] *= -1 / det;
        for (int i = 0; i < 14; ++i)
            for (int j = 0; j < 2; ++j) {
                double s = 0;
                for (int q = 0; q < 2; ++q) s += kG[i][q] * A[q][j];
                W[i][j] = s;
            }
        for (int i = 0; i < 14; ++i)
            for (int j = 0; j < 14; ++j) {
                double s = 0;
                for (int q = 0; q < 2; ++q) s += W[i][q] * kG[j][q];
                k[i][j] -= s;
            }
    } else {
        A[0][0] = 1 / A[0][0];
        for (int i = 0; i < 14; ++i) W[i][0] = kG[i][0] * A[0][0];
        for (int i = 0; i < 14; ++i)
            for (int j = 0; j < 14; ++j) k[i][j] -= W[i][0] * kG[j][0];
    }
}

static double regula_falsi(double p, double dp, double my, double dmy, double mz, double dmz)
{   /* frame.c:1397-1455.  The loop condition (phi_r <= 1-tol && phi_r >= 1+tol) is never true:
     * exactly one refinement of the secant estimate is made. */
    double tau_u = 1, tau_l = 0;
    double phi_u = fr_phi(p + tau_u * dp, my + tau_u * dmy, mz + tau_u * dmz);
    double phi_l = fr_phi(p + tau_l * dp, my + tau_l * dmy, mz + tau_l * dmz);
    double tau_r = tau_u - (phi_u - 1) * (tau_l - tau_u) / (phi_l - phi_u);
    double phi_r = fr_phi(p + tau_r * dp, my + tau_r * dmy, mz + tau_r * dmz);
    if ((phi_l - 1 > 0 && phi_r - 1 > 0) || (phi_l - 1 < 0 && phi_r - 1 < 0)) { tau_l = tau_r; phi_l = phi_r; }
    else { tau_u = tau_r; phi_u = phi_r; }
    return tau_u - (phi_u - 1) * (tau_l - tau_u) / (phi_l - phi_u);
}

static int frame_unload(const double *phi, const double *p, const double *my, const double *mz,
                        double Py, double Mpy, double Mpz, double k[14][14], const double *dl)
{   /* unload, frame.c:1457-1658: sign of the plastic multipliers (G^T k G)^-1 G^T k dd */
    double G[14][2], Gk[2][14], A[2][2], g[3];
    memset(G, 0, sizeof G);
    int nc, code1;
    if (phi[0] >= 1 - PHITOL && phi[1] >= 1 - PHITOL) {
        fr_grad(p[0], my[0], mz[0], Py, Mpy, Mpz, g); G[0][0] = g[0]; G[4][0] = g[1]; G[5][0] = g[2];
        fr_grad(p[1], my[1], mz[1], Py, Mpy, Mpz, g); G[7][1] = g[0]; G[11][1] = g[1]; G[12][1] = g[2];
        nc = 2; code1 = 0;
    } else if (phi[0] >= 1 - PHITOL) {
        fr_grad(p[0], my[0], mz[0], Py, Mpy, Mpz, g); G[0][0] = g[0]; G[4][0] = g[1]; G[5][0] = g[2];
        nc = 1; code1 = 2;
    } else if (phi[1] >= 1 - PHITOL) {
        fr_grad(p[1], my[1], mz[1], Py, Mpy, Mpz, g); G[7][0] = g[0]; G[11][0] = g[1]; G[12][0] = g[2];
        nc = 1; code1 = 3;
    } else return 0;
    for (int i = 0; i < nc; ++i)
        for (int j = 0; j < 14; ++j) {
            double s = 0;
            for (int q = 0; q < 14; ++q) s += G[q][i] * k[q][j];
            Gk[i][j] = s;
        }
    for (int i = 0; i < nc; ++i)
        for (int j = 0; j < nc; ++j) {
            double s = 0;
            for (int q = 0; q < 14; ++q) s += Gk[i][q] * G[q][j];
            A[i][j] = s;
        }
    if (nc == 2) {
        double lam[2];
        const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0], t = A[0][0];
        A[0][0] = A[1][1] / det; A[1][1] = t / det; A[0][1] *= -1 / det; A[1][0] *= -1 / det;
        for (int i = 0; i < 2; ++i) {
            double s = 0;
            for (int j = 0; j < 14; ++j) {
                double w = 0;
                for (int q = 0; q < 2; ++q) w += A[i][q] * Gk[q][j];
                s += w * dl[j];
            }
            lam[i] = s;
        }
        if (lam[0] < -1e-8 && lam[1] < -1e-8) return 1;
        if (lam[0] < -1e-8) return 2;
        if (lam[1] < -1e-8) return 3;
        return 0;
    }
    A[0][0] = 1 / A[0][0];
    double s = 0;
    for (int j = 0; j < 14; ++j) s += (A[0][0] * Gk[0][j]) * dl[j];
    return (s < -1e-8) ? code1 : 0;
}

static void frame_T(double T[14][14], const double *c1, const double *c2, const double *c3)
{   /* frame.c:286-295 */
    static const int b0[4] = {0, 3, 7, 10};
    memset(T, 0, 14 * 14 * sizeof(double));
    for (int g = 0; g < 4; ++g)
        for (int m = 0; m < 3; ++m) {
            T[b0[g]][b0[g] + m] = c1[m]; T[b0[g] + 1][b0[g] + m] = c2[m]; T[b0[g] + 2][b0[g] + m] = c3[m];
        }
    T[6][6] = T[13][13] = 1;
}

static void rigid_link_T(double Tr[14][14], const double *off)
{   /* TRANSPOSE of the rigid-link matrix as stiff_fr builds it (frame.c:306-320) */
    memset(Tr, 0, 14 * 14 * sizeof(double));
    for (int i = 0; i < 14; ++i) Tr[i][i] = 1;
    Tr[1][3] = -off[2]; Tr[2][3] = off[1]; Tr[0][4] = off[2]; Tr[2][4] = -off[0];
    Tr[0][5] = -off[1]; Tr[1][5] = off[0];
    Tr[8][10] = -off[5]; Tr[9][10] = off[4]; Tr[7][11] = off[5]; Tr[9][11] = -off[3];
    Tr[7][12] = -off[4]; Tr[8][12] = off[3];
}

static void frame_local_k(const orc_dims *D, long n, double k[14][14], double *eftot, const double *emod,
                          const double *gmod, const double *carea, const double *llength,
                          const double *defllen_ip, const double *istrong, const double *iweak,
                          const double *ipolar, const double *iwarp, const double *ef_ip,
                          const double *efFE_ip, const int *mendrel)
{
    const long T0 = D->NE_TR;
    memset(k, 0, 14 * 14 * sizeof(double));
    for (int i = 0; i < 14; ++i) eftot[i] = ef_ip[2 * T0 + n * 14 + i] + efFE_ip[n * 14 + i];
    frame_elastic(k, emod[T0 + n], gmod[n], carea[T0 + n], llength[T0 + n], istrong[n], iweak[n],
                  ipolar[n], iwarp[n]);
    if (D->ANAFLAG == 2 || D->ANAFLAG == 3)
        frame_geometric(k, eftot, defllen_ip[T0 + n], carea[T0 + n], ipolar[n]);
    if (D->ANAFLAG == 3 && (g_pl.yldflag[n * 2] != 2 || g_pl.yldflag[n * 2 + 1] != 2))   /* frame.c:268-277 */
        frame_plastic(k, eftot, g_pl.yldflag + n * 2, carea[T0 + n] * g_pl.yield[T0 + n],
                      g_pl.zweak[n] * g_pl.yield[T0 + n], g_pl.zstrong[n] * g_pl.yield[T0 + n]);
    if (mendrel[n * 5] == 1) frame_release(k, mendrel + n * 5 + 1);
}

void orc_stiff_fr(const orc_dims *D, double *ss, const double *emod, const double *gmod,
                  const double *carea, const double *offset, const int *osflag, const double *llength,
                  const double *defllen_ip, const double *istrong, const double *iweak,
                  const double *ipolar, const double *iwarp, const double *c1_ip, const double *c2_ip,
                  const double *c3_ip, const double *ef_ip, const double *efFE_ip, const int *mendrel,
                  const long *maxa, const long *mcode)
{   /* stiff_fr, frame.c:226-362, ANAFLAG 1 / 2 */
    const long T0 = D->NE_TR;
    for (long n = 0; n < D->NE_FR; ++n) {
        double k[14][14], T[14][14], K[196], K2[196], ef[14];
        frame_local_k(D, n, k, ef, emod, gmod, carea, llength, defllen_ip, istrong, iweak, ipolar,
                      iwarp, ef_ip, efFE_ip, mendrel);
        frame_T(T, c1_ip + T0 + n * 3, c2_ip + T0 + n * 3, c3_ip + T0 + n * 3);
        triple(&k[0][0], &T[0][0], K, 14);
        if (osflag[n] != 0) {
            double Tr[14][14];
            rigid_link_T(Tr, offset + n * 6);
            triple(K, &Tr[0][0], K2, 14);
            scatter(D, ss, K2, 14, mcode + 6 * T0 + n * 14, maxa);
        } else {
            scatter(D, ss, K, 14, mcode + 6 * T0 + n * 14, maxa);
        }
    }
}

int orc_forces_fr(const orc_dims *D, double *f_temp, const double *ef_ip, double *ef_i,
                   const double *efFE_ref, const double *efFE_ip, double *efFE_i, const double *dd,
                   const double *emod, const double *gmod, const double *carea, const double *offset,
                   const int *osflag, const double *llength, const double *defllen_ip,
                   const double *istrong, const double *iweak, const double *ipolar,
                   const double *iwarp, const double *c1_ip, const double *c2_ip, const double *c3_ip,
                   const double *c1_i, const double *c2_i, const double *c3_i, const int *mendrel,
                   const long *mcode, double *pdlpf, int itecnt)
{   /* forces_fr, frame.c:902-1312 (ANAFLAG 3: yldflag etc. from orc_set_plastic).  ef_ip / ef_i and
     * efFE_ip / efFE_i may alias, as in the linear call of main.c:1782 - rows are then updated in
     * place exactly like the reference does. */
    const long T0 = D->NE_TR;
    for (long n = 0; n < D->NE_FR; ++n) {
        double k[14][14], Tp[14][14], Ti[14][14], Tr[14][14], M[14][14];
        double eft[14], DD12[14], DDij[14], dl[14], def[14], EF[14];
        const long *mc = mcode + 6 * T0 + n * 14;
        const double *efp = ef_ip + 2 * T0 + n * 14;
        double *efi = ef_i + 2 * T0 + n * 14;
        frame_local_k(D, n, k, eft, emod, gmod, carea, llength, defllen_ip, istrong, iweak, ipolar,
                      iwarp, ef_ip, efFE_ip, mendrel);
        for (int i = 0; i < 14; ++i) DD12[i] = mc[i] ? dd[mc[i] - 1] : 0.0;
        frame_T(Tp, c1_ip + T0 + n * 3, c2_ip + T0 + n * 3, c3_ip + T0 + n * 3);
        if (osflag[n] == 0) {
            for (int i = 0; i < 14; ++i) dl[i] = dotn(Tp[i], DD12, 14);
        } else {
            /* forces_fr builds T_rl itself (frame.c:1018-1032) = transpose of stiff_fr's matrix */
            double TrT[14][14];
            rigid_link_T(TrT, offset + n * 6);
            for (int i = 0; i < 14; ++i) for (int j = 0; j < 14; ++j) Tr[i][j] = TrT[j][i];
            for (int i = 0; i < 14; ++i) {
                double s = 0;
                for (int j = 0; j < 14; ++j) s += Tr[j][i] * DD12[j];
                DDij[i] = s;
            }
            for (int i = 0; i < 14; ++i) dl[i] = dotn(Tp[i], DDij, 14);
        }
        for (int i = 0; i < 14; ++i) def[i] = dotn(k[i], dl, 14);
        frame_T(Ti, c1_i + T0 + n * 3, c2_i + T0 + n * 3, c3_i + T0 + n * 3);
        for (int i = 0; i < 14; ++i) for (int j = 0; j < 14; ++j) M[i][j] = dotn(Ti[i], Tp[j], 14);
        for (int i = 0; i < 14; ++i) {                 /* frame.c:1090-1096 */
            double s = 0;
            for (int j = 0; j < 14; ++j) s += M[i][j] * (efp[j] + def[j]);
            efi[i] = s;
        }
        const int *yld = (D->ANAFLAG == 3) ? g_pl.yldflag + n * 2 : NULL;
        double eti[14];
        for (int i = 0; i < 14; ++i) {                 /* frame.c:1100-1155 */
            const int yend = yld ? yld[i / 7] : 0;      /* an end that has yielded takes no increment */
            const double dlpf = *pdlpf;
            double s = 0;
            for (int j = 0; j < 14; ++j)
                s += M[i][j] * ((itecnt == 0 && yend == 0) ? (efFE_ip[n * 14 + j] + dlpf * efFE_ref[n * 14 + j])
                                                           : efFE_ip[n * 14 + j]);
            efFE_i[n * 14 + i] = s;
            eti[i] = efi[i] + s;
        }
        if (D->ANAFLAG == 3) {                         /* frame.c:1157-1268 */
            int *y = g_pl.yldflag + n * 2;
            const double Py = carea[T0 + n] * g_pl.yield[T0 + n], Mpy = g_pl.zweak[n] * g_pl.yield[T0 + n],
                         Mpz = g_pl.zstrong[n] * g_pl.yield[T0 + n];
            const double p[2] = {eti[0] / Py, eti[7] / Py}, my[2] = {eti[4] / Mpy, eti[11] / Mpy};
            const double mz[2] = {eti[5] / Mpz, eti[12] / Mpz};
            const double phi[2] = {fr_phi(p[0], my[0], mz[0]), fr_phi(p[1], my[1], mz[1])};
            if (phi[0] > phi[1] && phi[0] > 1 + PHITOL && y[0] != 2) {
                *pdlpf *= regula_falsi(eft[0] / Py, (eti[0] - eft[0]) / Py, eft[4] / Mpy, (eti[4] - eft[4]) / Mpy,
                                       eft[5] / Mpz, (eti[5] - eft[5]) / Mpz);
                y[0] = 1;
                return 1;
            } else if (phi[1] > phi[0] && phi[1] > 1 + PHITOL && y[1] != 2) {
                *pdlpf *= regula_falsi(eft[7] / Py, (eti[7] - eft[7]) / Py, eft[11] / Mpy, (eti[11] - eft[11]) / Mpy,
                                       eft[12] / Mpz, (eti[12] - eft[12]) / Mpz);
                y[1] = 1;
                return 1;
            } else if ((phi[0] >= 1 - PHITOL && y[0] != 2) && (phi[1] >= 1 - PHITOL && y[1] != 2)) {
                y[0] = y[1] = 1;
            } else if (phi[0] >= 1 - PHITOL && y[0] != 2) {
                y[0] = 1;
            } else if (phi[1] >= 1 - PHITOL && y[1] != 2) {
                y[1] = 1;
            }
            if (y[0] == 1 || y[1] == 1) {
                memset(k, 0, sizeof k);
                frame_elastic(k, emod[T0 + n], gmod[n], carea[T0 + n], llength[T0 + n], istrong[n], iweak[n],
                              ipolar[n], iwarp[n]);
                frame_geometric(k, eft, defllen_ip[T0 + n], carea[T0 + n], ipolar[n]);
                const int u = frame_unload(phi, p, my, mz, Py, Mpy, Mpz, k, dl);
                if (u == 1) { y[0] = y[1] = 2; return 2; }
                if (u == 2) { y[0] = 2; return 2; }
                if (u == 3) { y[1] = 2; return 2; }
            }
        }
        for (int i = 0; i < 14; ++i) {                 /* frame.c:1273-1309 */
            double s = 0;
            for (int j = 0; j < 14; ++j) s += Ti[j][i] * efi[j];
            EF[i] = s;
        }
        for (int i = 0; i < 14; ++i) {
            double s = EF[i];
            if (osflag[n] != 0) {
                s = 0;
                for (int j = 0; j < 14; ++j) s += Tr[i][j] * EF[j];
            }
            if (mc[i] != 0) f_temp[mc[i] - 1] += s;
        }
    }
    return 0;
}

void orc_mass_fr(const orc_dims *D, double *sm, const double *carea, double *llength, const double *dens,
                 const int *osflag, const double *offset, const double *x, double *xfr,
                 const long *minc, const long *mcode)
{   /* mass_fr, frame.c:1314-1393, SLVFLAG 0.  Reference indexing quirks kept: mcode+i*14,
     * dens+i, carea+i, llength+i WITHOUT the truss offsets (App. B.4), while the refreshed length
     * is stored at llength[NE_TR+i]. */
    const long pm = D->NE_TR * 2;
    for (long i = 0; i < D->NE_FR; ++i) {
        long j = minc[pm + i * 2] - 1, k = minc[pm + i * 2 + 1] - 1;
        double el[3];
        for (int l = 0; l < 3; ++l) {
            xfr[i * 6 + l] = x[j * 3 + l]; xfr[i * 6 + 3 + l] = x[k * 3 + l];
            if (osflag[i] != 0) {
                xfr[i * 6 + l] = x[j * 3 + l] + offset[i * 6 + l];
                xfr[i * 6 + 3 + l] = x[k * 3 + l] + offset[i * 6 + 3 + l];
            }
        }
        for (int l = 0; l < 3; ++l) el[l] = xfr[i * 6 + 3 + l] - xfr[i * 6 + l];
        llength[D->NE_TR + i] = sqrt(dotn(el, el, 3));
        const double mt = (dens[i] * carea[i] * llength[i]) / 24 * 12;
        const double mr = (dens[i] * carea[i] * llength[i]) / 24 * (llength[i] * llength[i]);
        for (int ie = 0; ie < 14; ++ie) {
            long q = mcode[i * 14 + ie];
            int r = ie % 7;
            if (q != 0 && r < 6) sm[q - 1] += (r < 3) ? mt : mr;
            else if (q != 0) sm[q - 1] += 0.0;
        }
    }
}

/* --------------------------------------------------------------------------------- brick.c */
static const int BG[8] = {+1, -1, +1, -1, -1, +1, -1, +1};   /* overall sign of dh/d(r,s,t)   */
static const int BR[8] = {+1, -1, -1, +1, +1, -1, -1, +1};   /* R/8 + BR/8                    */
static const int BS[8] = {+1, +1, -1, -1, +1, +1, -1, -1};   /* S + BS                        */
static const int BT[8] = {+1, +1, +1, +1, -1, -1, -1, -1};   /* T + BT                        */

void orc_stiff_br(const orc_dims *D, double *ss, const double *x, const double *emod, const double *nu,
                  const long *minc, const long *mcode)
{   /* stiff_br + jacob, brick.c:79-397, 541-699: 2x2x2 Gauss, B^T C B detJ, dense scatter */
    const long pe = D->NE_TR + D->NE_FR + D->NE_SH;
    const long pm = 2 * D->NE_TR + 2 * D->NE_FR + 3 * D->NE_SH;
    const long pmc = 6 * D->NE_TR + 14 * D->NE_FR + 18 * D->NE_SH;
    const double gp[2] = {1.0 / sqrt(3), -1.0 / sqrt(3)};
    for (long e = 0; e < D->NE_BR; ++e) {
        double kbr[24][24], C[6][6];
        const double E = emod[pe + e], v = nu[pe + e];
        const double e1 = E * v / ((1 + v) * (1 - 2 * v));
        const double e2 = .5 * (E / ((1 + v)));
        const double e3 = E * (1 - v) / ((1 + v) * (1 - 2 * v));
        memset(kbr, 0, sizeof kbr); memset(C, 0, sizeof C);
        C[0][0] = C[1][1] = C[2][2] = e3;
        C[0][1] = C[0][2] = C[1][0] = C[1][2] = C[2][0] = C[2][1] = e1;
        C[3][3] = C[4][4] = C[5][5] = e2;
        for (int r = 0; r < 2; ++r) for (int s = 0; s < 2; ++s) for (int t = 0; t < 2; ++t) {
            const double R = gp[r], S = gp[s], T = gp[t];
            double jac[9], Ji[9], B[6][24], CBm[6][24], dh[3][8];
            memset(jac, 0, sizeof jac);
            for (int n = 0; n < 8; ++n) {              /* jacob, brick.c:572-697 */
                const long jt = minc[pm + e * 8 + n] - 1;
                const double cr = (BG[n] * ((S + BS[n]) * (T + BT[n]))) / 8.0;
                const double cs = (BG[n] * (R / 8.0 + BR[n] * (1.0 / 8.0))) * (T + BT[n]);
                const double ct = (BG[n] * (R / 8.0 + BR[n] * (1.0 / 8.0))) * (S + BS[n]);
                for (int m = 0; m < 3; ++m) {
                    jac[0 + m] += cr * x[jt * 3 + m];
                    jac[3 + m] += cs * x[jt * 3 + m];
                    jac[6 + m] += ct * x[jt * 3 + m];
                }
            }
            memcpy(Ji, jac, sizeof Ji);
            gj_inverse(Ji, 3);
            const double detJ = jac[0] * jac[4] * jac[8] - jac[0] * jac[5] * jac[7] - jac[1] * jac[3] * jac[8] +
                                jac[1] * jac[5] * jac[6] + jac[2] * jac[3] * jac[7] - jac[2] * jac[4] * jac[6];
            for (int n = 0; n < 8; ++n)                /* brick.c:196-316 */
                for (int m = 0; m < 3; ++m)
                    dh[m][n] = Ji[m * 3 + 0] * (BG[n] * ((S + BS[n]) * (T + BT[n]))) / 8.0 +
                               Ji[m * 3 + 1] * (BG[n] * (R / 8.0 + BR[n] * (1.0 / 8.0))) * (T + BT[n]) +
                               Ji[m * 3 + 2] * (BG[n] * (R / 8.0 + BR[n] * (1.0 / 8.0))) * (S + BS[n]);
            memset(B, 0, sizeof B);
            for (int n = 0; n < 8; ++n) {              /* brick.c:326-348 */
                B[0][3 * n] = dh[0][n]; B[1][3 * n + 1] = dh[1][n]; B[2][3 * n + 2] = dh[2][n];
                B[3][3 * n] = dh[1][n]; B[3][3 * n + 1] = dh[0][n];
                B[4][3 * n + 1] = dh[2][n]; B[4][3 * n + 2] = dh[1][n];
                B[5][3 * n] = dh[2][n]; B[5][3 * n + 2] = dh[0][n];
            }
            for (int l = 0; l < 24; ++l)
                for (int j = 0; j < 6; ++j) {
                    double sum = 0;
                    for (int q = 0; q < 6; ++q) sum += C[j][q] * B[q][l];
                    CBm[j][l] = sum;
                }
            for (int l = 0; l < 24; ++l)
                for (int j = 0; j < 24; ++j) {
                    double sum = 0;
                    for (int q = 0; q < 6; ++q) sum += B[q][j] * CBm[q][l];
                    kbr[j][l] += sum * detJ;
                }
        }
        for (int ie = 0; ie < 24; ++ie)                /* brick.c:384-395 */
            for (int je = 0; je < 24; ++je) {
                long j = mcode[pmc + e * 24 + ie], k = mcode[pmc + e * 24 + je];
                if (j != 0 && k != 0) ss[(j - 1) * D->NEQ + k - 1] += kbr[je][ie];
            }
    }
}

void orc_mass_br(const orc_dims *D, double *sm, const double *dens, const double *x, const long *minc,
                 const long *mcode)
{   /* mass_br + jacob, brick.c:399-537: consistent mass rho H^T H detJ, full-order scatter */
    const long pe = D->NE_TR + D->NE_FR + D->NE_SH;
    const long pm = 2 * D->NE_TR + 2 * D->NE_FR + 3 * D->NE_SH;
    const long pmc = 6 * D->NE_TR + 14 * D->NE_FR + 18 * D->NE_SH;
    const double gp[2] = {1.0 / sqrt(3), -1.0 / sqrt(3)};
    for (long e = 0; e < D->NE_BR; ++e) {
        double mbr[24][24];
        memset(mbr, 0, sizeof mbr);
        for (int r = 0; r < 2; ++r) for (int s = 0; s < 2; ++s) for (int t = 0; t < 2; ++t) {
            const double R = gp[r], S = gp[s], T = gp[t];
            double jac[9], h[8];
            memset(jac, 0, sizeof jac);
            for (int n = 0; n < 8; ++n) {              /* jacob, brick.c:572-697 */
                const long jt = minc[pm + e * 8 + n] - 1;
                const double cr = (BG[n] * ((S + BS[n]) * (T + BT[n]))) / 8.0;
                const double cs = (BG[n] * (R / 8.0 + BR[n] * (1.0 / 8.0))) * (T + BT[n]);
                const double ct = (BG[n] * (R / 8.0 + BR[n] * (1.0 / 8.0))) * (S + BS[n]);
                for (int m = 0; m < 3; ++m) {
                    jac[0 + m] += cr * x[jt * 3 + m];
                    jac[3 + m] += cs * x[jt * 3 + m];
                    jac[6 + m] += ct * x[jt * 3 + m];
                }
            }
            const double detJ = jac[0] * jac[4] * jac[8] - jac[0] * jac[5] * jac[7] - jac[1] * jac[3] * jac[8] +
                                jac[1] * jac[5] * jac[6] + jac[2] * jac[3] * jac[7] - jac[2] * jac[4] * jac[6];
            for (int n = 0; n < 8; ++n)                /* brick.c:468-475 */
                h[n] = (1.0 + BR[n] * R) * (1.0 + BS[n] * S) * (1.0 + BT[n] * T) / 8.0;
            for (int l = 0; l < 24; ++l)               /* brick.c:505-522 */
                for (int j = 0; j < 24; ++j) {
                    double sum = 0;
                    for (int q = 0; q < 3; ++q)
                        sum += ((j % 3 == q) ? h[j / 3] : 0.0) * ((l % 3 == q) ? h[l / 3] : 0.0);
                    mbr[j][l] += dens[pe + e] * sum * detJ;
                }
        }
        for (int ie = 0; ie < 24; ++ie)                /* brick.c:525-536 */
            for (int je = 0; je < 24; ++je) {
                long j = mcode[pmc + e * 24 + ie], k = mcode[pmc + e * 24 + je];
                if (j != 0 && k != 0) sm[(j - 1) * D->NEQ + k - 1] += mbr[je][ie];
            }
    }
}

/* --------------------------------------------------------------------------------- solve.c */
long orc_dense_to_csc(long neq, const double *ss, double tol, int *Ap, int *Ai, double *Ax)
{   /* solve.c:110-119 (with Ap written for every column) */
    long nz = 0;
    Ap[0] = 0;
    for (long i = 0; i < neq; ++i) {
        for (long j = 0; j < neq; ++j)
            if (fabs(ss[i * neq + j]) > tol) { Ai[nz] = (int)j; Ax[nz] = ss[i * neq + j]; ++nz; }
        Ap[i + 1] = (int)nz;
    }
    return nz;
}
