// cb_shell_plastic.cuh - material-nonlinear DKT shell (ANAFLAG 3): Ivanov's yield criterion in
// stress resultants evaluated at the three vertices, the plastic flow directions, the elasto-plastic
// membrane / bending / coupling stiffness at the controlling vertex and the strain / curvature
// increments.  Included by cb_forces.cu only (compiled with -fmad=false, reference operation order);
// the stiffness pass consumes the local matrix these routines leave in sh_kpl.
//
// reference: stiff_sh shell.c:171-283, stiffm_sh 842-1133, stiffm_m_sh 1134-1199, stiffm_b_sh
// 1201-1368, stiffm_mb_sh 1370-1503, forces_sh 1786-2325, strn_curv 2448-2553.
// The reference evaluates x^2..x^4 with libm pow() and the hardening law with libm exp(); products
// and CUDA's exp() are used here (<= 2 ulp apart, tolerance 1e-12).
#ifndef CB_SHELL_PLASTIC_CUH
#define CB_SHELL_PLASTIC_CUH

#define CB_SH_PHITOL 1e-4          // shell.c:37
#define CB_SH_PL 21                // per shell and generation: chi[3], efN[3][3], efM[3][3]

struct Ivanov { double alpha, Me, Nbar, Mbar, MNbar, q, r, s, phi; int h; };
struct IvFlow { double fn[3], fm[3], fnC[3], fmC[3], jf, kf, Bf, df_da, da_dchi; };

__device__ __forceinline__ double sq(double x) { return x * x; }

// shell.c:178-234 (the same block at 1833-1876, 1996-2041, 2145-2188); r, s, h, phi keep their
// previous values when q < 1e-4, like the reference's per-vertex arrays
__device__ __forceinline__ void ivanov_eval(Ivanov &v, const double *N, const double *M, double chi,
                                            double fy, double t, double No)
{
    v.alpha = 1.0 - 0.4 * exp(-2.6 * sqrt(chi));
    v.Me = v.alpha * 0.25 * fy * sq(t);
    v.Nbar = sq(N[0]) + sq(N[1]) - N[0] * N[1] + 3 * sq(N[2]);
    v.Mbar = sq(M[0]) + sq(M[1]) - M[0] * M[1] + 3 * sq(M[2]);
    v.MNbar = M[0] * N[0] + M[1] * N[1] - 0.5 * M[0] * N[1] - 0.5 * M[1] * N[0] + 3 * M[2] * N[2];
    v.q = v.Nbar * sq(v.Me) + 0.48 * v.Mbar * sq(No);
    if (v.q >= 1e-4) {
        v.r = sqrt(sq(No) * sq(v.Mbar) + 4 * sq(v.Me) * sq(v.MNbar));
        v.h = (v.r / (2 * sq(v.Me) * No) >= 1e-4) ? 1 : 0;
        v.s = v.Nbar * v.Mbar - sq(v.MNbar);
        if (v.h == 1)
            v.phi = v.Nbar / sq(No) + 0.5 * v.Mbar / sq(v.Me) - 0.25 * v.s / v.q + v.r / (2 * sq(v.Me) * No);
        else
            v.phi = v.Nbar / sq(No) + 0.5 * v.Mbar / sq(v.Me) - 0.25 * v.s / v.q;
    }
}

// plastic flow directions and the factors of the elasto-plastic moduli, shell.c:868-937 (the same
// block at 1880-1960 and 2052-2122).  C = plane-stress matrix (C00, C01, C22).
__device__ __forceinline__ void ivanov_flow(IvFlow &f, const Ivanov &v, const double *N, const double *M,
                                            const double (*C)[3], double E, double t, double fy,
                                            double chi, double No)
{
    const double c_fact = 1 / sq(No) - v.Mbar / (4 * v.q) + v.s * sq(v.Me) / (4 * sq(v.q));
    double g_fact, d_fact;
    if (v.h == 1) {
        g_fact = v.MNbar * (1 / (4 * v.q) + 1 / (No * v.r));
        d_fact = 1 / (2 * sq(v.Me)) - v.Nbar / (4 * v.q) + 0.12 * sq(No) * v.s / sq(v.q) +
                 v.Mbar * No / (2 * sq(v.Me) * v.r);
    } else {
        g_fact = v.MNbar / (4 * v.q);
        d_fact = 1 / (2 * sq(v.Me)) - v.Nbar / (4 * v.q) + 0.12 * sq(No) * v.s / sq(v.q);
    }
    const double gN[3] = {2 * N[0] - N[1], 2 * N[1] - N[0], 6 * N[2]};
    const double gM[3] = {2 * M[0] - M[1], 2 * M[1] - M[0], 6 * M[2]};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        f.fn[j] = c_fact * gN[j] + g_fact * gM[j];
        f.fm[j] = g_fact * gN[j] + d_fact * gM[j];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double s1 = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) { s1 += f.fn[k] * C[k][j]; s2 += f.fm[k] * C[k][j]; }
        f.fnC[j] = s1; f.fmC[j] = s2;
    }
    double s1 = 0, s2 = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) { s1 += f.fnC[j] * f.fn[j]; s2 += f.fmC[j] * f.fm[j]; }
    f.jf = t * s1;
    f.kf = (t * t * t) * s2 / 12;
    f.Bf = 2 * sqrt(sq(g_fact) * v.Nbar + sq(d_fact) * v.Mbar + 2 * d_fact * g_fact * v.MNbar);
    if (v.h == 1)
        f.df_da = -(v.Mbar / (v.alpha * sq(v.Me))) + v.s * v.Nbar * sq(v.Me) / (2 * sq(v.q) * v.alpha) -
                  v.r / (v.alpha * sq(v.Me) * No) + 2 * sq(v.MNbar) / (v.alpha * No * v.r);
    else
        f.df_da = -(v.Mbar / (v.alpha * sq(v.Me))) + v.s * v.Nbar * sq(v.Me) / (2 * sq(v.q) * v.alpha);
    if (chi >= 1e-6) f.da_dchi = 0.52 * E * t * exp(-2.6 * sqrt(chi)) / (3 * fy * sqrt(chi));
    else f.da_dchi = 0;
}

// sc = shell constants with sc[5..7] = x2, x3, y3 and sc[8..10] = the side lengths to use
// (dkt_alpha_T / membrane_B of cb_forces.cu read them there)
__device__ void dkt_alpha_T(const double *sc, double aT[9][9]);

// stiffm_sh and its three helpers: the 18x18 local matrix at the controlling yielded vertex.
// Entries are assigned (the geometric part is added by the stiffness pass).  k must be zeroed.
__device__ void shell_plastic_k(double (*k)[18], const IvFlow &f, const double (*C)[3], const double *sc,
                                double t, double Adef)
{
    const int FM[6] = {0, 1, 6, 7, 12, 13}, FB[9] = {2, 3, 4, 8, 9, 10, 14, 15, 16};
    const double den = f.jf + f.kf - f.Bf * f.df_da * f.da_dchi;
    const double t3 = t * t * t;
    double Bm[3][6], aT[9][9];
    membrane_B(sc, Adef, Bm);
    dkt_alpha_T(sc, aT);
    {   // membrane, shell.c:1134-1199
        const double xi = t / den;
        double xNC[3][3], Cs[3][3], BC[6][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += (f.fn[i] * f.fn[q]) * C[q][j];
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
    {   // bending, shell.c:1201-1368
        const double xi = t3 / (12 * den);
        double xMC[3][3], Ds[3][3], Q[9][9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += (f.fm[i] * f.fm[q]) * C[q][j];
                xMC[i][j] = xi * s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += C[i][q] * xMC[q][j];
                Ds[i][j] = t3 * (C[i][j] - s) / 12;
            }
        for (int i = 0; i < 9; ++i) {        // shell.c:1314-1357: the three row blocks are alike
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
                k[FB[i]][FB[j]] = s / (2 * Adef);
            }
        k[5][5] = k[3][3] / 10000; k[11][11] = k[9][9] / 10000; k[17][17] = k[15][15] / 10000;
    }
    {   // membrane-bending coupling, shell.c:1370-1503
        const double xi = -(t3 * t) / (12 * den);
        double xNMC[3][3], cd[3][3], BcL[6][9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += (f.fn[i] * f.fm[q]) * C[q][j];
                xNMC[i][j] = xi * s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += C[i][q] * xNMC[q][j];
                cd[i][j] = s;
            }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 9; ++j) {
                double s = 0;
                for (int q = 0; q < 3; ++q) s += Bm[q][i] * (cd[q][j / 3] / 6);
                BcL[i][j] = s;
            }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 9; ++j) {
                double s = 0;
                for (int q = 0; q < 9; ++q) s += BcL[i][q] * aT[j][q];
                k[FM[i]][FB[j]] = s; k[FB[j]][FM[i]] = s;
            }
    }
}

// strn_curv, shell.c:2448-2553: membrane strain increment and the curvature increments at the
// three vertices (the rows of its LL_alpha are columns of alpha^T: LL_alpha[v][j][k] = aT[k][3j+v])
__device__ void strain_curvature(double *strn, double (*curv)[3], const double *ddm, const double *ddb,
                                 const double *sc, double Adef)
{
    double Bm[3][6], aT[9][9];
    membrane_B(sc, Adef, Bm);
    for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int j = 0; j < 6; ++j) s += Bm[i][j] * ddm[j];
        strn[i] = s;
    }
    dkt_alpha_T(sc, aT);
    for (int v = 0; v < 3; ++v)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int q = 0; q < 9; ++q) s += aT[q][3 * j + v] * ddb[q];
            curv[v][j] = s / (2 * Adef);
        }
}

// the yielded vertex that controls the element: smallest phi among those on the surface
__device__ __forceinline__ void pick_vertex(int &yv, const double *phi, int i)
{
    if (yv == 0) yv = i + 1;
    else if (phi[i] < phi[yv - 1]) yv = i + 1;
}

#endif
