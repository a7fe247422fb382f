// cb_frame_math.cuh - local (element-axes) tangent stiffness of the 14-DOF space frame, shared
// by the force path (compiled with -fmad=false: reference rounding) and the stiffness path.
//
// DOF order per end: u, v, w, theta_x (torsion), theta_y, theta_z, warping; end 2 at +7.
// stiffe_fr frame.c:364-408 (elastic, warping torsion included), stiffg_fr frame.c:410-579
// (geometric, closed form in the total end forces of the previous configuration), release
// frame.c:798-900 (member-end bending releases by static condensation).
#ifndef CB_FRAME_MATH_CUH
#define CB_FRAME_MATH_CUH

#include "cb_internal.h"

// fc = fr_const record: E, G, A, L0, L0^2, L0^3 (host libm), Iz, Iy, J, Cw, aux xyz
__device__ __forceinline__ void frame_elastic(double (*k)[14], const double *fc)
{
    const double E = fc[0], G = fc[1], A = fc[2], L = fc[3], L3 = fc[5];
    const double Iz = fc[6], Iy = fc[7], J = fc[8], Cw = fc[9];
#define CB_SYM(i, j, v) k[i][j] = k[j][i] = (v)
    k[0][0] = k[7][7] = E * A / L;                 CB_SYM(7, 0, -k[0][0]);
    k[1][1] = k[8][8] = 12 * E * Iz / L3;          CB_SYM(8, 1, -k[1][1]);
    k[2][2] = k[9][9] = 12 * E * Iy / L3;          CB_SYM(9, 2, -k[2][2]);
    k[3][3] = k[10][10] = 6 * G * J / (5 * L) + 12 * E * Cw / L3;   CB_SYM(10, 3, -k[3][3]);
    k[5][5] = k[12][12] = 4 * E * Iz / L;
    k[4][4] = k[11][11] = 4 * E * Iy / L;
    k[6][6] = k[13][13] = 2 * G * J * L / 15 + 4 * E * Cw / L;
    CB_SYM(5, 1, 6 * E * Iz / (L * L)); CB_SYM(12, 1, k[5][1]);
    CB_SYM(8, 5, -k[5][1]);             CB_SYM(12, 8, -k[5][1]);
    CB_SYM(9, 4, 6 * E * Iy / (L * L)); CB_SYM(11, 9, k[9][4]);
    CB_SYM(4, 2, -k[9][4]);             CB_SYM(11, 2, -k[9][4]);
    CB_SYM(6, 3, G * J / 10 + 6 * E * Cw / (L * L)); CB_SYM(13, 3, k[6][3]);
    CB_SYM(10, 6, -k[6][3]);            CB_SYM(13, 10, -k[6][3]);
    CB_SYM(12, 5, 2 * E * Iz / L);
    CB_SYM(11, 4, 2 * E * Iy / L);
    CB_SYM(13, 6, -(G * J * L / 30 - 2 * E * Cw / L));
#undef CB_SYM
}

// ef = total end forces (ef_ip + efFE_ip) in the previous local frame, L = deformed length
__device__ __forceinline__ void frame_geometric(double (*k)[14], const double *ef, double L,
                                                double A, double J)
{
    const double P = ef[7], M4 = ef[4], M5 = ef[5], M10 = ef[10], M11 = ef[11], M12 = ef[12];
#define CB_ADD(i, j, v) do { k[i][j] += (v); k[j][i] += (v); } while (0)
#define CB_SUB(i, j, v) do { k[i][j] -= (v); k[j][i] -= (v); } while (0)
    k[0][0] += P / L; k[7][7] += P / L; CB_SUB(7, 0, P / L);
    k[1][1] += 6 * P / (5 * L); k[8][8] += 6 * P / (5 * L);
    k[2][2] += 6 * P / (5 * L); k[9][9] += 6 * P / (5 * L);
    CB_SUB(8, 1, 6 * P / (5 * L)); CB_SUB(9, 2, 6 * P / (5 * L));
    k[3][3] += 6 * P * J / (5 * A * L); k[10][10] += 6 * P * J / (5 * A * L);
    CB_SUB(10, 3, 6 * P * J / (5 * A * L));
    k[4][4] += 2 * P * L / 15; k[11][11] += 2 * P * L / 15;
    k[5][5] += 2 * P * L / 15; k[12][12] += 2 * P * L / 15;
    k[6][6] += 2 * P * J / (15 * A); k[13][13] += 2 * P * J / (15 * A);
    CB_ADD(3, 1, (11 * M4 - M11) / (10 * L)); CB_SUB(8, 3, (11 * M4 - M11) / (10 * L));
    CB_ADD(4, 1, M10 / L); CB_ADD(5, 2, M10 / L); CB_ADD(11, 8, M10 / L); CB_ADD(12, 9, M10 / L);
    CB_SUB(11, 1, M10 / L); CB_SUB(12, 2, M10 / L); CB_SUB(8, 4, M10 / L); CB_SUB(9, 5, M10 / L);
    CB_ADD(5, 1, P / 10); CB_ADD(12, 1, P / 10); CB_ADD(9, 4, P / 10); CB_ADD(11, 9, P / 10);
    CB_SUB(4, 2, P / 10); CB_SUB(11, 2, P / 10); CB_SUB(8, 5, P / 10); CB_SUB(12, 8, P / 10);
    CB_ADD(6, 1, M4 / 10); CB_SUB(8, 6, M4 / 10);
    CB_ADD(10, 8, (M4 - 11 * M11) / (10 * L)); CB_SUB(10, 1, (M4 - 11 * M11) / (10 * L));
    CB_ADD(13, 8, M11 / 10); CB_SUB(13, 1, M11 / 10);
    CB_ADD(3, 2, (11 * M5 - M12) / (10 * L)); CB_SUB(9, 3, (11 * M5 - M12) / (10 * L));
    CB_ADD(6, 2, M5 / 10); CB_SUB(9, 6, M5 / 10);
    CB_ADD(10, 9, (M5 - 11 * M12) / (10 * L)); CB_SUB(10, 2, (M5 - 11 * M12) / (10 * L));
    CB_ADD(13, 9, M12 / 10); CB_SUB(13, 2, M12 / 10);
    CB_SUB(4, 3, (2 * M5 - M12) / 5); CB_ADD(5, 3, (2 * M4 - M11) / 5);
    CB_ADD(6, 3, P * J / (10 * A)); CB_ADD(13, 3, P * J / (10 * A));
    CB_SUB(10, 6, P * J / (10 * A)); CB_SUB(13, 10, P * J / (10 * A));
    CB_SUB(11, 3, (2 * M5 + M12) / 10); CB_ADD(12, 3, (2 * M4 + M11) / 10);
    CB_SUB(6, 4, (3 * M5 - M12) * L / 30); CB_SUB(10, 4, (M5 + 2 * M12) / 10);
    CB_SUB(11, 4, P * L / 30); CB_SUB(12, 5, P * L / 30);
    CB_ADD(12, 4, M10 / 2); CB_SUB(11, 5, M10 / 2);
    CB_ADD(13, 4, M5 * L / 30);
    CB_ADD(6, 5, (3 * M4 - M11) * L / 30); CB_ADD(10, 5, (M4 + 2 * M11) / 10);
    CB_SUB(13, 5, M4 * L / 30);
    CB_SUB(11, 6, M12 * L / 30); CB_ADD(12, 6, M11 * L / 30);
    CB_SUB(13, 6, P * J / (30 * A));
    CB_ADD(11, 10, (M5 - 2 * M12) / 5); CB_SUB(12, 10, (M4 - 2 * M11) / 5);
    CB_SUB(13, 11, (M5 - 3 * M12) * L / 30); CB_ADD(13, 12, (M4 - 3 * M11) * L / 30);
#undef CB_ADD
#undef CB_SUB
}

// the per-member records are AoS with 16-byte multiples as strides (80, 112, 128, 48 bytes): 16-byte
// vector accesses halve the sectors each request touches
template <int N>
__device__ __forceinline__ void ldv2(const double *p, double *r)
{
    static_assert(N % 2 == 0, "even record length");
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const double2 v = reinterpret_cast<const double2 *>(p)[i];
        r[2 * i] = v.x; r[2 * i + 1] = v.y;
    }
}
template <int N>
__device__ __forceinline__ void stv2(double *p, const double *r)
{
    static_assert(N % 2 == 0, "even record length");
#pragma unroll
    for (int i = 0; i < N / 2; ++i) reinterpret_cast<double2 *>(p)[i] = make_double2(r[2 * i], r[2 * i + 1]);
}

// ---- packed variants for the force path: elastic + geometric tangent in a thread-private column of
// SHARED memory, upper triangle only (both are bitwise symmetric: every entry and its mirror get
// the same values in the same order, frame.c:364-579), entry (i, j) of thread t at
// sk[tri(i, j) * CB_FR_TPB + t].  Same expressions and operation order as frame_elastic /
// frame_geometric above; the point is that the 14x14 matrix never lives in local memory.
#define CB_FR_TPB 64
#define CB_FR_TRI(i, j) ((i) >= (j) ? (i) * ((i) + 1) / 2 + (j) : (j) * ((j) + 1) / 2 + (i))
#define KP(i, j) sk[CB_FR_TRI(i, j) * CB_FR_TPB]
__device__ __forceinline__ void frame_elastic_packed(double *sk, const double *fc)
{
    const double E = fc[0], G = fc[1], A = fc[2], L = fc[3], L3 = fc[5];
    const double Iz = fc[6], Iy = fc[7], J = fc[8], Cw = fc[9];
#define CB_SYM(i, j, v) KP(i, j) = (v)
    KP(0, 0) = KP(7, 7) = E * A / L;                 CB_SYM(7, 0, -KP(0, 0));
    KP(1, 1) = KP(8, 8) = 12 * E * Iz / L3;          CB_SYM(8, 1, -KP(1, 1));
    KP(2, 2) = KP(9, 9) = 12 * E * Iy / L3;          CB_SYM(9, 2, -KP(2, 2));
    KP(3, 3) = KP(10, 10) = 6 * G * J / (5 * L) + 12 * E * Cw / L3;   CB_SYM(10, 3, -KP(3, 3));
    KP(5, 5) = KP(12, 12) = 4 * E * Iz / L;
    KP(4, 4) = KP(11, 11) = 4 * E * Iy / L;
    KP(6, 6) = KP(13, 13) = 2 * G * J * L / 15 + 4 * E * Cw / L;
    CB_SYM(5, 1, 6 * E * Iz / (L * L)); CB_SYM(12, 1, KP(5, 1));
    CB_SYM(8, 5, -KP(5, 1));             CB_SYM(12, 8, -KP(5, 1));
    CB_SYM(9, 4, 6 * E * Iy / (L * L)); CB_SYM(11, 9, KP(9, 4));
    CB_SYM(4, 2, -KP(9, 4));             CB_SYM(11, 2, -KP(9, 4));
    CB_SYM(6, 3, G * J / 10 + 6 * E * Cw / (L * L)); CB_SYM(13, 3, KP(6, 3));
    CB_SYM(10, 6, -KP(6, 3));            CB_SYM(13, 10, -KP(6, 3));
    CB_SYM(12, 5, 2 * E * Iz / L);
    CB_SYM(11, 4, 2 * E * Iy / L);
    CB_SYM(13, 6, -(G * J * L / 30 - 2 * E * Cw / L));
#undef CB_SYM
}

__device__ __forceinline__ void frame_geometric_packed(double *sk, const double *ef, double L, double A,
                                                       double J)
{
    const double P = ef[7], M4 = ef[4], M5 = ef[5], M10 = ef[10], M11 = ef[11], M12 = ef[12];
#define CB_ADD(i, j, v) do { KP(i, j) += (v); } while (0)
#define CB_SUB(i, j, v) do { KP(i, j) -= (v); } while (0)
    KP(0, 0) += P / L; KP(7, 7) += P / L; CB_SUB(7, 0, P / L);
    KP(1, 1) += 6 * P / (5 * L); KP(8, 8) += 6 * P / (5 * L);
    KP(2, 2) += 6 * P / (5 * L); KP(9, 9) += 6 * P / (5 * L);
    CB_SUB(8, 1, 6 * P / (5 * L)); CB_SUB(9, 2, 6 * P / (5 * L));
    KP(3, 3) += 6 * P * J / (5 * A * L); KP(10, 10) += 6 * P * J / (5 * A * L);
    CB_SUB(10, 3, 6 * P * J / (5 * A * L));
    KP(4, 4) += 2 * P * L / 15; KP(11, 11) += 2 * P * L / 15;
    KP(5, 5) += 2 * P * L / 15; KP(12, 12) += 2 * P * L / 15;
    KP(6, 6) += 2 * P * J / (15 * A); KP(13, 13) += 2 * P * J / (15 * A);
    CB_ADD(3, 1, (11 * M4 - M11) / (10 * L)); CB_SUB(8, 3, (11 * M4 - M11) / (10 * L));
    CB_ADD(4, 1, M10 / L); CB_ADD(5, 2, M10 / L); CB_ADD(11, 8, M10 / L); CB_ADD(12, 9, M10 / L);
    CB_SUB(11, 1, M10 / L); CB_SUB(12, 2, M10 / L); CB_SUB(8, 4, M10 / L); CB_SUB(9, 5, M10 / L);
    CB_ADD(5, 1, P / 10); CB_ADD(12, 1, P / 10); CB_ADD(9, 4, P / 10); CB_ADD(11, 9, P / 10);
    CB_SUB(4, 2, P / 10); CB_SUB(11, 2, P / 10); CB_SUB(8, 5, P / 10); CB_SUB(12, 8, P / 10);
    CB_ADD(6, 1, M4 / 10); CB_SUB(8, 6, M4 / 10);
    CB_ADD(10, 8, (M4 - 11 * M11) / (10 * L)); CB_SUB(10, 1, (M4 - 11 * M11) / (10 * L));
    CB_ADD(13, 8, M11 / 10); CB_SUB(13, 1, M11 / 10);
    CB_ADD(3, 2, (11 * M5 - M12) / (10 * L)); CB_SUB(9, 3, (11 * M5 - M12) / (10 * L));
    CB_ADD(6, 2, M5 / 10); CB_SUB(9, 6, M5 / 10);
    CB_ADD(10, 9, (M5 - 11 * M12) / (10 * L)); CB_SUB(10, 2, (M5 - 11 * M12) / (10 * L));
    CB_ADD(13, 9, M12 / 10); CB_SUB(13, 2, M12 / 10);
    CB_SUB(4, 3, (2 * M5 - M12) / 5); CB_ADD(5, 3, (2 * M4 - M11) / 5);
    CB_ADD(6, 3, P * J / (10 * A)); CB_ADD(13, 3, P * J / (10 * A));
    CB_SUB(10, 6, P * J / (10 * A)); CB_SUB(13, 10, P * J / (10 * A));
    CB_SUB(11, 3, (2 * M5 + M12) / 10); CB_ADD(12, 3, (2 * M4 + M11) / 10);
    CB_SUB(6, 4, (3 * M5 - M12) * L / 30); CB_SUB(10, 4, (M5 + 2 * M12) / 10);
    CB_SUB(11, 4, P * L / 30); CB_SUB(12, 5, P * L / 30);
    CB_ADD(12, 4, M10 / 2); CB_SUB(11, 5, M10 / 2);
    CB_ADD(13, 4, M5 * L / 30);
    CB_ADD(6, 5, (3 * M4 - M11) * L / 30); CB_ADD(10, 5, (M4 + 2 * M11) / 10);
    CB_SUB(13, 5, M4 * L / 30);
    CB_SUB(11, 6, M12 * L / 30); CB_ADD(12, 6, M11 * L / 30);
    CB_SUB(13, 6, P * J / (30 * A));
    CB_ADD(11, 10, (M5 - 2 * M12) / 5); CB_SUB(12, 10, (M4 - 2 * M11) / 5);
    CB_SUB(13, 11, (M5 - 3 * M12) * L / 30); CB_ADD(13, 12, (M4 - 3 * M11) * L / 30);
#undef CB_ADD
#undef CB_SUB
}

// ---- stiffness-path variants: the same closed forms with every division replaced by a product with
// a reciprocal formed once (three FP64 divisions per block instead of ~60, each of which is a
// ~30-instruction sequence on the FP64 pipe).  K_t has no cancellation beyond its own magnitude, so
// the 1-2 ulp per entry this moves stays far inside the 1e-12 norm-wise tolerance; the force path
// keeps the reference's rounding sequence above.
__device__ __forceinline__ void frame_elastic_rcp(double (*k)[14], const double *fc)
{
    const double E = fc[0], G = fc[1], A = fc[2], L = fc[3];
    const double Iz = fc[6], Iy = fc[7], J = fc[8], Cw = fc[9];
    const double iL = 1.0 / L, iL2 = iL * iL, iL3 = iL2 * iL;
    const double EIz = E * Iz, EIy = E * Iy, GJ = G * J, ECw = E * Cw;
#define CB_SYM(i, j, v) k[i][j] = k[j][i] = (v)
    k[0][0] = k[7][7] = E * A * iL;                CB_SYM(7, 0, -k[0][0]);
    k[1][1] = k[8][8] = 12 * EIz * iL3;            CB_SYM(8, 1, -k[1][1]);
    k[2][2] = k[9][9] = 12 * EIy * iL3;            CB_SYM(9, 2, -k[2][2]);
    k[3][3] = k[10][10] = 1.2 * GJ * iL + 12 * ECw * iL3;   CB_SYM(10, 3, -k[3][3]);
    k[5][5] = k[12][12] = 4 * EIz * iL;
    k[4][4] = k[11][11] = 4 * EIy * iL;
    k[6][6] = k[13][13] = (2.0 / 15.0) * GJ * L + 4 * ECw * iL;
    CB_SYM(5, 1, 6 * EIz * iL2); CB_SYM(12, 1, k[5][1]);
    CB_SYM(8, 5, -k[5][1]);      CB_SYM(12, 8, -k[5][1]);
    CB_SYM(9, 4, 6 * EIy * iL2); CB_SYM(11, 9, k[9][4]);
    CB_SYM(4, 2, -k[9][4]);      CB_SYM(11, 2, -k[9][4]);
    CB_SYM(6, 3, 0.1 * GJ + 6 * ECw * iL2); CB_SYM(13, 3, k[6][3]);
    CB_SYM(10, 6, -k[6][3]);     CB_SYM(13, 10, -k[6][3]);
    CB_SYM(12, 5, 2 * EIz * iL);
    CB_SYM(11, 4, 2 * EIy * iL);
    CB_SYM(13, 6, -(GJ * L * (1.0 / 30.0) - 2 * ECw * iL));
#undef CB_SYM
}

__device__ __forceinline__ void frame_geometric_rcp(double (*k)[14], const double *ef, double L,
                                                    double A, double J)
{
    const double P = ef[7], M4 = ef[4], M5 = ef[5], M10 = ef[10], M11 = ef[11], M12 = ef[12];
    const double iL = 1.0 / L, JA = J / A;
    const double PL = P * iL, PL65 = 1.2 * PL, PJ = P * JA, PL30 = P * L * (1.0 / 30.0);
    const double L30 = L * (1.0 / 30.0), iL10 = 0.1 * iL;
#define CB_ADD(i, j, v) do { const double v_ = (v); k[i][j] += v_; k[j][i] += v_; } while (0)
#define CB_SUB(i, j, v) do { const double v_ = (v); k[i][j] -= v_; k[j][i] -= v_; } while (0)
    k[0][0] += PL; k[7][7] += PL; CB_SUB(7, 0, PL);
    k[1][1] += PL65; k[8][8] += PL65;
    k[2][2] += PL65; k[9][9] += PL65;
    CB_SUB(8, 1, PL65); CB_SUB(9, 2, PL65);
    k[3][3] += PL65 * JA; k[10][10] += PL65 * JA;
    CB_SUB(10, 3, PL65 * JA);
    k[4][4] += 4 * PL30; k[11][11] += 4 * PL30;
    k[5][5] += 4 * PL30; k[12][12] += 4 * PL30;
    k[6][6] += (2.0 / 15.0) * PJ; k[13][13] += (2.0 / 15.0) * PJ;
    CB_ADD(3, 1, (11 * M4 - M11) * iL10); CB_SUB(8, 3, (11 * M4 - M11) * iL10);
    CB_ADD(4, 1, M10 * iL); CB_ADD(5, 2, M10 * iL); CB_ADD(11, 8, M10 * iL); CB_ADD(12, 9, M10 * iL);
    CB_SUB(11, 1, M10 * iL); CB_SUB(12, 2, M10 * iL); CB_SUB(8, 4, M10 * iL); CB_SUB(9, 5, M10 * iL);
    CB_ADD(5, 1, 0.1 * P); CB_ADD(12, 1, 0.1 * P); CB_ADD(9, 4, 0.1 * P); CB_ADD(11, 9, 0.1 * P);
    CB_SUB(4, 2, 0.1 * P); CB_SUB(11, 2, 0.1 * P); CB_SUB(8, 5, 0.1 * P); CB_SUB(12, 8, 0.1 * P);
    CB_ADD(6, 1, 0.1 * M4); CB_SUB(8, 6, 0.1 * M4);
    CB_ADD(10, 8, (M4 - 11 * M11) * iL10); CB_SUB(10, 1, (M4 - 11 * M11) * iL10);
    CB_ADD(13, 8, 0.1 * M11); CB_SUB(13, 1, 0.1 * M11);
    CB_ADD(3, 2, (11 * M5 - M12) * iL10); CB_SUB(9, 3, (11 * M5 - M12) * iL10);
    CB_ADD(6, 2, 0.1 * M5); CB_SUB(9, 6, 0.1 * M5);
    CB_ADD(10, 9, (M5 - 11 * M12) * iL10); CB_SUB(10, 2, (M5 - 11 * M12) * iL10);
    CB_ADD(13, 9, 0.1 * M12); CB_SUB(13, 2, 0.1 * M12);
    CB_SUB(4, 3, (2 * M5 - M12) * 0.2); CB_ADD(5, 3, (2 * M4 - M11) * 0.2);
    CB_ADD(6, 3, 0.1 * PJ); CB_ADD(13, 3, 0.1 * PJ);
    CB_SUB(10, 6, 0.1 * PJ); CB_SUB(13, 10, 0.1 * PJ);
    CB_SUB(11, 3, (2 * M5 + M12) * 0.1); CB_ADD(12, 3, (2 * M4 + M11) * 0.1);
    CB_SUB(6, 4, (3 * M5 - M12) * L30); CB_SUB(10, 4, (M5 + 2 * M12) * 0.1);
    CB_SUB(11, 4, PL30); CB_SUB(12, 5, PL30);
    CB_ADD(12, 4, 0.5 * M10); CB_SUB(11, 5, 0.5 * M10);
    CB_ADD(13, 4, M5 * L30);
    CB_ADD(6, 5, (3 * M4 - M11) * L30); CB_ADD(10, 5, (M4 + 2 * M11) * 0.1);
    CB_SUB(13, 5, M4 * L30);
    CB_SUB(11, 6, M12 * L30); CB_ADD(12, 6, M11 * L30);
    CB_SUB(13, 6, PJ * (1.0 / 30.0));
    CB_ADD(11, 10, (M5 - 2 * M12) * 0.2); CB_SUB(12, 10, (M4 - 2 * M11) * 0.2);
    CB_SUB(13, 11, (M5 - 3 * M12) * L30); CB_ADD(13, 12, (M4 - 3 * M11) * L30);
#undef CB_ADD
#undef CB_SUB
}

// ---- concentrated plasticity (ANAFLAG 3): yield surface of frame.c:617-621, its gradients
// (frame.c:625-648) and the plastic reduction of the tangent, stiffm_fr frame.c:581-796.
// The reference evaluates the powers with libm pow(); products are used here (<= 2 ulp apart).
#define CB_PHITOL 1e-4                      // frame.c:37
__device__ __forceinline__ double fr_phi(double p, double my, double mz)
{
    const double p2 = p * p, mz2 = mz * mz, my2 = my * my;
    return p2 + mz2 + my2 * my2 + 3.5 * p2 * mz2 + 3 * (p2 * p2 * p2) * my2 + 4.5 * (mz2 * mz2) * my2;
}
// gradient w.r.t. (axial force, weak-axis moment, strong-axis moment) of one member end
__device__ __forceinline__ void fr_grad(double p, double my, double mz, double Py, double Mpy,
                                        double Mpz, double *g)
{
    const double p2 = p * p, mz2 = mz * mz, my2 = my * my;
    g[0] = 2 * p / Py + 7 * p * mz2 / Py + 18 * (p2 * p2 * p) * my2 / Py;
    g[1] = 4 * (my2 * my) / Mpy + 6 * (p2 * p2 * p2) * my / Mpy + 9 * (mz2 * mz2) * my / Mpy;
    g[2] = 2 * mz / Mpz + 7 * p2 * mz / Mpz + 18 * (mz2 * mz) * my2 / Mpz;
}

// k <- k - kG (G^T k G)^-1 G^T k with G the yield-surface gradients of the ends on the surface.
// pl = Py, Mpy, Mpz; y0, y1 = yldflag of the two ends.  Every index is static.
__device__ __forceinline__ void frame_plastic(double (*k)[14], const double *eft, int y0, int y1,
                                              const double *pl)
{
    const double Py = pl[0], Mpy = pl[1], Mpz = pl[2];
    const double p0 = eft[0] / Py, p1 = eft[7] / Py, my0 = eft[4] / Mpy, my1 = eft[11] / Mpy;
    const double mz0 = eft[5] / Mpz, mz1 = eft[12] / Mpz;
    const double phi0 = fr_phi(p0, my0, mz0), phi1 = fr_phi(p1, my1, mz1);
    const bool on0 = phi0 >= 1 - CB_PHITOL, on1 = phi1 >= 1 - CB_PHITOL;
    if (on0 && on1) {
        double g0[3] = {0, 0, 0}, g1[3] = {0, 0, 0};
        if (y0 != 2) fr_grad(p0, my0, mz0, Py, Mpy, Mpz, g0);
        if (y1 != 2) fr_grad(p1, my1, mz1, Py, Mpy, Mpz, g1);
        double kG[14][2];
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            kG[i][0] = k[i][0] * g0[0] + k[i][4] * g0[1] + k[i][5] * g0[2];
            kG[i][1] = k[i][7] * g1[0] + k[i][11] * g1[1] + k[i][12] * g1[2];
        }
        double a00 = kG[0][0] * g0[0] + kG[4][0] * g0[1] + kG[5][0] * g0[2];
        double a01 = kG[7][0] * g1[0] + kG[11][0] * g1[1] + kG[12][0] * g1[2];
        double a10 = kG[0][1] * g0[0] + kG[4][1] * g0[1] + kG[5][1] * g0[2];
        double a11 = kG[7][1] * g1[0] + kG[11][1] * g1[1] + kG[12][1] * g1[2];
        const double det = a00 * a11 - a01 * a10, tmp = a00;
        a00 = a11 / det; a11 = tmp / det; a01 *= -1 / det; a10 *= -1 / det;
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            const double w0 = kG[i][0] * a00 + kG[i][1] * a10, w1 = kG[i][0] * a01 + kG[i][1] * a11;
#pragma unroll
            for (int j = 0; j < 14; ++j) k[i][j] -= w0 * kG[j][0] + w1 * kG[j][1];
        }
    } else if (on0 && y0 != 2) {
        double g[3], kG[14];
        fr_grad(p0, my0, mz0, Py, Mpy, Mpz, g);
#pragma unroll
        for (int i = 0; i < 14; ++i) kG[i] = k[i][0] * g[0] + k[i][4] * g[1] + k[i][5] * g[2];
        const double inv = 1 / (kG[0] * g[0] + kG[4] * g[1] + kG[5] * g[2]);
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            const double w = kG[i] * inv;
#pragma unroll
            for (int j = 0; j < 14; ++j) k[i][j] -= w * kG[j];
        }
    } else if (on1 && y1 != 2) {
        double g[3], kG[14];
        fr_grad(p1, my1, mz1, Py, Mpy, Mpz, g);
#pragma unroll
        for (int i = 0; i < 14; ++i) kG[i] = k[i][7] * g[0] + k[i][11] * g[1] + k[i][12] * g[2];
        const double inv = 1 / (kG[7] * g[0] + kG[11] * g[1] + kG[12] * g[2]);
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            const double w = kG[i] * inv;
#pragma unroll
            for (int j = 0; j < 14; ++j) k[i][j] -= w * kG[j];
        }
    }
}

// scale of the load increment that brings a member end back onto the yield surface
// (regula_falsi, frame.c:1397-1455; its loop condition is never true, App. B.8: one refinement)
__device__ __forceinline__ double fr_regula_falsi(double p, double dp, double my, double dmy,
                                                  double mz, double dmz)
{
    double tau_u = 1, tau_l = 0;
    double phi_u = fr_phi(p + tau_u * dp, my + tau_u * dmy, mz + tau_u * dmz);
    double phi_l = fr_phi(p + tau_l * dp, my + tau_l * dmy, mz + tau_l * dmz);
    double tau_r = tau_u - (phi_u - 1) * (tau_l - tau_u) / (phi_l - phi_u);
    double phi_r = fr_phi(p + tau_r * dp, my + tau_r * dmy, mz + tau_r * dmz);
    if ((phi_l - 1 > 0 && phi_r - 1 > 0) || (phi_l - 1 < 0 && phi_r - 1 < 0)) { tau_l = tau_r; phi_l = phi_r; }
    else { tau_u = tau_r; phi_u = phi_r; }
    tau_r = tau_u - (phi_u - 1) * (tau_l - tau_u) / (phi_l - phi_u);
    return tau_r;
}

// elastic unloading test of yielded member ends (unload, frame.c:1457-1658): sign of the plastic
// multipliers lambda = (G^T k G)^-1 G^T k dd.  k = elastic + geometric tangent, dl = local
// displacement increment.  Returns 0, or 1 / 2 / 3 = both ends / end 1 / end 2 unload.
__device__ __forceinline__ int fr_unload(const double *phi, const double *p, const double *my,
                                         const double *mz, const double *pl, double (*k)[14],
                                         const double *dl)
{
    const double Py = pl[0], Mpy = pl[1], Mpz = pl[2];
    const bool on0 = phi[0] >= 1 - CB_PHITOL, on1 = phi[1] >= 1 - CB_PHITOL;
    if (on0 && on1) {
        double g0[3], g1[3], Gk[2][14];
        fr_grad(p[0], my[0], mz[0], Py, Mpy, Mpz, g0);
        fr_grad(p[1], my[1], mz[1], Py, Mpy, Mpz, g1);
        for (int j = 0; j < 14; ++j) {
            Gk[0][j] = g0[0] * k[0][j] + g0[1] * k[4][j] + g0[2] * k[5][j];
            Gk[1][j] = g1[0] * k[7][j] + g1[1] * k[11][j] + g1[2] * k[12][j];
        }
        double a00 = Gk[0][0] * g0[0] + Gk[0][4] * g0[1] + Gk[0][5] * g0[2];
        double a01 = Gk[0][7] * g1[0] + Gk[0][11] * g1[1] + Gk[0][12] * g1[2];
        double a10 = Gk[1][0] * g0[0] + Gk[1][4] * g0[1] + Gk[1][5] * g0[2];
        double a11 = Gk[1][7] * g1[0] + Gk[1][11] * g1[1] + Gk[1][12] * g1[2];
        const double det = a00 * a11 - a01 * a10, tmp = a00;
        a00 = a11 / det; a11 = tmp / det; a01 *= -1 / det; a10 *= -1 / det;
        double l0 = 0, l1 = 0;
        for (int j = 0; j < 14; ++j) {
            l0 += (a00 * Gk[0][j] + a01 * Gk[1][j]) * dl[j];
            l1 += (a10 * Gk[0][j] + a11 * Gk[1][j]) * dl[j];
        }
        if (l0 < -1e-8 && l1 < -1e-8) return 1;
        if (l0 < -1e-8) return 2;
        if (l1 < -1e-8) return 3;
    } else if (on0 || on1) {
        const int o = on0 ? 0 : 7, e = on0 ? 0 : 1;
        double g[3], Gk[14];
        fr_grad(p[e], my[e], mz[e], Py, Mpy, Mpz, g);
        for (int j = 0; j < 14; ++j) Gk[j] = g[0] * k[o][j] + g[1] * k[o + 4][j] + g[2] * k[o + 5][j];
        const double inv = 1 / (Gk[o] * g[0] + Gk[o + 4] * g[1] + Gk[o + 5] * g[2]);
        double lam = 0;
        for (int j = 0; j < 14; ++j) lam += (inv * Gk[j]) * dl[j];
        if (lam < -1e-8) return on0 ? 2 : 3;
    }
    return 0;
}

// Gauss-Jordan inverse without pivoting, as misc.c:284-343 behaves for SPD input (n <= 4)
__device__ __forceinline__ void gj_inverse4(double *A, int n)
{
    double aug[4][8];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 2 * n; ++j)
            aug[i][j] = (j < n) ? A[n * i + j] : ((j - n == i) ? 1.0 : 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            if (j != i) {
                const double m = aug[j][i] / aug[i][i];
                for (int q = 0; q < 2 * n; ++q) aug[j][q] -= m * aug[i][q];
            }
    for (int i = 0; i < n; ++i)
        for (int j = n; j < 2 * n; ++j) A[n * i + j - n] = aug[i][j] / aug[i][i];
}

// rel4 = mendrel[1..4]: strong/weak axis releases at end 1, end 2 -> DOFs 5, 4, 12, 11
__device__ __forceinline__ void frame_release(double (*k)[14], const int *rel4)
{
    const int dof[4] = {5, 4, 12, 11};
    int idx[4], r = 0;
    for (int i = 0; i < 4; ++i)
        if (rel4[i] == 1) idx[r++] = dof[i];
    if (r == 0) return;
    double kG[14][4], GkG[16], W[14][4];
    for (int i = 0; i < 14; ++i)
        for (int j = 0; j < r; ++j) kG[i][j] = 0 + k[i][idx[j]];
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < r; ++j) GkG[i * r + j] = kG[idx[j]][i];
    if (r == 1) GkG[0] = 1 / GkG[0];
    else if (r == 2) {
        const double det = GkG[0] * GkG[3] - GkG[1] * GkG[2], t0 = GkG[0];
        GkG[0] = GkG[3] / det; GkG[3] = t0 / det; GkG[1] *= -1 / det; GkG[2] *= -1 / det;
    } else gj_inverse4(GkG, r);
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

// full local tangent of frame e for ANAFLAG 1 / 2; eftot receives ef_ip + efFE_ip
__device__ __forceinline__ void frame_local_k(const CbDev &d, long e, const double *ef_ip,
                                              const double *efFE_ip, double defllen_ip,
                                              double (*k)[14], double *eftot)
{
    const double *fc = d.fr_const + e * CB_FR_CONST;
    for (int i = 0; i < 14; ++i) {
        for (int j = 0; j < 14; ++j) k[i][j] = 0;
        eftot[i] = ef_ip[e * 14 + i] + efFE_ip[e * 14 + i];
    }
    frame_elastic(k, fc);
    if (d.ANAFLAG >= 2) frame_geometric(k, eftot, defllen_ip, fc[2], fc[8]);
    if (d.ANAFLAG == 3) {                            // frame.c:268-277
        const int y0 = d.fr_yldflag[e * 2], y1 = d.fr_yldflag[e * 2 + 1];
        if (y0 != 2 || y1 != 2) frame_plastic(k, eftot, y0, y1, d.fr_plast + e * 3);
    }
    if (d.fr_mendrel[e * 5] == 1) frame_release(k, d.fr_mendrel + e * 5 + 1);
}

#endif
