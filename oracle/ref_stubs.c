/* TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * Link-time stand-ins for the third-party solvers the reference's solve.c calls but this image
 * does not ship: UMFPACK (solve.c:122-134,240-242,412,461), LAPACK dgesv_/dgetrf_/dgetrs_
 * (solve.c:95,224,389,397,409,516,526) and cblas_dgemv (solve.c:363).  None of them is on the
 * element/assembly path; every shipped sample deck uses SLVFLAG=0 (the reference's own skyline
 * LDL^T), which never reaches these symbols.
 *
 * umfpack_di_symbolic() CAPTURES the CSC arrays it receives so tests can read back exactly
 * what solve.c:110-119 built from the dense matrix; umfpack_di_solve() then solves the system
 * with a small dense Gaussian elimination so the SLVFLAG=2 static branch returns a usable
 * answer in puc (sizes in the tests are tiny). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "umfpack.h"

static int     cap_n = 0;
static long    cap_nnz = 0;
static int    *cap_Ap = NULL, *cap_Ai = NULL;
static double *cap_Ax = NULL;

int umfpack_di_symbolic(int n_row, int n_col, const int *Ap, const int *Ai, const double *Ax,
                        void **Symbolic, const double *Control, double *Info)
{
    (void)n_col; (void)Control; (void)Info;
    free(cap_Ap); free(cap_Ai); free(cap_Ax);
    cap_n = n_row;
    cap_nnz = Ap[n_row];
    cap_Ap = (int *)malloc(sizeof(int) * (size_t)(n_row + 1));
    cap_Ai = (int *)malloc(sizeof(int) * (size_t)(cap_nnz > 0 ? cap_nnz : 1));
    cap_Ax = (double *)malloc(sizeof(double) * (size_t)(cap_nnz > 0 ? cap_nnz : 1));
    memcpy(cap_Ap, Ap, sizeof(int) * (size_t)(n_row + 1));
    memcpy(cap_Ai, Ai, sizeof(int) * (size_t)cap_nnz);
    memcpy(cap_Ax, Ax, sizeof(double) * (size_t)cap_nnz);
    if (Symbolic) *Symbolic = NULL;
    return 0;
}

int umfpack_di_numeric(const int *Ap, const int *Ai, const double *Ax, void *Symbolic,
                       void **Numeric, const double *Control, double *Info)
{
    (void)Ap; (void)Ai; (void)Ax; (void)Symbolic; (void)Control; (void)Info;
    if (Numeric) *Numeric = NULL;
    return 0;
}

/* Dense partial-pivot solve of the captured column-compressed system (test sizes only). */
int umfpack_di_solve(int sys, const int *Ap, const int *Ai, const double *Ax, double *X,
                     const double *B, void *Numeric, const double *Control, double *Info)
{
    (void)sys; (void)Numeric; (void)Control; (void)Info;
    int n = cap_n, i, j, k;
    double *A = (double *)calloc((size_t)n * (size_t)n, sizeof(double));
    double *b = (double *)malloc(sizeof(double) * (size_t)n);
    if (!A || !b) { free(A); free(b); return -1; }
    for (j = 0; j < n; ++j)
        for (k = Ap[j]; k < Ap[j + 1]; ++k)
            A[(size_t)Ai[k] * n + j] = Ax[k];
    memcpy(b, B, sizeof(double) * (size_t)n);
    for (k = 0; k < n; ++k) {
        int p = k; double big = fabs(A[(size_t)k * n + k]);
        for (i = k + 1; i < n; ++i)
            if (fabs(A[(size_t)i * n + k]) > big) { big = fabs(A[(size_t)i * n + k]); p = i; }
        if (p != k) {
            for (j = 0; j < n; ++j) {
                double t = A[(size_t)k * n + j]; A[(size_t)k * n + j] = A[(size_t)p * n + j];
                A[(size_t)p * n + j] = t;
            }
            double t = b[k]; b[k] = b[p]; b[p] = t;
        }
        for (i = k + 1; i < n; ++i) {
            double m = A[(size_t)i * n + k] / A[(size_t)k * n + k];
            if (m == 0.0) continue;
            for (j = k; j < n; ++j) A[(size_t)i * n + j] -= m * A[(size_t)k * n + j];
            b[i] -= m * b[k];
        }
    }
    for (i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (j = i + 1; j < n; ++j) s -= A[(size_t)i * n + j] * X[j];
        X[i] = s / A[(size_t)i * n + i];
    }
    free(A); free(b);
    return 0;
}

void umfpack_di_free_symbolic(void **Symbolic) { if (Symbolic) *Symbolic = NULL; }
void umfpack_di_free_numeric(void **Numeric)   { if (Numeric)  *Numeric  = NULL; }

/* Accessors used by tests (ctypes). */
int    ref_csc_n(void)   { return cap_n; }
long   ref_csc_nnz(void) { return cap_nnz; }
void   ref_csc_copy(int *Ap, int *Ai, double *Ax)
{
    memcpy(Ap, cap_Ap, sizeof(int) * (size_t)(cap_n + 1));
    memcpy(Ai, cap_Ai, sizeof(int) * (size_t)cap_nnz);
    memcpy(Ax, cap_Ax, sizeof(double) * (size_t)cap_nnz);
}

/* LAPACK / CBLAS entry points: only reachable with SLVFLAG=1 or the FSI dynamic branch, which
 * the oracle never selects.  Fail loudly if that ever changes. */
static void unreachable(const char *name)
{
    fprintf(stderr, "oracle/ref_stubs.c: %s called - SLVFLAG=1 / FSI paths are not part of "
                    "the oracle\n", name);
    abort();
}
void dgesv_(void)  { unreachable("dgesv_");  }
void dgetrf_(void) { unreachable("dgetrf_"); }
void dgetrs_(void) { unreachable("dgetrs_"); }
void cblas_dgemv(void) { unreachable("cblas_dgemv"); }
