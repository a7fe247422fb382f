/* dense-LU stand-in for the five UMFPACK entry points (see umfpack.h next to this file); counts its calls */
#include <stdlib.h>
#include <math.h>
#include "umfpack.h"

typedef struct { int n; } Sym;
typedef struct { int n; double *LU; int *piv; } Num;
int umfpack_double_calls[5];

int umfpack_di_symbolic(int n_row, int n_col, const int Ap[], const int Ai[], const double Ax[], void **Symbolic,
                        const double Control[], double Info[])
{
    (void)Ap; (void)Ai; (void)Ax; (void)Control; (void)Info;
    ++umfpack_double_calls[0];
    if (n_row != n_col) return -1;
    Sym *s = (Sym *)malloc(sizeof *s); s->n = n_row; *Symbolic = s;
    return UMFPACK_OK;
}

int umfpack_di_numeric(const int Ap[], const int Ai[], const double Ax[], void *Symbolic, void **Numeric,
                       const double Control[], double Info[])
{
    (void)Control; (void)Info;
    ++umfpack_double_calls[1];
    const int n = ((Sym *)Symbolic)->n;
    Num *m = (Num *)malloc(sizeof *m);
    m->n = n; m->LU = (double *)calloc((size_t)n * n, sizeof(double)); m->piv = (int *)malloc(n * sizeof(int));
    for (int j = 0; j < n; ++j)
        for (int p = Ap[j]; p < Ap[j + 1]; ++p) m->LU[(size_t)Ai[p] * n + j] += Ax[p];
    for (int k = 0; k < n; ++k) {
        int r = k;
        for (int i = k + 1; i < n; ++i) if (fabs(m->LU[(size_t)i * n + k]) > fabs(m->LU[(size_t)r * n + k])) r = i;
        m->piv[k] = r;
        if (m->LU[(size_t)r * n + k] == 0.0) { free(m->LU); free(m->piv); free(m); return UMFPACK_WARNING_singular_matrix; }
        if (r != k) for (int j = 0; j < n; ++j) { double t = m->LU[(size_t)k * n + j]; m->LU[(size_t)k * n + j] = m->LU[(size_t)r * n + j]; m->LU[(size_t)r * n + j] = t; }
        for (int i = k + 1; i < n; ++i) {
            const double l = m->LU[(size_t)i * n + k] /= m->LU[(size_t)k * n + k];
            for (int j = k + 1; j < n; ++j) m->LU[(size_t)i * n + j] -= l * m->LU[(size_t)k * n + j];
        }
    }
    *Numeric = m;
    return UMFPACK_OK;
}

int umfpack_di_solve(int sys, const int Ap[], const int Ai[], const double Ax[], double X[], const double B[],
                     void *Numeric, const double Control[], double Info[])
{
    (void)Ap; (void)Ai; (void)Ax; (void)Control; (void)Info;
    ++umfpack_double_calls[2];
    if (sys != UMFPACK_A) return -1;
    const Num *m = (const Num *)Numeric;
    const int n = m->n;
    for (int i = 0; i < n; ++i) X[i] = B[i];
    for (int k = 0; k < n; ++k) {
        if (m->piv[k] != k) { double t = X[k]; X[k] = X[m->piv[k]]; X[m->piv[k]] = t; }
        for (int i = k + 1; i < n; ++i) X[i] -= m->LU[(size_t)i * n + k] * X[k];
    }
    for (int i = n - 1; i >= 0; --i) {
        for (int j = i + 1; j < n; ++j) X[i] -= m->LU[(size_t)i * n + j] * X[j];
        X[i] /= m->LU[(size_t)i * n + i];
    }
    return UMFPACK_OK;
}

void umfpack_di_free_symbolic(void **Symbolic) { ++umfpack_double_calls[3]; free(*Symbolic); *Symbolic = 0; }
void umfpack_di_free_numeric(void **Numeric)
{
    ++umfpack_double_calls[4];
    Num *m = (Num *)*Numeric;
    if (m) { free(m->LU); free(m->piv); free(m); }
    *Numeric = 0;
}
