/* cb_sparse.c - the solver side of the CSC hand-off.
 *
 * The reference's SLVFLAG==2 path scans a dense NEQ^2 array into Ap/Ai/Ax and calls
 * umfpack_di_symbolic / _numeric / _solve on it (solve.c:107-135, 199-243).  With the device path
 * the host receives Ap/Ai (once) and Ax (every refactorisation) straight from cb_csc_pattern /
 * cb_get_csc_values, so that scan - and the NEQ^2 allocations of main.c:1323-1334 - go away.
 *
 * Two back ends behind one interface:
 *   - UMFPACK, the reference's own call sequence, compiled in with -DCB_HAVE_UMFPACK (SuiteSparse is
 *     not in this image, so that branch is built only where umfpack.h exists);
 *   - a built-in sparse LDL^T (up-looking, elimination-tree based, natural ordering, no pivoting) -
 *     the sparse counterpart of the reference's skyline LDL^T (skyfact / skysolve, solve.c:539-698):
 *     same factors D and L up to rounding, so the pivots and the sign of the determinant the
 *     arc-length driver needs (solve.c:563-572) come out the same way.
 * K_t is symmetric; the factorisation reads the upper triangle of the full (unsymmetric-storage)
 * CSC matrix that UMFPACK would be given.
 *
 * Also here: K v on the full CSC (skymult, solve.c:700-756), the diagonal addresses for
 * Keff = K + c M (solve.c:199-207) and matpart() on the CSC (solve.c:758-824). */
#include "cb_host.h"
#include <stdlib.h>
#include <string.h>
#ifdef CB_HAVE_UMFPACK
#include <umfpack.h>
#endif

struct cb_csc_solver {
    long n;
    const int *Ap, *Ai;          /* caller-owned pattern (full matrix, sorted or unsorted rows) */
    long *parent;                /* elimination tree                                            */
    long *Lp, *Lnz, *Li;         /* L by columns, unit diagonal not stored                      */
    double *Lx, *D, *Y;
    long *pattern, *flag;
    int factored;
#ifdef CB_HAVE_UMFPACK
    void *Symbolic, *Numeric;
    const double *Ax_last;
    double *tmp;
#endif
};

void cb_csc_solver_destroy(cb_csc_solver *s)
{
    if (!s) return;
#ifdef CB_HAVE_UMFPACK
    if (s->Numeric) umfpack_di_free_numeric(&s->Numeric);
    if (s->Symbolic) umfpack_di_free_symbolic(&s->Symbolic);
    free(s->tmp);
#endif
    free(s->parent); free(s->Lp); free(s->Lnz); free(s->Li); free(s->Lx); free(s->D); free(s->Y);
    free(s->pattern); free(s->flag);
    free(s);
}

/* symbolic analysis: elimination tree and column counts of L from the upper triangle of A */
int cb_csc_solver_create(long n, const int *Ap, const int *Ai, cb_csc_solver **out)
{
    if (!out || n < 0 || !Ap || !Ai) return CB_ERR_ARG;
    cb_csc_solver *s = (cb_csc_solver *)calloc(1, sizeof *s);
    if (!s) return CB_ERR_ARG;
    s->n = n; s->Ap = Ap; s->Ai = Ai;
    s->parent = (long *)malloc((size_t)(n + 1) * sizeof(long));
    s->Lp = (long *)malloc((size_t)(n + 1) * sizeof(long));
    s->Lnz = (long *)calloc((size_t)(n + 1), sizeof(long));
    s->D = (double *)malloc((size_t)(n + 1) * sizeof(double));
    s->Y = (double *)calloc((size_t)(n + 1), sizeof(double));
    s->pattern = (long *)malloc((size_t)(n + 1) * sizeof(long));
    s->flag = (long *)malloc((size_t)(n + 1) * sizeof(long));
    if (!s->parent || !s->Lp || !s->Lnz || !s->D || !s->Y || !s->pattern || !s->flag) {
        cb_csc_solver_destroy(s);
        return CB_ERR_ARG;
    }
    for (long k = 0; k < n; ++k) {
        s->parent[k] = -1;
        s->flag[k] = k;
        for (long p = Ap[k]; p < Ap[k + 1]; ++p) {
            long i = Ai[p];
            if (i >= k) continue;
            for (; s->flag[i] != k; i = s->parent[i]) {     /* walk up the tree from i to k */
                if (s->parent[i] == -1) s->parent[i] = k;
                ++s->Lnz[i];
                s->flag[i] = k;
            }
        }
    }
    s->Lp[0] = 0;
    for (long k = 0; k < n; ++k) s->Lp[k + 1] = s->Lp[k] + s->Lnz[k];
    const long lnz = s->Lp[n] > 0 ? s->Lp[n] : 1;
    s->Li = (long *)malloc((size_t)lnz * sizeof(long));
    s->Lx = (double *)malloc((size_t)lnz * sizeof(double));
    if (!s->Li || !s->Lx) { cb_csc_solver_destroy(s); return CB_ERR_ARG; }
#ifdef CB_HAVE_UMFPACK
    s->tmp = (double *)malloc((size_t)(n + 1) * sizeof(double));
    /* solve.c:122 */
    if (umfpack_di_symbolic((int)n, (int)n, Ap, Ai, NULL, &s->Symbolic, NULL, NULL) != UMFPACK_OK) {
        cb_csc_solver_destroy(s);
        return CB_ERR_ARG;
    }
#endif
    *out = s;
    return CB_OK;
}

long cb_csc_solver_lnz(const cb_csc_solver *s) { return s ? s->Lp[s->n] : 0; }

/* numeric LDL^T.  allow_indefinite == 0: a non-positive pivot fails (solve.c:596-604, "Non-positive
 * definite stiffness matrix"); != 0: only a zero pivot fails, *det_neg is set when any pivot is
 * negative and pivots (may be NULL) receives D (the ALGFLAG==3 branch, solve.c:563-572).  0 / 1. */
int cb_csc_solver_factor(cb_csc_solver *s, const double *Ax, int allow_indefinite, int *det_neg,
                         double *pivots)
{
    const long n = s->n;
    const int *Ap = s->Ap, *Ai = s->Ai;
    double *Y = s->Y, *D = s->D, *Lx = s->Lx;
    long *Li = s->Li, *Lp = s->Lp, *Lnz = s->Lnz, *pattern = s->pattern, *flag = s->flag;
    if (det_neg) *det_neg = 0;
    s->factored = 0;
#ifdef CB_HAVE_UMFPACK
    if (!allow_indefinite && !pivots) {
        if (s->Numeric) umfpack_di_free_numeric(&s->Numeric);
        /* solve.c:124 */
        if (umfpack_di_numeric(Ap, Ai, Ax, s->Symbolic, &s->Numeric, NULL, NULL) != UMFPACK_OK) return 1;
        s->Ax_last = Ax;
        s->factored = 2;
        return 0;
    }
#endif
    for (long k = 0; k < n; ++k) {
        /* nonzero pattern of row k of L, in topological order, and the scattered column of A */
        long top = n;
        flag[k] = k;
        Lnz[k] = 0;
        Y[k] = 0;
        for (long p = Ap[k]; p < Ap[k + 1]; ++p) {
            long i = Ai[p];
            if (i > k) continue;
            Y[i] += Ax[p];
            long len = 0;
            for (; flag[i] != k; i = s->parent[i]) {
                pattern[len++] = i;
                flag[i] = k;
            }
            while (len > 0) pattern[--top] = pattern[--len];
        }
        double dk = Y[k];
        Y[k] = 0;
        for (; top < n; ++top) {
            const long i = pattern[top];
            const double yi = Y[i];
            Y[i] = 0;
            const long p2 = Lp[i] + Lnz[i];
            for (long p = Lp[i]; p < p2; ++p) Y[Li[p]] -= Lx[p] * yi;
            const double lki = yi / D[i];
            dk -= lki * yi;
            Li[p2] = k;
            Lx[p2] = lki;
            ++Lnz[i];
        }
        D[k] = dk;
        if (pivots) pivots[k] = dk;
        if (!allow_indefinite) {
            if (dk <= 0) return 1;
        } else {
            if (dk == 0) return 1;
            if (dk < 0 && det_neg) *det_neg = 1;
        }
    }
    s->factored = 1;
    return 0;
}

/* rhs <- K^-1 rhs with the last factorisation */
int cb_csc_solver_solve(cb_csc_solver *s, double *x)
{
    if (!s || !s->factored) return CB_ERR_ARG;
    const long n = s->n;
#ifdef CB_HAVE_UMFPACK
    if (s->factored == 2) {
        /* solve.c:126: x in tmp; the reference forgets the copy back into dd */
        if (umfpack_di_solve(UMFPACK_A, s->Ap, s->Ai, s->Ax_last, s->tmp, x, s->Numeric, NULL, NULL) != UMFPACK_OK)
            return CB_ERR_ARG;
        memcpy(x, s->tmp, (size_t)n * sizeof(double));
        return CB_OK;
    }
#endif
    const long *Lp = s->Lp, *Li = s->Li, *Lnz = s->Lnz;
    const double *Lx = s->Lx, *D = s->D;
    for (long j = 0; j < n; ++j) {                         /* L y = b   */
        const double xj = x[j];
        const long p2 = Lp[j] + Lnz[j];
        for (long p = Lp[j]; p < p2; ++p) x[Li[p]] -= Lx[p] * xj;
    }
    for (long j = 0; j < n; ++j) x[j] /= D[j];             /* D z = y   */
    for (long j = n - 1; j >= 0; --j) {                    /* L^T x = z */
        double xj = x[j];
        const long p2 = Lp[j] + Lnz[j];
        for (long p = Lp[j]; p < p2; ++p) xj -= Lx[p] * x[Li[p]];
        x[j] = xj;
    }
    return CB_OK;
}

/* v <- K v on the full CSC matrix (skymult, solve.c:700-756); tmp [n] scratch */
void cb_csc_mult(long n, const int *Ap, const int *Ai, const double *Ax, double *v, double *tmp)
{
    for (long i = 0; i < n; ++i) tmp[i] = 0;
    for (long j = 0; j < n; ++j) {
        const double vj = v[j];
        for (long p = Ap[j]; p < Ap[j + 1]; ++p) tmp[Ai[p]] += Ax[p] * vj;
    }
    memcpy(v, tmp, (size_t)n * sizeof(double));
}

/* position of the diagonal entry of every column (-1 when the pattern has none) */
void cb_csc_diag(long n, const int *Ap, const int *Ai, long *diag)
{
    for (long j = 0; j < n; ++j) {
        diag[j] = -1;
        for (long p = Ap[j]; p < Ap[j + 1]; ++p)
            if (Ai[p] == j) { diag[j] = p; break; }
    }
}

/* matpart() on the CSC (solve.c:758-824): the equations with pmot != 0 carry prescribed motion uc;
 * their columns go to the right-hand side of the free rows, their rows and columns leave the
 * matrix (unit diagonal). */
void cb_csc_partition(long n, const int *Ap, const int *Ai, double *Ax, double *qtot, const double *uc,
                      const int *pmot)
{
    for (long j = 0; j < n; ++j)
        for (long p = Ap[j]; p < Ap[j + 1]; ++p) {
            const long i = Ai[p];
            if (pmot[j]) {
                if (pmot[i]) Ax[p] = (i == j) ? 1.0 : 0.0;
                else { qtot[i] -= Ax[p] * uc[j]; Ax[p] = 0; }
            } else if (pmot[i]) {
                Ax[p] = 0;
            }
        }
}

/* ---- cb_lin: what the drivers see - the skyline of SLVFLAG 0 or the CSC of SLVFLAG 2 ----------- */
struct cb_lin {
    int csc;
    long neq, nval;              /* nval = lss or nnz                              */
    const long *maxa;            /* skyline                                        */
    int *Ap, *Ai;                /* CSC pattern (owned)                            */
    long *diag;
    cb_csc_solver *S;
    double *tmp;
};

void cb_lin_destroy(cb_lin *L)
{
    if (!L) return;
    cb_csc_solver_destroy(L->S);
    free(L->Ap); free(L->Ai); free(L->diag); free(L->tmp);
    free(L);
}

/* maxa != NULL: skyline (maxa, lss as skylin() produced them); maxa == NULL: CSC, pattern taken
 * from the handle (the handle must have been created with a CSC layout) */
int cb_lin_create(cb_handle *h, long neq, const long *maxa, long lss, cb_lin **out)
{
    cb_lin *L = (cb_lin *)calloc(1, sizeof *L);
    if (!L) return CB_ERR_ARG;
    L->neq = neq;
    if (maxa) {
        L->maxa = maxa; L->nval = lss;
        *out = L;
        return CB_OK;
    }
    L->csc = 1;
    L->nval = cb_csc_nnz(h);
    if (L->nval <= 0) { free(L); return CB_ERR_ARG; }
    L->Ap = (int *)malloc((size_t)(neq + 1) * sizeof(int));
    L->Ai = (int *)malloc((size_t)L->nval * sizeof(int));
    L->diag = (long *)malloc((size_t)(neq + 1) * sizeof(long));
    L->tmp = (double *)malloc((size_t)(neq + 1) * sizeof(double));
    int rc = CB_ERR_ARG;
    if (L->Ap && L->Ai && L->diag && L->tmp && (rc = cb_csc_pattern(h, L->Ap, L->Ai)) == CB_OK &&
        (rc = cb_csc_solver_create(neq, L->Ap, L->Ai, &L->S)) == CB_OK) {
        cb_csc_diag(neq, L->Ap, L->Ai, L->diag);
        for (long i = 0; i < neq; ++i)
            if (L->diag[i] < 0) rc = CB_ERR_ARG;
    }
    if (rc != CB_OK) { cb_lin_destroy(L); return rc; }
    *out = L;
    return CB_OK;
}

long cb_lin_nval(const cb_lin *L) { return L->nval; }
int  cb_lin_is_scalar(const cb_lin *L) { return !L->csc && L->nval == 1; }

/* K_t of the last cb_stiff into K [nval] */
int cb_lin_fetch(cb_lin *L, cb_handle *h, double *K)
{
    return L->csc ? cb_get_csc_values(h, K) : cb_get_skyline(h, K, L->nval);
}

/* K [nval] += diag(num * m / den), evaluated in that order (solve.c:199-207) */
void cb_lin_add_diag(const cb_lin *L, double *K, double num, double den, const double *m)
{
    if (L->csc) for (long i = 0; i < L->neq; ++i) K[L->diag[i]] += num * m[i] / den;
    else for (long i = 0; i < L->neq; ++i) K[L->maxa[i] - 1] += num * m[i] / den;
}

double cb_lin_diag0(const cb_lin *L, const double *K) { return L->csc ? K[L->diag[0]] : K[0]; }

/* factorise K in place (skyline) or into the solver's own storage (CSC; K is kept) */
int cb_lin_factor(cb_lin *L, double *K, double *pivots, int *det_neg, int allow_indefinite)
{
    if (L->csc) return cb_csc_solver_factor(L->S, K, allow_indefinite, det_neg, pivots);
    return cb_sky_factor(L->neq, L->maxa, K, pivots, det_neg, allow_indefinite);
}

void cb_lin_solve(cb_lin *L, const double *K, double *rhs)
{
    if (L->csc) cb_csc_solver_solve(L->S, rhs);
    else cb_sky_solve(L->neq, L->maxa, K, rhs);
}

void cb_lin_mult(cb_lin *L, const double *K, double *v)
{
    if (L->csc) cb_csc_mult(L->neq, L->Ap, L->Ai, K, v, L->tmp);
    else cb_sky_mult(L->neq, L->maxa, K, v);
}

void cb_lin_partition(cb_lin *L, double *K, double *qtot, const double *uc, const int *pmot, long nbc,
                      const long *ii, const long *ij)
{
    if (L->csc) cb_csc_partition(L->neq, L->Ap, L->Ai, K, qtot, uc, pmot);
    else cb_sky_partition(L->neq, nbc, L->maxa, K, qtot, uc, ii, ij);
}
