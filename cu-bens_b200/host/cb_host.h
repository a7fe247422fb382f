/* cb_host.h - C host drivers that sit where main.c's loops sit and drive the device path
 * through the C-ABI of include/cubens_b200.h.  The reference's host stays C (north_star); this
 * is the static Newton-Raphson / modified Newton-Raphson load-increment loop of
 * main.c:1824-2152 written against cb_*, with the skyline LDL^T solve that solve.c:78-80 does
 * (skyfact/skysolve, solve.c:539-698 - Bathe's COLSOL) and the convergence test of
 * misc.c:187-250.  "Next" rows 1-2 of SURVEY.md section 8(f): host-side consumers of the path. */
#ifndef CB_HOST_H
#define CB_HOST_H
#include "../../include/cubens_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* active-column (skyline) LDL^T, layout of model.c:1269-1278: maxa 1-based diagonal addresses.
 * allow_indefinite != 0 reproduces the ALGFLAG==3 branch (pivots returned in ssd, *det_neg set
 * when a pivot is negative) instead of failing on a non-positive pivot.  Returns 0 / 1.       */
int  cb_sky_factor(long neq, const long *maxa, double *ss, double *ssd, int *det_neg,
                   int allow_indefinite);
void cb_sky_solve(long neq, const long *maxa, const double *ss, double *rhs);

/* ---- the CSC hand-off to the solver (cb_sparse.c) = solve.c:107-135, 199-243 without the dense
 * NEQ^2 scan: Ap / Ai once from cb_csc_pattern, Ax from cb_get_csc_values at every refactorisation.
 * Back ends: UMFPACK (the reference's calls, -DCB_HAVE_UMFPACK) or the built-in sparse LDL^T
 * (natural ordering, no pivoting - the sparse twin of skyfact / skysolve, solve.c:539-698, so the
 * pivots and determinant sign of the arc-length branch, solve.c:563-572, are available).        */
typedef struct cb_csc_solver cb_csc_solver;
int  cb_csc_solver_create(long n, const int *Ap, const int *Ai, cb_csc_solver **out);
int  cb_csc_solver_factor(cb_csc_solver *s, const double *Ax, int allow_indefinite, int *det_neg,
                          double *pivots);                       /* 0 / 1 like skyfact            */
int  cb_csc_solver_solve(cb_csc_solver *s, double *rhs);
long cb_csc_solver_lnz(const cb_csc_solver *s);
void cb_csc_solver_destroy(cb_csc_solver *s);
void cb_csc_mult(long n, const int *Ap, const int *Ai, const double *Ax, double *v, double *tmp);
void cb_csc_diag(long n, const int *Ap, const int *Ai, long *diag);
void cb_csc_partition(long n, const int *Ap, const int *Ai, double *Ax, double *qtot, const double *uc,
                      const int *pmot);

/* what the drivers below factorise: the skyline of SLVFLAG 0 (maxa != NULL) or the device-built CSC
 * of SLVFLAG 2 (maxa == NULL; the handle must own a CSC layout).  Every driver takes (maxa, lss):
 * pass maxa = NULL, lss = 0 to run it on the CSC hand-off.                                        */
typedef struct cb_lin cb_lin;
int    cb_lin_create(cb_handle *h, long neq, const long *maxa, long lss, cb_lin **out);
void   cb_lin_destroy(cb_lin *L);
long   cb_lin_nval(const cb_lin *L);
int    cb_lin_is_scalar(const cb_lin *L);
int    cb_lin_fetch(cb_lin *L, cb_handle *h, double *K);
void   cb_lin_add_diag(const cb_lin *L, double *K, double num, double den, const double *m);
double cb_lin_diag0(const cb_lin *L, const double *K);
int    cb_lin_factor(cb_lin *L, double *K, double *pivots, int *det_neg, int allow_indefinite);
void   cb_lin_solve(cb_lin *L, const double *K, double *rhs);
void   cb_lin_mult(cb_lin *L, const double *K, double *v);
void   cb_lin_partition(cb_lin *L, double *K, double *qtot, const double *uc, const int *pmot, long nbc,
                        const long *ii, const long *ij);

/* the solver controls main.c reads after the loads (main.c:1809-1812) */
typedef struct cb_nr_params {
    double lpfmax, lpf, dlpf, dlpfmax, dlpfmin;
    int itemax, submax, solmin;
    double toldisp, tolforc, tolener;
    int algflag;                 /* 1 = Newton-Raphson, 2 = modified Newton-Raphson */
} cb_nr_params;

typedef struct cb_nr_result {
    int status;                  /* 0 = "Solution successful", else the reference's error exit   */
    int increments;              /* converged load increments                                    */
    int iterations;              /* total equilibrium iterations                                 */
    int stiff_calls, force_calls;
    double lpf;                  /* load proportionality factor of the last converged increment  */
} cb_nr_result;

/* q [NEQ] reference load vector; d_out [NEQ] displacements of the last converged increment;
 * hist (may be NULL) receives up to max_hist rows (lpf, iterations, d[hist_dof]) like
 * results2.txt.  maxa / lss as skylin() produced them.                                       */
int cb_newton_static(cb_handle *h, long neq, const long *maxa, long lss, const double *q,
                     const cb_nr_params *p, double *d_out, cb_nr_result *res, double *hist,
                     int max_hist, long hist_dof);

/* ---- transient drivers (cb_newmark.c) --------------------------------------------------------
 * pinpt [NEQ][ntstps]: the load histories load() stores (model.c:1490-1535), dt = ttot / (ntstps-1),
 * alpham / alphaf: generalized-alpha parameters (main.c:3337-3351).  hist receives one row per time
 * step: time, iterations, d[0..NEQ) - what the reference hands to output().                      */
void cb_sky_mult(long neq, const long *maxa, const double *ss, double *v);   /* solve.c:700-756 */
/* matpart() (solve.c:758-824): ii / ij = equations without / with prescribed motion (0-based)  */
void cb_sky_partition(long neq, long nbc, const long *maxa, double *ss, double *qtot, const double *uc,
                      const long *ii, const long *ij);
int cb_newmark_nonlinear(cb_handle *h, long neq, const long *maxa, long lss, const double *pinpt,
                         long ntstps, double dt, double alpham, double alphaf,
                         const cb_nr_params *p, double *hist, cb_nr_result *res);
/* same with prescribed support motion: pdisp [NEQ][ntstps] (model.c:1537-1572) and pmot [NEQ] = 1 at
 * the equations listed as moved supports (skylin(), model.c:1149-1201); the effective matrix is
 * partitioned as matpart() does (solve.c:758-824).  pdisp / pmot may be NULL (no support motion). */
int cb_newmark_nonlinear_bc(cb_handle *h, long neq, const long *maxa, long lss, const double *pinpt,
                            const double *pdisp, const int *pmot, long ntstps, double dt, double alpham,
                            double alphaf, const cb_nr_params *p, double *hist, cb_nr_result *res);
int cb_newmark_linear(cb_handle *h, long neq, const long *maxa, long lss, const double *pinpt,
                      long ntstps, double dt, double alpham, double alphaf, const double *um0,
                      const double *vm0, const double *am0, double *hist, cb_nr_result *res);

/* ---- modified spherical arc-length driver (cb_arclength.c) = main.c:2158-3141 (ALGFLAG 3) --------
 * Bathe & Dvorkin's arc-length iteration with Crisfield & Shi's psi factor: first increment by a
 * prescribed displacement dk at equation dkdof (0-based, msal() arc.c:41-68), then arc-length
 * controlled increments until |lpf| > lpfmax or |d[dkdof]| > dkimax.  K_t is assembled from the
 * committed state once per increment (cb_stiff(CB_GEN_COMMITTED)), factorised with the indefinite
 * skyline LDL^T (pivots -> psi, sign of the determinant -> direction), f_int every iteration.     */
typedef struct cb_arc_params {
    double dk; long dkdof;
    double alpha, psi_thresh; int iteopt;
    double lpfmax, dkimax;
    int itemax, submax, imagmax, negmax;
    double toldisp, tolforc, tolener;
} cb_arc_params;
/* hist: up to max_rows rows (lpf, iterations, d[0..NEQ)); res->increments = rows written */
int cb_arclength_static(cb_handle *h, long neq, const long *maxa, long lss, const double *q,
                        const cb_arc_params *p, double *hist, int max_rows, cb_nr_result *res);

#ifdef __cplusplus
}
#endif
#endif
