/* Stand-in for SuiteSparse's umfpack.h so that the reference's solve.c (#include "umfpack.h",
 * solve.c:43) compiles in this image, which has no SuiteSparse.  TEST INFRASTRUCTURE ONLY.
 * The five entry points are defined in oracle/ref_stubs.c; umfpack_di_symbolic() records the
 * CSC arrays it is handed (that is "what solve.c hands to SuiteSparse", solve.c:122) so the
 * parity tests can compare them with the device-assembled CSC. */
#ifndef CB_ORACLE_UMFPACK_STUB_H
#define CB_ORACLE_UMFPACK_STUB_H
#define UMFPACK_A 0
int  umfpack_di_symbolic(int n_row, int n_col, const int *Ap, const int *Ai, const double *Ax,
                         void **Symbolic, const double *Control, double *Info);
int  umfpack_di_numeric(const int *Ap, const int *Ai, const double *Ax, void *Symbolic,
                        void **Numeric, const double *Control, double *Info);
int  umfpack_di_solve(int sys, const int *Ap, const int *Ai, const double *Ax, double *X,
                      const double *B, void *Numeric, const double *Control, double *Info);
void umfpack_di_free_symbolic(void **Symbolic);
void umfpack_di_free_numeric(void **Numeric);
#endif
