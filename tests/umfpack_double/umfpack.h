/* TEST DOUBLE of SuiteSparse's umfpack.h (SuiteSparse is not in this image): the five umfpack_di_* entry
 * points solve.c:107-135 uses, with their published prototypes, implemented in umfpack_double.c by a dense
 * LU with partial pivoting.  It exists so that the -DCB_HAVE_UMFPACK branch of cu-bens_b200/host/cb_sparse.c
 * is compiled, linked and run by tests/test_umfpack_binding_cpu.py; it is not a sparse solver. */
#ifndef CB_TEST_UMFPACK_DOUBLE_H
#define CB_TEST_UMFPACK_DOUBLE_H
#define UMFPACK_OK 0
#define UMFPACK_A 0
#define UMFPACK_WARNING_singular_matrix 1
int  umfpack_di_symbolic(int n_row, int n_col, const int Ap[], const int Ai[], const double Ax[],
                         void **Symbolic, const double Control[], double Info[]);
int  umfpack_di_numeric(const int Ap[], const int Ai[], const double Ax[], void *Symbolic,
                        void **Numeric, const double Control[], double Info[]);
int  umfpack_di_solve(int sys, const int Ap[], const int Ai[], const double Ax[], double X[],
                      const double B[], void *Numeric, const double Control[], double Info[]);
void umfpack_di_free_symbolic(void **Symbolic);
void umfpack_di_free_numeric(void **Numeric);
#endif
