/* cubens_oracle.h - TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's per-iteration element / assembly hot path, written
 * from the reference's algorithm (file:line cited at every routine in cubens_oracle.c), with
 * the reference's host array layout so the same numpy arrays can be handed to this library, to
 * the compiled reference (oracle/_ref) and to the CUDA path.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks every routine here against
 *   (a) the unmodified reference compiled from /root/reference (oracle/_ref), bit for bit, on
 *       generated meshes, and
 *   (b) the committed golden vectors under tests/golden/ (made by tests/golden/make_golden.py
 *       from oracle/_ref; the reference ships no tests or expected outputs of its own).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#ifndef CUBENS_ORACLE_H
#define CUBENS_ORACLE_H

typedef struct orc_dims {
    long NJ, NE_TR, NE_FR, NE_SH, NE_BR, NEQ;
    int ANAFLAG, SLVFLAG;
} orc_dims;

/* model.c:937-1142 / 1204-1281 */
long orc_codes(const orc_dims *D, long *mcode, long *jcode, const long *minc);
long orc_skylin(const orc_dims *D, long *maxa, long *kht, const long *mcode);

/* misc.c:71-185 */
void orc_updatc(const orc_dims *D, double *x_temp, double *x_ip, double *xfr_temp, const double *dd,
                double *defllen_i, double *deffarea_i, double *defslen_i, const double *offset,
                const int *osflag, const double *auxpt, double *c1_i, double *c2_i, double *c3_i,
                const long *minc, const long *jcode);

/* shell.c */
void orc_stiff_sh(const orc_dims *D, double *ss, const double *emod, const double *nu,
                  const double *x_temp, const double *xlocal, const double *thick,
                  const double *farea, const double *deffarea_ip, const double *slength,
                  const double *c1_ip, const double *c2_ip, const double *c3_ip, const long *maxa,
                  const long *minc, const long *mcode);
/* returns forces_sh's code: 0, or 1 when a vertex ends beyond 1 + 10*phitol (shell.c:2044-2046) */
int  orc_forces_sh(const orc_dims *D, double *f_temp, double *ef_ip, double *ef_i, const double *dd,
                   const double *d_temp, const double *x_temp, const double *emod, const double *nu,
                   const double *xlocal, const double *thick, const double *farea,
                   const double *slength, const double *c1_ip, const double *c2_ip,
                   const double *c3_ip, const double *c1_i, const double *c2_i, const double *c3_i,
                   const long *minc, const long *mcode);
void orc_mass_sh(const orc_dims *D, double *sm, const double *dens, const double *thick, double *farea,
                 double *slength, const double *x, const long *minc, const long *mcode);
/* one element's 18x18 matrix in global axes (what stiff_sh scatters), for element-level tests */
void orc_shell_element_K(const orc_dims *D, long n, double *K18, const double *emod, const double *nu,
                         const double *x_temp, const double *xlocal, const double *thick,
                         const double *farea, const double *deffarea_ip, const double *slength,
                         const double *c1_ip, const double *c2_ip, const double *c3_ip,
                         const long *minc);

/* ANAFLAG 3: the material arrays main.c owns (yield [TR+FR+SH+BR], zstrong / zweak [FR]) and the
 * member-end flags yldflag [FR*2] that forces_fr mutates; read by the truss and frame routines */
void orc_set_plastic(const double *yield, const double *zstrong, const double *zweak, int *yldflag);
/* shells: equivalent plastic curvature chi_temp [SH*3] and vertex stress resultants efN_temp /
 * efM_temp [SH*9] (mutated by orc_forces_sh), previous-iterate coordinates, deformed area / sides */
void orc_set_plastic_sh(double *chi_temp, double *efN_temp, double *efM_temp, const double *x_ip,
                        const double *deffarea_ip, const double *defslen_ip);

/* truss.c */
void orc_stiff_tr(const orc_dims *D, double *ss, const double *emod, const double *carea,
                  const double *llength, const double *defllen_ip, const double *c1_ip,
                  const double *c2_ip, const double *c3_ip, const double *ef_ip, const long *maxa,
                  const long *mcode);
void orc_forces_tr(const orc_dims *D, double *f_temp, double *ef_i, const double *d, const double *emod,
                   const double *carea, const double *llength, const double *defllen_i,
                   const double *c1_i, const double *c2_i, const double *c3_i, const long *mcode);
void orc_mass_tr(const orc_dims *D, double *sm, const double *carea, double *llength, const double *dens,
                 const double *x, const long *minc, const long *mcode);

/* frame.c */
void orc_stiff_fr(const orc_dims *D, double *ss, const double *emod, const double *gmod,
                  const double *carea, const double *offset, const int *osflag, const double *llength,
                  const double *defllen_ip, const double *istrong, const double *iweak,
                  const double *ipolar, const double *iwarp, const double *c1_ip, const double *c2_ip,
                  const double *c3_ip, const double *ef_ip, const double *efFE_ip, const int *mendrel,
                  const long *maxa, const long *mcode);
/* returns forces_fr's code: 0, 1 (yield surface overshot, *dlpf rescaled), 2 (elastic unloading) */
int  orc_forces_fr(const orc_dims *D, double *f_temp, const double *ef_ip, double *ef_i,
                   const double *efFE_ref, const double *efFE_ip, double *efFE_i, const double *dd,
                   const double *emod, const double *gmod, const double *carea, const double *offset,
                   const int *osflag, const double *llength, const double *defllen_ip,
                   const double *istrong, const double *iweak, const double *ipolar,
                   const double *iwarp, const double *c1_ip, const double *c2_ip, const double *c3_ip,
                   const double *c1_i, const double *c2_i, const double *c3_i, const int *mendrel,
                   const long *mcode, double *dlpf, int itecnt);
void orc_mass_fr(const orc_dims *D, double *sm, const double *carea, double *llength, const double *dens,
                 const int *osflag, const double *offset, const double *x, double *xfr,
                 const long *minc, const long *mcode);

/* brick.c */
void orc_stiff_br(const orc_dims *D, double *ss, const double *x, const double *emod, const double *nu,
                  const long *minc, const long *mcode);

void orc_mass_br(const orc_dims *D, double *sm, const double *dens, const double *x, const long *minc,
                 const long *mcode);

/* solve.c:110-119: dense row-major [NEQ][NEQ] -> Ap/Ai/Ax with |a| > tol; returns nnz */
long orc_dense_to_csc(long neq, const double *ss, double tol, int *Ap, int *Ai, double *Ax);

#endif
