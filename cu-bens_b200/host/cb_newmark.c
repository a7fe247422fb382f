/* cb_newmark.c - the reference's transient drivers on the device path, through the C-ABI:
 *
 *   cb_newmark_nonlinear  main.c:3305-3960 (ALGFLAG 5): generalized-alpha Newmark with Newton
 *                         iterations inside every time step, load-factor sub-incrementation and
 *                         time-step halving.  K_t, the lumped mass and f_int are rebuilt on the
 *                         device every iteration (cb_stiff, cb_mass, cb_update_forces); the
 *                         effective matrix K + a0 (1-am)/(1-af) M and its right-hand side are the
 *                         dynamic branch of solve() (solve.c:139-186, 470-536).
 *   cb_newmark_linear     main.c:3143-3303 + solve.c:199-458 (ALGFLAG 4): one stiffness / mass
 *                         assembly, one factorisation, a time loop of effective-load solves.
 *
 * Both follow the reference statement by statement (same operation order), because the parity
 * target is its displacement history to 1e-9.  Prescribed support motion (NBC != 0) is carried by
 * cb_newmark_nonlinear_bc: heavy masses at the moved supports, matpart() (solve.c:758-824) on the
 * effective matrix, inertial reactions (main.c:3707-3727); not by the linear driver. */
#include "cb_host.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* K v for the skyline layout (skymult, solve.c:700-756): upper triangle, diagonal, lower triangle */
void cb_sky_mult(long neq, const long *maxa, const double *ss, double *v)
{
    double *out = (double *)calloc((size_t)neq, sizeof(double));
    for (long n = neq; n >= 2; --n) {
        const long kl = maxa[n - 1] + 1, ku = maxa[n] - 1;
        long k = n;
        for (long kk = kl; kk <= ku; ++kk) { --k; out[k - 1] += ss[kk - 1] * v[n - 1]; }
    }
    for (long n = 0; n < neq; ++n) out[n] += v[n] * ss[maxa[n] - 1];
    for (long n = neq; n >= 1; --n) {
        const long kl = maxa[n - 1] + 1, ku = maxa[n] - 1;
        if (ku - kl >= 0) {
            long k = n;
            double c = 0;
            for (long kk = kl; kk <= ku; ++kk) { --k; c += ss[kk - 1] * v[k - 1]; }
            out[n - 1] += c;
        }
    }
    memcpy(v, out, (size_t)neq * sizeof(double));
    free(out);
}

/* misc.c:187-250 with the dynamic out-of-balance vector in the place of qtot */
static int conv_test(long neq, const double *d_temp, const double *dd, const double *f_temp,
                     const double *fp, const double *qtot, const double *f_ip, double intener1,
                     const cb_nr_params *p, int *convchk)
{
    *convchk = 0;
    if (p->toldisp < 1) {
        double deltad = 0, totald = 0;
        for (long i = 0; i < neq; ++i) deltad += dd[i] * dd[i];
        for (long i = 0; i < neq; ++i) totald += d_temp[i] * d_temp[i];
        if (totald == 0) return 1;
        if (sqrt(deltad) / sqrt(totald) > p->toldisp) *convchk += 10;
    }
    if (p->tolforc < 1) {
        double unbfi = 0, unbfp = 0;
        for (long i = 0; i < neq; ++i) {
            unbfi += (qtot[i] - f_temp[i]) * (qtot[i] - f_temp[i]);
            unbfp += (qtot[i] - fp[i]) * (qtot[i] - fp[i]);
        }
        if (unbfp == 0) return 1;
        if (sqrt(unbfi) / sqrt(unbfp) > p->tolforc) *convchk += 100;
    }
    if (p->tolener < 1) {
        double inteneri = 0;
        for (long i = 0; i < neq; ++i) inteneri += dd[i] * (qtot[i] - f_ip[i]);
        if (intener1 == 0) return 1;
        if (fabs(inteneri / intener1) > p->tolener) *convchk += 1000;
    }
    return 0;
}

static void newmark_constants(double alpham, double alphaf, double dt_temp, double *a)
{   /* main.c:3353-3354, 3457-3464 == solve.c:171-183 */
    const double alpha = (1 - alpham + alphaf) * (1 - alpham + alphaf) / 4;
    const double delta = 0.5 - alpham + alphaf;
    a[0] = 1 / (alpha * (dt_temp * dt_temp));
    a[1] = delta / (alpha * dt_temp);
    a[2] = 1 / (alpha * dt_temp);
    a[3] = 1 / (2 * alpha) - 1;
    a[4] = delta / alpha - 1;
    a[5] = (dt_temp) / 2 * (delta / alpha - 2);
    a[6] = (dt_temp) * (1 - delta);
    a[7] = delta * (dt_temp);
}

/* matpart(), solve.c:758-824: rows / columns of the moved supports are taken out of the skyline
 * matrix (unit diagonal), their known displacements uc go to the right-hand side.  ii / ij = the
 * equations without / with support motion (main.c:1341-1353).  Statement order as the reference. */
void cb_sky_partition(long neq, long nbc, const long *maxa, double *ss, double *qtot, const double *uc,
                      const long *ii, const long *ij)
{
    long n = ij[nbc - 1] + 1;
    for (long i = 1; i <= nbc; ++i) {
        const long kl = maxa[n - 1] + 1, ku = maxa[n] - 1;
        if (ku - kl >= 0) {
            long k = n - 1;
            for (long kk = kl; kk <= ku; ++kk) {
                --k;
                for (long j = 0; j < nbc; ++j)
                    if (k == ij[j]) ss[kk - 1] = 0;
            }
            k = n - 1;
            for (long kk = kl; kk <= ku; ++kk) {
                --k;
                qtot[k] -= ss[kk - 1] * uc[n - 1];
                ss[kk - 1] = 0;
            }
        }
        if (i < nbc) n = ij[nbc - 1 - i] + 1;
    }
    for (long m = 0; m < nbc; ++m) ss[maxa[ij[m]] - 1] = 1;
    const long nf = neq - nbc;
    if (nf <= 0) return;
    n = ii[nf - 1] + 1;
    for (long i = 1; i <= nf; ++i) {
        const long kl = maxa[n - 1] + 1, ku = maxa[n] - 1;
        if (ku - kl >= 0) {
            long k = n;
            for (long kk = kl; kk <= ku; ++kk) {
                --k;
                for (long j = 0; j < nbc; ++j)
                    if (k - 1 == ij[j]) { qtot[n - 1] -= ss[kk - 1] * uc[k - 1]; ss[kk - 1] = 0; }
            }
        }
        if (i < nf) n = ii[nf - 1 - i] + 1;
    }
}

int cb_newmark_nonlinear(cb_handle *h, long neq, const long *maxa, long lss, const double *pinpt,
                         long ntstps, double dt, double alpham, double alphaf,
                         const cb_nr_params *p, double *hist, cb_nr_result *res)
{
    return cb_newmark_nonlinear_bc(h, neq, maxa, lss, pinpt, NULL, NULL, ntstps, dt, alpham, alphaf, p, hist, res);
}

int cb_newmark_nonlinear_bc(cb_handle *h, long neq, const long *maxa, long lss, const double *pinpt,
                            const double *pdisp, const int *pmot, long ntstps, double dt, double alpham,
                            double alphaf, const cb_nr_params *p, double *hist, cb_nr_result *res)
{
    if (!h || !pinpt || !p || !hist || !res) return CB_ERR_ARG;
    memset(res, 0, sizeof *res);
    cb_lin *L = NULL;                                       /* skyline, or CSC when maxa == NULL */
    if (cb_lin_create(h, neq, maxa, lss, &L) != CB_OK) return CB_ERR_ARG;
    lss = cb_lin_nval(L);
    long nbc = 0, *ii = NULL, *ij = NULL;
    if (pmot && pdisp) {
        for (long i = 0; i < neq; ++i) nbc += pmot[i] != 0;
        if (nbc) {
            ii = (long *)malloc((size_t)(neq - nbc + 1) * sizeof(long));
            ij = (long *)malloc((size_t)nbc * sizeof(long));
            long ci = 0, cj = 0;
            for (long i = 0; i < neq; ++i) { if (pmot[i] == 0) ii[ci++] = i; else ij[cj++] = i; }
        }
    }
    double *buf = (double *)calloc((size_t)neq * 20 + 2 * (size_t)lss, sizeof(double));
    if (!buf) { cb_lin_destroy(L); free(ii); free(ij); return CB_ERR_ARG; }
    double *qtot = buf, *d = qtot + neq, *d_temp = d + neq, *f = d_temp + neq, *f_temp = f + neq,
           *fp = f_temp + neq, *f_ip = fp + neq, *r = f_ip + neq, *dd = r + neq, *sm = dd + neq,
           *uc = sm + neq, *vc = uc + neq, *ac = vc + neq, *uc_i = ac + neq, *vc_i = uc_i + neq,
           *ac_i = vc_i + neq, *um = ac_i + neq, *vm = um + neq, *am = vm + neq, *dyn = am + neq,
           *ss = dyn + neq, *Keff = ss + lss;
    double lpf = p->lpf, dlpf = p->dlpf, dlpfp = dlpf, lpfi = lpf, dlpfi = dlpf, intener1 = 0;
    double ddt = 1, sub_dt = 1, dt_temp = dt, a[8];
    int tsflag = 0, solcnt = 0, subcnt = 0, convchk = 0, frcchk_fr = 0, frcchk_sh = 0, itecnt = 0;
    int status = 0, rc;
    long k = 0;
#define FAIL(code) do { status = (code); goto done; } while (0)
    do {                                                     /* time steps, main.c:3447 */
        dt_temp = ddt * dt; sub_dt = ddt; tsflag = 0;
        do {                                                 /* sub-steps, main.c:3454 */
            newmark_constants(alpham, alphaf, dt_temp, a);
            solcnt = subcnt = 0;
            lpf = lpfi; dlpf = dlpfi;
            do {                                             /* load factor, main.c:3470 */
                if (lpf > p->lpfmax) lpf = p->lpfmax;
                for (long i = 0; i < neq; ++i) { fp[i] = f[i]; d_temp[i] = d[i]; f_temp[i] = f[i]; }
                dlpfp = dlpf;
                if ((rc = cb_begin_increment(h)) != CB_OK) FAIL(100 + rc);    /* main.c:3487-3531 */
                for (long i = 0; i < neq; ++i) {
                    um[i] = uc_i[i] = uc[i]; vm[i] = vc_i[i] = vc[i]; am[i] = ac_i[i] = ac[i];
                }
                itecnt = 0;
                frcchk_fr = frcchk_sh = 0;
                do {                                         /* iterations, main.c:3544 */
                    if (nbc)                                 /* moved supports, main.c:3546-3551 */
                        for (long i = 0; i < neq; ++i)
                            if (pdisp[i * ntstps + k] != 0) uc_i[i] = (pdisp[i * ntstps + k] - um[i]) * sub_dt * lpf;
                    if (itecnt == 0) {                       /* predictor, main.c:3553-3571 */
                        if (k == 0) {
                            for (long i = 0; i < neq; ++i) {
                                qtot[i] = pinpt[i * ntstps + k] * sub_dt * lpf;
                                r[i] = qtot[i] - f_temp[i];
                            }
                        } else {
                            for (long i = 0; i < neq; ++i) {
                                const double q0 = pinpt[i * ntstps + k - 1], q1 = pinpt[i * ntstps + k];
                                qtot[i] = (q0 + (q1 - q0) * (sub_dt)) * lpf;
                                r[i] = (qtot[i] - f_temp[i]) + alphaf / (1 - alphaf) * (q0 - f_temp[i]);
                            }
                        }
                    } else {                                 /* corrector, main.c:3572-3588 */
                        for (long i = 0; i < neq; ++i) {
                            r[i] = (f_temp[i] - qtot[i]);
                            r[i] = r[i] + sm[i] * ac_i[i] -
                                   sm[i] * ((a[2] * vc_i[i] + a[3] * ac_i[i]) * (1 - alpham) - ac_i[i] * alpham) / (1 - alphaf);
                        }
                    }
                    /* ss = sm = 0; stiff_xx then mass_xx per element type (main.c:3590-3619) */
                    if ((rc = cb_stiff(h, CB_GEN_IP)) != CB_OK) FAIL(100 + rc);
                    if ((rc = cb_lin_fetch(L, h, ss)) != CB_OK) FAIL(100 + rc);
                    if ((rc = cb_mass(h)) != CB_OK) FAIL(100 + rc);
                    if ((rc = cb_get_mass(h, sm)) != CB_OK) FAIL(100 + rc);
                    ++res->stiff_calls;
                    if (cb_lin_is_scalar(L)) {
                        dd[0] = r[0] / ss[0];
                    } else {                                 /* solve(), solve.c:139-186, 470-536 */
                        if (nbc)                             /* heavy masses at the moved supports */
                            for (long i = 0; i < neq; ++i)
                                if (pdisp[i * ntstps + k] != 0) sm[i] = 1000000 * sm[i];
                        for (long i = 0; i < lss; ++i) Keff[i] = ss[i];
                        cb_lin_add_diag(L, Keff, a[0] * (1 - alpham), 1 - alphaf, sm);
                        if (!nbc && cb_lin_factor(L, Keff, NULL, NULL, 0)) FAIL(2);
                        for (long i = 0; i < neq; ++i) {
                            if (nbc && pdisp[i * ntstps + k] != 0 && itecnt > 0) dd[i] = 0;
                            else if (nbc && pdisp[i * ntstps + k] != 0 && itecnt == 0) dd[i] = uc_i[i];
                            else dd[i] = r[i] + sm[i] * ((1 - alpham) * (vc_i[i] * a[2] + ac_i[i] * a[3]) - alpham * ac_i[i]) / (1 - alphaf);
                        }
                        if (nbc) {
                            cb_lin_partition(L, Keff, dd, uc_i, pmot, nbc, ii, ij);
                            if (cb_lin_factor(L, Keff, NULL, NULL, 0)) FAIL(2);
                        }
                        cb_lin_solve(L, Keff, dd);
                    }
                    for (long i = 0; i < neq; ++i) {         /* main.c:3642-3656 */
                        if (itecnt > 0) dd[i] = dd[i] * (-1);
                        d_temp[i] += dd[i];
                        f_ip[i] = f_temp[i];
                    }
                    for (long i = 0; i < neq; ++i) {
                        uc_i[i] = d_temp[i];
                        ac_i[i] = (uc_i[i] - um[i]) * a[0] - a[2] * vm[i] - a[3] * am[i];
                        vc_i[i] = vm[i] + a[6] * am[i] + a[7] * ac_i[i];
                    }
                    if ((rc = cb_update_forces(h, dd, &dlpf, itecnt, f_temp, &frcchk_fr, &frcchk_sh)) != CB_OK)
                        FAIL(100 + rc);
                    ++res->force_calls; ++res->iterations;
                    for (long i = 0; i < neq; ++i) dyn[i] = qtot[i] - sm[i] * ac_i[i];
                    if (itecnt == 0) {
                        intener1 = 0;
                        for (long i = 0; i < neq; ++i) intener1 += dd[i] * (dyn[i] - fp[i]);
                    }
                    if (nbc)                                 /* main.c:3707-3727 */
                        for (long i = 0; i < neq; ++i)
                            if (pmot[i] != 0) f_temp[i] = -sm[i] * ac_i[i];
                    if (conv_test(neq, d_temp, dd, f_temp, fp, dyn, f_ip, intener1, p, &convchk)) FAIL(3);
                    if ((rc = cb_end_iteration(h)) != CB_OK) FAIL(100 + rc);
                    ++itecnt;
                } while (convchk != 0 && frcchk_fr == 0 && frcchk_sh == 0 && itecnt <= p->itemax);

                if (frcchk_fr == 2) {                        /* main.c:3773-3817 */
                    dlpf = dlpfp;
                } else if ((convchk != 0 || frcchk_fr != 0 || frcchk_sh != 0) && subcnt <= p->submax) {
                    if (lpf == p->lpfmax) break;
                    else if (dlpfp == p->dlpfmin) break;
                    if (frcchk_fr != 1) dlpf = dlpfp / 2;
                    if (dlpf < p->dlpfmin) dlpf = p->dlpfmin;
                    lpf = lpf - dlpfp + dlpf;
                    ++subcnt; solcnt = 0;
                } else if (subcnt > p->submax) {
                    break;
                } else {
                    ++solcnt; subcnt = 0;
                    if (solcnt >= p->solmin) { dlpf *= 2; solcnt = 0; }
                    lpf += dlpf;
                }
            } while (lpf <= p->lpfmax);

            if (convchk != 0) {                              /* main.c:3819-3839 */
                ddt = ddt / 2;
                dt_temp = ddt * dt;
                if (tsflag == 0) sub_dt = ddt;
                else if (tsflag == 1) sub_dt = sub_dt - ddt;
            } else {                                         /* commit, main.c:3841-3920 */
                for (long i = 0; i < neq; ++i) {
                    d[i] = d_temp[i]; f[i] = f_temp[i];
                    uc[i] = uc_i[i]; vc[i] = vc_i[i]; ac[i] = ac_i[i];
                }
                if ((rc = cb_commit(h)) != CB_OK) FAIL(100 + rc);
                ++res->increments;
            }
            if (frcchk_fr == 0 && frcchk_sh == 0) {          /* main.c:3922-3936 */
                if (ddt < 1) {
                    if (sub_dt <= 1) { sub_dt = sub_dt + ddt; tsflag = 1; }
                    if (sub_dt > 1) { ddt = ddt * 2; tsflag = 2; }
                } else if (ddt == 1) {
                    tsflag = 2;
                }
            }
        } while (ddt >= 0.0001 && tsflag != 2);
        if (convchk != 0 || frcchk_fr != 0 || frcchk_sh != 0) FAIL(8);   /* minimum time increment */
        hist[k * (neq + 2)] = (double)k * dt;                /* output(), main.c:3947-3950 */
        hist[k * (neq + 2) + 1] = itecnt;
        memcpy(hist + k * (neq + 2) + 2, d, (size_t)neq * sizeof(double));
        ++k;
    } while (k < ntstps);
done:
    res->status = status; res->lpf = lpf;
    free(ii); free(ij);
    free(buf);
    cb_lin_destroy(L);
    return status == 0 ? CB_OK : CB_ERR_ARG;
#undef FAIL
}

int cb_newmark_linear(cb_handle *h, long neq, const long *maxa, long lss, const double *pinpt,
                      long ntstps, double dt, double alpham, double alphaf, const double *um0,
                      const double *vm0, const double *am0, double *hist, cb_nr_result *res)
{
    if (!h || !pinpt || !hist || !res) return CB_ERR_ARG;
    memset(res, 0, sizeof *res);
    cb_lin *L = NULL;                                       /* skyline, or CSC when maxa == NULL */
    if (cb_lin_create(h, neq, maxa, lss, &L) != CB_OK) return CB_ERR_ARG;
    lss = cb_lin_nval(L);
    double *buf = (double *)calloc((size_t)neq * 10 + 2 * (size_t)lss, sizeof(double));
    if (!buf) { cb_lin_destroy(L); return CB_ERR_ARG; }
    double *sm = buf, *um = sm + neq, *vm = um + neq, *am = vm + neq, *uc = am + neq, *vc = uc + neq,
           *ac = vc + neq, *Meff = ac + neq, *Reff = Meff + neq, *dd = Reff + neq, *ss = dd + neq,
           *Keff = ss + lss;
    double a[8];
    int status = 0, rc;
#define FAIL(code) do { status = (code); goto done; } while (0)
    /* main.c:3226-3254: one stiffness + mass assembly at the initial configuration */
    if ((rc = cb_stiff(h, CB_GEN_COMMITTED)) != CB_OK) FAIL(100 + rc);
    if ((rc = cb_lin_fetch(L, h, ss)) != CB_OK) FAIL(100 + rc);
    if ((rc = cb_mass(h)) != CB_OK) FAIL(100 + rc);
    if ((rc = cb_get_mass(h, sm)) != CB_OK) FAIL(100 + rc);
    ++res->stiff_calls;
    for (long i = 0; i < neq; ++i) {
        um[i] = um0 ? um0[i] : 0.0; vm[i] = vm0 ? vm0[i] : 0.0; am[i] = am0 ? am0[i] : 0.0;
    }
    newmark_constants(alpham, alphaf, dt, a);
    for (long i = 0; i < lss; ++i) Keff[i] = ss[i];          /* solve.c:199-207 */
    cb_lin_add_diag(L, Keff, a[0] * (1 - alpham), 1 - alphaf, sm);
    if (cb_lin_factor(L, Keff, NULL, NULL, 0)) FAIL(2);
    for (long k = 0; k < ntstps; ++k) {                      /* solve.c:291-433 */
        for (long i = 0; i < neq; ++i) {
            dd[i] = um[i];
            Meff[i] = sm[i] * ((1 - alpham) * (um[i] * a[0] + vm[i] * a[2] + am[i] * a[3]) - alpham * am[i]) / (1 - alphaf);
        }
        if (k == 0) {
            for (long i = 0; i < neq; ++i) Reff[i] = pinpt[i * ntstps + k] + Meff[i];
        } else {
            for (long i = 0; i < neq; ++i)
                Reff[i] = pinpt[i * ntstps + k] + alphaf / (1 - alphaf) * pinpt[i * ntstps + k - 1] + Meff[i];
        }
        if (alphaf != 0) {
            cb_lin_mult(L, ss, dd);
            for (long i = 0; i < neq; ++i) Reff[i] -= alphaf / (1 - alphaf) * dd[i];
        }
        cb_lin_solve(L, Keff, Reff);
        for (long i = 0; i < neq; ++i) {
            uc[i] = Reff[i];
            ac[i] = a[0] * (uc[i] - um[i]) - a[2] * vm[i] - a[3] * am[i];
            vc[i] = vm[i] + a[6] * am[i] + a[7] * ac[i];
        }
        hist[k * (neq + 2)] = (double)k * dt; hist[k * (neq + 2) + 1] = 0;
        memcpy(hist + k * (neq + 2) + 2, uc, (size_t)neq * sizeof(double));
        for (long i = 0; i < neq; ++i) { um[i] = uc[i]; vm[i] = vc[i]; am[i] = ac[i]; }
        ++res->increments;
    }
done:
    res->status = status;
    free(buf);
    cb_lin_destroy(L);
    return status == 0 ? CB_OK : CB_ERR_ARG;
#undef FAIL
}
