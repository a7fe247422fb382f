/* cb_newton.c - the reference's static NR / MNR load-increment loop (main.c:1824-2152) driving
 * the device through the C-ABI: every stiff_xx / updatc / forces_xx / generation-copy block of the
 * reference is one cb_* call, everything else is the reference's
 * bookkeeping restated line for line (load-factor halving / doubling, counters, error exits). */
#include "cb_host.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* misc.c:187-250 */
static int conv_test(long neq, const double *d_temp, const double *dd, const double *f_temp,
                     const double *fp, const double *qtot, const double *f_ip, double intener1,
                     const cb_nr_params *p, int *convchk)
{
    *convchk = 0;
    if (p->toldisp < 1) {
        double deltad = 0, totald = 0;
        for (long i = 0; i < neq; ++i) deltad += dd[i] * dd[i];
        for (long i = 0; i < neq; ++i) totald += d_temp[i] * d_temp[i];
        if (totald == 0) return 1;                          /* "Displacements are zero"   */
        if (sqrt(deltad) / sqrt(totald) > p->toldisp) *convchk += 10;
    }
    if (p->tolforc < 1) {
        double unbfi = 0, unbfp = 0;
        for (long i = 0; i < neq; ++i) {
            unbfi += (qtot[i] - f_temp[i]) * (qtot[i] - f_temp[i]);
            unbfp += (qtot[i] - fp[i]) * (qtot[i] - fp[i]);
        }
        if (unbfp == 0) return 1;                           /* "Force increment is zero"  */
        if (sqrt(unbfi) / sqrt(unbfp) > p->tolforc) *convchk += 100;
    }
    if (p->tolener < 1) {
        double inteneri = 0;
        for (long i = 0; i < neq; ++i) inteneri += dd[i] * (qtot[i] - f_ip[i]);
        if (intener1 == 0) return 1;                        /* "Energy increment is zero" */
        if (fabs(inteneri / intener1) > p->tolener) *convchk += 1000;
    }
    return 0;
}

int cb_newton_static(cb_handle *h, long neq, const long *maxa, long lss, const double *q,
                     const cb_nr_params *p, double *d_out, cb_nr_result *res, double *hist,
                     int max_hist, long hist_dof)
{
    if (!h || !q || !p || !d_out || !res) return CB_ERR_ARG;
    memset(res, 0, sizeof *res);
    cb_lin *L = NULL;                                       /* skyline, or CSC when maxa == NULL */
    if (cb_lin_create(h, neq, maxa, lss, &L) != CB_OK) return CB_ERR_ARG;
    lss = cb_lin_nval(L);
    double *buf = (double *)calloc((size_t)neq * 9 + (size_t)lss, sizeof(double));
    if (!buf) { cb_lin_destroy(L); return CB_ERR_ARG; }
    double *qtot = buf, *d = qtot + neq, *d_temp = d + neq, *f = d_temp + neq, *f_temp = f + neq,
           *fp = f_temp + neq, *f_ip = fp + neq, *r = f_ip + neq, *dd = r + neq, *ss = dd + neq;
    double lpf = p->lpf, dlpf = p->dlpf, dlpfp, intener1 = 0;
    int solcnt = 0, subcnt = 0, convchk = 0, frcchk_fr = 0, frcchk_sh = 0, itecnt = 0;
    int status = 0, nh = 0, rc;
#define FAIL(code) do { status = (code); goto done; } while (0)
    do {
        if (lpf > p->lpfmax) lpf = p->lpfmax;
        for (long i = 0; i < neq; ++i) {                   /* main.c:1833-1842 */
            qtot[i] = q[i] * lpf; fp[i] = f[i]; d_temp[i] = d[i]; f_temp[i] = f[i];
        }
        dlpfp = dlpf;
        if ((rc = cb_begin_increment(h)) != CB_OK) FAIL(100 + rc);       /* main.c:1846-1882 */
        itecnt = 0;
        frcchk_fr = frcchk_sh = 0;
        do {
            for (long i = 0; i < neq; ++i) r[i] = qtot[i] - f_temp[i];  /* main.c:1893-1895 */
            const int refactor = (p->algflag == 1 || (p->algflag == 2 && itecnt == 0));
            if (refactor) {                                 /* main.c:1897-1922 */
                if ((rc = cb_stiff(h, CB_GEN_IP)) != CB_OK) FAIL(100 + rc);
                if ((rc = cb_lin_fetch(L, h, ss)) != CB_OK) FAIL(100 + rc);
                ++res->stiff_calls;
            }
            for (long i = 0; i < neq; ++i) dd[i] = r[i];    /* solve.c:71-73 */
            if (cb_lin_is_scalar(L)) {
                dd[0] = r[0] / ss[0];                       /* main.c:1925-1928 */
            } else {
                if (refactor && cb_lin_factor(L, ss, NULL, NULL, 0)) FAIL(2);
                cb_lin_solve(L, ss, dd);
            }
            for (long i = 0; i < neq; ++i) { d_temp[i] += dd[i]; f_ip[i] = f_temp[i]; }
            /* main.c:1949-1984: f_temp <- 0; updatc; forces_*; ef_ip <- ef_i */
            if ((rc = cb_update_forces(h, dd, &dlpf, itecnt, f_temp, &frcchk_fr, &frcchk_sh)) != CB_OK)
                FAIL(100 + rc);
            ++res->force_calls; ++res->iterations;
            if (itecnt == 0) {                              /* main.c:1986-1992 */
                intener1 = 0;
                for (long i = 0; i < neq; ++i) intener1 += dd[i] * (qtot[i] - fp[i]);
            }
            if (conv_test(neq, d_temp, dd, f_temp, fp, qtot, f_ip, intener1, p, &convchk)) FAIL(3);
            if ((rc = cb_end_iteration(h)) != CB_OK) FAIL(100 + rc);     /* main.c:2006-2028 */
            ++itecnt;
        } while (convchk != 0 && frcchk_fr == 0 && frcchk_sh == 0 && itecnt <= p->itemax);

        if (frcchk_fr == 2) {
            dlpf = dlpfp;
        } else if ((convchk != 0 || frcchk_fr != 0 || frcchk_sh != 0) && subcnt <= p->submax) {
            if (lpf == p->lpfmax) FAIL(4);      /* max load factor attempted without convergence */
            else if (dlpfp == p->dlpfmin) FAIL(5);
            if (frcchk_fr != 1) dlpf = dlpfp / 2;
            if (dlpf < p->dlpfmin) dlpf = p->dlpfmin;
            lpf = lpf - dlpfp + dlpf;
            ++subcnt; solcnt = 0;
        } else if (subcnt > p->submax) {
            FAIL(6);
        } else {
            ++res->increments;
            for (long i = 0; i < neq; ++i) { d[i] = d_temp[i]; f[i] = f_temp[i]; }
            if ((rc = cb_commit(h)) != CB_OK) FAIL(100 + rc);            /* main.c:2078-2134 */
            res->lpf = lpf;
            if (hist && nh < max_hist) {
                hist[nh * 3] = lpf; hist[nh * 3 + 1] = itecnt;
                hist[nh * 3 + 2] = (hist_dof >= 0 && hist_dof < neq) ? d[hist_dof] : 0.0;
                ++nh;
            }
            ++solcnt; subcnt = 0;
            if (solcnt >= p->solmin) {
                dlpf *= 2;
                if (dlpf > p->dlpfmax) dlpf = p->dlpfmax;
                solcnt = 0;
            }
            lpf += dlpf;
        }
    } while (lpf <= p->lpfmax);
    if (!(lpf >= p->lpfmax && convchk == 0)) status = 7;
done:
    memcpy(d_out, d, (size_t)neq * sizeof(double));
    res->status = status;
    free(buf);
    cb_lin_destroy(L);
    return status == 0 ? CB_OK : CB_ERR_ARG;
#undef FAIL
}
