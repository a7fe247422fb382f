/* cb_arclength.c - the reference's modified spherical arc-length (MSAL) driver, main.c:2158-3141
 * with quad() of arc.c:70-157, on the device path through the C-ABI.  Statement order follows the
 * reference (parity target: its load-factor / displacement history to 1e-9). */
#include "cb_host.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static double dotv(const double *a, const double *b, long n)      /* misc.c:252-262 */
{
    double s = 0;
    for (long i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* misc.c:187-250 */
static int conv_test(long neq, const double *d_temp, const double *dd, const double *f_temp,
                     const double *fp, const double *qtot, const double *f_ip, double intener1,
                     const cb_arc_params *p, int *convchk)
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

/* quad(), arc.c:70-157: roots of a x^2 + b x + c, the one whose displacement increment makes the
 * smaller angle with the previous increment.  0 ok, 2 imaginary roots, 3 no admissible root. */
static int quad(long neq, double a, double b, double c, double *dt, const double *dp, const double *ddr,
                const double *ddq, double *dd, double *dlpf, double *lpft, double *d1, double *d2)
{
    const double radical = b * b - 4 * a * c;
    if (!(radical > 0)) return 2;
    const double root1 = (-b + sqrt(b * b - 4 * a * c)) / (2 * a);
    double gamma1 = 0, gamma2 = 0;
    for (long i = 0; i < neq; ++i) {
        d1[i] = dt[i] - dp[i] + ddr[i] + root1 * ddq[i];
        gamma1 += (dt[i] - dp[i]) * d1[i];
    }
    const double root2 = (-b - sqrt(b * b - 4 * a * c)) / (2 * a);
    for (long i = 0; i < neq; ++i) {
        d2[i] = dt[i] - dp[i] + ddr[i] + root2 * ddq[i];
        gamma2 += (dt[i] - dp[i]) * d2[i];
    }
    if (gamma1 > gamma2 && gamma1 > 0) {
        for (long i = 0; i < neq; ++i) { dt[i] = d1[i] + dp[i]; dd[i] = ddr[i] + root1 * ddq[i]; }
        *dlpf = root1; *lpft += *dlpf;
    } else if (gamma2 > gamma1 && gamma2 > 0) {
        for (long i = 0; i < neq; ++i) { dt[i] = d2[i] + dp[i]; dd[i] = ddr[i] + root2 * ddq[i]; }
        *dlpf = root2; *lpft += *dlpf;
    } else return 3;
    return 0;
}

/* solve() for ALGFLAG 3 / SLVFLAG 0 (solve.c:71-82): dd <- r; skyfact (fact == 0) + skysolve.
 * skyfact resets *det on every call, also when it does not factorise (solve.c:545). */
static int sky_solve3(cb_lin *L, long neq, double *ss, double *ssd, const double *r, double *out,
                      int fact, int *det)
{
    for (long i = 0; i < neq; ++i) out[i] = r[i];
    *det = 0;
    if (fact == 0 && cb_lin_factor(L, ss, ssd, det, 1)) return 1;
    cb_lin_solve(L, ss, out);
    return 0;
}

int cb_arclength_static(cb_handle *h, long neq, const long *maxa, long lss, const double *q,
                        const cb_arc_params *p, double *hist, int max_rows, cb_nr_result *res)
{
    if (!h || !q || !p || !hist || !res) return CB_ERR_ARG;
    memset(res, 0, sizeof *res);
    cb_lin *L = NULL;                                       /* skyline, or CSC when maxa == NULL */
    if (cb_lin_create(h, neq, maxa, lss, &L) != CB_OK) return CB_ERR_ARG;
    lss = cb_lin_nval(L);
    const int scalar = cb_lin_is_scalar(L);
    double *buf = (double *)calloc((size_t)neq * 17 + (size_t)lss, sizeof(double));
    if (!buf) { cb_lin_destroy(L); return CB_ERR_ARG; }
    double *d = buf, *dp = d + neq, *dpp = dp + neq, *f = dpp + neq, *fp = f + neq, *dd = fp + neq,
           *ddq = dd + neq, *ddr = ddq + neq, *ssd_o = ddr + neq, *ssd = ssd_o + neq, *qtot = ssd + neq,
           *r = qtot + neq, *f_ip = r + neq, *d_temp = f_ip + neq, *f_temp = d_temp + neq,
           *w1 = f_temp + neq, *w2 = w1 + neq, *ss = w2 + neq;
    const long k = p->dkdof;
    double lpf, dlpf, lpfp = 0, lpfpp, lpf_temp, intener1 = 0, arc = 0, beta, psi, a = 0, b, c, dnorm,
           dnormallow, dkc, lpfc, temp;
    int det = 0, itecnt = 0, convchk = 0, frcchk_fr = 0, frcchk_sh = 0, errchk2, subcnt, imagcnt, negcnt;
    int status = 0, rc, nrow = 0;
#define FAIL(code) do { status = (code); goto done; } while (0)
#define ROW() do { if (nrow < max_rows) { hist[(long)nrow * (neq + 2)] = lpf; hist[(long)nrow * (neq + 2) + 1] = itecnt; \
                   memcpy(hist + (long)nrow * (neq + 2) + 2, d, (size_t)neq * sizeof(double)); ++nrow; } } while (0)

    /* ---- first increment: prescribed displacement dk at DOF k (main.c:2297-2560) ------------- */
    if ((rc = cb_begin_increment(h)) != CB_OK) FAIL(100 + rc);
    if ((rc = cb_stiff(h, CB_GEN_COMMITTED)) != CB_OK) FAIL(100 + rc);
    if ((rc = cb_lin_fetch(L, h, ss)) != CB_OK) FAIL(100 + rc);
    ++res->stiff_calls;
    if (scalar) { ddq[0] = q[0] / ss[0]; ssd[0] = ss[0]; }
    else if (sky_solve3(L, neq, ss, ssd, q, ddq, 0, &det)) FAIL(2);
    lpf = p->dk / ddq[k];
    for (long i = 0; i < neq; ++i) {
        d[i] = dd[i] = lpf * ddq[i];
        intener1 += dd[i] * (lpf * q[i]);
        ssd_o[i] = ssd[i];
    }
    itecnt = 0;
    if ((rc = cb_update_forces(h, dd, &lpf, itecnt, f, &frcchk_fr, &frcchk_sh)) != CB_OK) FAIL(100 + rc);
    ++res->force_calls;
    itecnt = 1;
    if ((rc = cb_end_iteration(h)) != CB_OK) FAIL(100 + rc);
    frcchk_fr = frcchk_sh = 0;
    do {
        for (long i = 0; i < neq; ++i) { qtot[i] = q[i] * lpf; r[i] = qtot[i] - f[i]; }
        if (scalar) { ddr[0] = r[0] / ss[0]; ssd[0] = ss[0]; }
        else if (sky_solve3(L, neq, ss, ssd, r, ddr, 1, &det)) FAIL(2);
        dlpf = -ddr[k] / ddq[k];
        lpf += dlpf;
        for (long i = 0; i < neq; ++i) { dd[i] = ddr[i] + dlpf * ddq[i]; d[i] += dd[i]; f_ip[i] = f[i]; }
        if ((rc = cb_update_forces(h, dd, &dlpf, itecnt, f, &frcchk_fr, &frcchk_sh)) != CB_OK) FAIL(100 + rc);
        ++res->force_calls; ++res->iterations;
        if (conv_test(neq, d, dd, f, fp, qtot, f_ip, intener1, p, &convchk)) FAIL(3);
        ++itecnt;
        if ((rc = cb_end_iteration(h)) != CB_OK) FAIL(100 + rc);
    } while (convchk != 0 && frcchk_fr == 0 && frcchk_sh == 0 && itecnt <= p->itemax);
    if (convchk != 0 || frcchk_fr != 0 || frcchk_sh != 0) FAIL(9);    /* initial displacement too large */
    if ((rc = cb_commit(h)) != CB_OK) FAIL(100 + rc);
    ROW();
    dkc = fabs(d[k]); lpfc = fabs(lpf);

    /* ---- arc-length controlled increments (main.c:2562-3134) ------------------------------- */
    for (long i = 0; i < neq; ++i) { dpp[i] = dp[i]; dp[i] = d[i]; fp[i] = f[i]; }
    lpfpp = lpfp; lpfp = lpf;
    if ((rc = cb_begin_increment(h)) != CB_OK) FAIL(100 + rc);
    dnorm = sqrt(dotv(dp, dp, neq));
    dnormallow = p->alpha * sqrt(dotv(dp, dp, neq));
    beta = sqrt(((double)p->iteopt) / (double)itecnt) * (dnormallow / dnorm);
    psi = 1;
    for (long i = 0; i < neq; ++i) { temp = fabs(ssd[i] / ssd_o[i]); if (temp < psi) psi = temp; }
    errchk2 = subcnt = imagcnt = negcnt = 0;
    while (lpfc <= p->lpfmax && dkc <= p->dkimax) {
        for (long i = 0; i < neq; ++i) d_temp[i] = d[i];
        lpf_temp = lpf;
        if (errchk2 == 0) {
            double dotprod = 0;
            for (long i = 0; i < neq; ++i) dotprod += (dp[i] - dpp[i]) * (dp[i] - dpp[i]);
            if (psi >= p->psi_thresh) arc = beta * sqrt(dotprod + (lpfp - lpfpp) * (lpfp - lpfpp));
            else arc = beta * sqrt(dotprod);
        } else {
            if ((rc = cb_begin_increment(h)) != CB_OK) FAIL(100 + rc);      /* main.c:2667-2707 */
            errchk2 = 0;
        }
        if ((rc = cb_stiff(h, CB_GEN_COMMITTED)) != CB_OK) FAIL(100 + rc);
        if ((rc = cb_lin_fetch(L, h, ss)) != CB_OK) FAIL(100 + rc);
        ++res->stiff_calls;
        if (scalar) { ddq[0] = q[0] / ss[0]; ssd[0] = ss[0]; det = (ss[0] > 0) ? 0 : 1; }
        else if (sky_solve3(L, neq, ss, ssd, q, ddq, 0, &det)) FAIL(2);
        if (psi >= p->psi_thresh) a = dotv(q, q, neq) + dotv(ddq, ddq, neq);
        else a = dotv(ddq, ddq, neq);
        if (det == 0) dlpf = arc * sqrt(1 / a);
        else dlpf = -arc * sqrt(1 / a);
        lpf_temp += dlpf;
        intener1 = 0;
        for (long i = 0; i < neq; ++i) {
            dd[i] = dlpf * ddq[i];
            d_temp[i] += dd[i];
            intener1 += dd[i] * (dlpf * q[i]);
        }
        {   /* predictor: the reference ignores the return codes of these calls (main.c:2785-2808) */
            int ffr = 0, fsh = 0;
            if ((rc = cb_update_forces(h, dd, &dlpf, itecnt, f_temp, &ffr, &fsh)) != CB_OK) FAIL(100 + rc);
            ++res->force_calls;
        }
        if ((rc = cb_end_iteration(h)) != CB_OK) FAIL(100 + rc);
        itecnt = 1;
        frcchk_fr = frcchk_sh = 0;
        do {
            for (long i = 0; i < neq; ++i) {
                qtot[i] = q[i] * lpf_temp;
                r[i] = qtot[i] - f_temp[i];
                f_ip[i] = f_temp[i];
            }
            if (scalar) ddr[0] = r[0] / ss[0];
            else if (sky_solve3(L, neq, ss, ssd, r, ddr, 1, &det)) FAIL(2);
            if (psi >= p->psi_thresh) {
                b = 2 * (dotv(d_temp, ddq, neq) - dotv(dp, ddq, neq) + dotv(ddr, ddq, neq) +
                         (lpf_temp - lpfp) * dotv(q, q, neq));
                c = 2 * (dotv(d_temp, ddr, neq) - dotv(dp, ddr, neq) - dotv(d_temp, dp, neq)) +
                    dotv(d_temp, d_temp, neq) + dotv(dp, dp, neq) + dotv(ddr, ddr, neq) +
                    (lpf_temp - lpfp) * (lpf_temp - lpfp) * dotv(q, q, neq) - arc * arc;
            } else {
                b = 2 * (dotv(d_temp, ddq, neq) - dotv(dp, ddq, neq) + dotv(ddr, ddq, neq));
                c = 2 * (dotv(d_temp, ddr, neq) - dotv(dp, ddr, neq) - dotv(d_temp, dp, neq)) +
                    dotv(d_temp, d_temp, neq) + dotv(dp, dp, neq) + dotv(ddr, ddr, neq) - arc * arc;
            }
            errchk2 = quad(neq, a, b, c, d_temp, dp, ddr, ddq, dd, &dlpf, &lpf_temp, w1, w2);
            if (errchk2 == 0) {
                if ((rc = cb_update_forces(h, dd, &dlpf, itecnt, f_temp, &frcchk_fr, &frcchk_sh)) != CB_OK)
                    FAIL(100 + rc);
                ++res->force_calls; ++res->iterations;
                if (conv_test(neq, d_temp, dd, f_temp, fp, qtot, f_ip, intener1, p, &convchk)) FAIL(3);
                if (convchk == 0) {
                    dnorm = 0;
                    for (long i = 0; i < neq; ++i) dnorm += (d_temp[i] - dp[i]) * (d_temp[i] - dp[i]);
                    dnorm = sqrt(dnorm);
                    if (dnorm > 100 * dnormallow) {
                        arc /= beta; beta = dnormallow / dnorm; arc *= beta;
                        for (long i = 0; i < neq; ++i) f_temp[i] = f[i];
                        errchk2 = 1;
                    }
                } else if (frcchk_fr == 2) {
                    for (long i = 0; i < neq; ++i) f_temp[i] = f[i];
                    errchk2 = 1;
                } else if ((frcchk_fr != 0 || frcchk_sh != 0) && subcnt <= p->submax) {
                    arc *= 0.5;
                    for (long i = 0; i < neq; ++i) f_temp[i] = f[i];
                    ++subcnt; errchk2 = 1;
                } else {
                    ++itecnt;
                    if (itecnt > p->itemax) {
                        arc *= 0.5;
                        for (long i = 0; i < neq; ++i) f_temp[i] = f[i];
                        ++subcnt; errchk2 = 1;
                    }
                    if ((rc = cb_end_iteration(h)) != CB_OK) FAIL(100 + rc);
                }
            } else if (errchk2 == 2) {
                arc *= 0.5;
                for (long i = 0; i < neq; ++i) f_temp[i] = f[i];
                ++imagcnt;
            } else if (errchk2 == 3) {
                arc *= 0.5;
                for (long i = 0; i < neq; ++i) f_temp[i] = f[i];
                ++negcnt;
            }
        } while (convchk != 0 && errchk2 == 0 && subcnt <= p->submax && imagcnt <= p->imagmax &&
                 negcnt <= p->negmax);
        if (subcnt > p->submax || imagcnt > p->imagmax || negcnt > p->negmax) FAIL(6);
        else if (errchk2 == 0) {
            beta = sqrt(((double)p->iteopt) / ((double)itecnt)) * (dnormallow / dnorm);
            for (long i = 0; i < neq; ++i) {
                dpp[i] = dp[i];
                d[i] = dp[i] = d_temp[i];
                f[i] = fp[i] = f_temp[i];
            }
            lpfpp = lpfp;
            lpf = lpfp = lpf_temp;
            if ((rc = cb_commit(h)) != CB_OK) FAIL(100 + rc);
            /* the reference leaves *_ip at the second-to-last iterate here (main.c:2925-2940) */
            if ((rc = cb_keep_ip(h)) != CB_OK) FAIL(100 + rc);
            dkc = fabs(d[k]); lpfc = fabs(lpf);
            psi = 1;
            for (long i = 0; i < neq; ++i) { temp = fabs(ssd[i] / ssd_o[i]); if (temp < psi) psi = temp; }
            subcnt = imagcnt = negcnt = 0;
            ROW();
        }
    }
done:
    res->status = status; res->increments = nrow; res->lpf = (nrow > 0) ? hist[(long)(nrow - 1) * (neq + 2)] : 0;
    free(buf);
    cb_lin_destroy(L);
    return status == 0 ? CB_OK : CB_ERR_ARG;
#undef FAIL
#undef ROW
}
