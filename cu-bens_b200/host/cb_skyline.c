/* cb_skyline.c - active-column LDL^T for the reference's skyline storage (host side).
 * Column j is stored from its diagonal upwards: ss[maxa[j]-1 ... maxa[j+1]-2] (0-based j),
 * i.e. entry (i <= j) at ss[maxa[j] - 1 + (j - i)]  (model.c:1269-1278).  Algorithm: Bathe's
 * COLSOL, which is what solve.c:539-698 implements; the reduction order below is the
 * textbook's, so factors agree with the reference's to rounding of identical operations. */
#include "cb_host.h"

int cb_sky_factor(long neq, const long *maxa, double *ss, double *ssd, int *det_neg,
                  int allow_indefinite)
{
    if (det_neg) *det_neg = 0;
    for (long n = 1; n <= neq; ++n) {
        const long kn = maxa[n - 1];            /* diagonal of column n (1-based address)    */
        const long kl = kn + 1, ku = maxa[n] - 1;
        const long kh = ku - kl;                /* column height above the diagonal, minus 1 */
        if (kh > 0) {
            /* reduce the off-diagonal entries of column n, top row first */
            long k = n - kh, klt = ku;
            for (long ic = 1; ic <= kh; ++ic) {
                --klt;
                const long ki = maxa[k - 1];
                const long nd = maxa[k] - ki - 1;
                if (nd > 0) {
                    const long kk = nd < ic ? nd : ic;
                    double c = 0;
                    for (long l = 1; l <= kk; ++l) c += ss[ki - 1 + l] * ss[klt - 1 + l];
                    ss[klt - 1] -= c;
                }
                ++k;
            }
        }
        if (kh >= 0) {
            /* scale by the pivots and reduce the diagonal */
            long k = n;
            double b = 0;
            for (long kk = kl; kk <= ku; ++kk) {
                --k;
                const double c = ss[kk - 1] / ss[maxa[k - 1] - 1];
                b += c * ss[kk - 1];
                ss[kk - 1] = c;
            }
            ss[kn - 1] -= b;
        }
        const double piv = ss[kn - 1];
        if (!allow_indefinite) {
            if (piv <= 0) return 1;             /* "Non-positive definite stiffness matrix"  */
        } else {
            if (ssd) ssd[n - 1] = piv;
            if (piv == 0) return 1;             /* "Singular stiffness matrix"               */
            if (piv < 0 && det_neg) *det_neg = 1;
        }
    }
    return 0;
}

void cb_sky_solve(long neq, const long *maxa, const double *ss, double *v)
{
    for (long n = 1; n <= neq; ++n) {           /* forward reduction                          */
        const long kl = maxa[n - 1] + 1, ku = maxa[n] - 1;
        if (ku - kl >= 0) {
            long k = n;
            double c = 0;
            for (long kk = kl; kk <= ku; ++kk) { --k; c += ss[kk - 1] * v[k - 1]; }
            v[n - 1] -= c;
        }
    }
    for (long n = 0; n < neq; ++n) v[n] /= ss[maxa[n] - 1];
    for (long n = neq; n >= 2; --n) {           /* back substitution                          */
        const long kl = maxa[n - 1] + 1, ku = maxa[n] - 1;
        if (ku - kl >= 0) {
            long k = n;
            for (long kk = kl; kk <= ku; ++kk) { --k; v[k - 1] -= ss[kk - 1] * v[n - 1]; }
        }
    }
}
