/* ref_capture.c - TEST INFRASTRUCTURE ONLY.
 *
 * Runs the UNMODIFIED reference driver (main.c, compiled in place from /root/reference with
 * -Dmain=ben_main) on the model_def.txt of the current directory and records, at FULL double
 * precision, every row the reference hands to output() (misc.c:345): load factor or time,
 * iteration count and the NEQ displacements of each converged increment / time step.  main.c and
 * solve.c are compiled with -Doutput=cap_output so their calls land here first; the reference's own
 * output() (misc.c, compiled as is) still writes results2.txt with its 7 digits.
 *
 * The rows go to capture.bin: int64 NEQ, int64 nrows, then nrows x (2 + NEQ) doubles.  They are the
 * golden histories for the drivers in cu-bens_b200/host (arc-length, Newmark), which results2.txt
 * (%e) could only pin to 1e-6. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern long NEQ;
void output(double *plpf, int *pitecnt, double *pd, double *pef, int flag);
int ben_main(int argc, char **argv);

static double *rows;
static long nrows, cap, width;

void cap_output(double *plpf, int *pitecnt, double *pd, double *pef, int flag)
{
    if (flag == 1) {
        if (width == 0) width = NEQ + 2;
        if (nrows == cap) {
            cap = cap ? 2 * cap : 64;
            rows = (double *)realloc(rows, (size_t)cap * width * sizeof(double));
            if (!rows) abort();
        }
        double *r = rows + nrows * width;
        r[0] = *plpf; r[1] = (double)*pitecnt;
        memcpy(r + 2, pd, (size_t)NEQ * sizeof(double));
        ++nrows;
    }
    output(plpf, pitecnt, pd, pef, flag);
}

int main(int argc, char **argv)
{
    int rc = ben_main(argc, argv);
    FILE *f = fopen("capture.bin", "wb");
    long long hdr[2] = {(long long)NEQ, (long long)nrows};
    fwrite(hdr, sizeof hdr, 1, f);
    if (nrows) fwrite(rows, sizeof(double), (size_t)(nrows * width), f);
    fclose(f);
    return rc;
}
