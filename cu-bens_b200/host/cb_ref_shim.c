/* cb_ref_shim.c - the reference's OWN driver on the C-ABI, without editing it.
 *
 * This translation unit exports the hot-path routines of the reference under their own names and
 * signatures (prototypes.h:91-93 stiff_tr, :100-102 forces_tr, :106-107 mass_tr, :122-127 stiff_fr, :149-155
 * forces_fr, :159-161 mass_fr, :187-191 stiff_sh, :240-242 mass_sh, :246-251 forces_sh) and forwards them to
 * libcubens_b200 (include/cubens_b200.h).  Linked with the UNMODIFIED main.c / model.c / solve.c / arc.c /
 * memory.c / misc.c (the recipe that compiles the reference in place: ben_b200.exe, ben_b200_capture.exe) it turns the reference program
 * itself into a host of the device path: input decks, Newton / arc-length / Newmark loops, step halving,
 * output files - all the reference's, every element stiffness, internal force and mass evaluated on the GPU.
 *
 * How it hooks in: the reference's definitions of these routines are renamed when their files are compiled
 * (-Dstiff_sh=ref_stiff_sh ..., no source is touched), so main.c's calls land here.  The set-up routines
 * prop_tr / prop_fr / prop_sh, codes, skylin and load are wrapped the same way: the wrapper calls the
 * reference's own routine and remembers the arrays main() handed it - that is how the shim learns about
 * arrays (jcode, dens, auxpt ...) that are locals of main() and never reach a hot routine.
 *
 * main.c owns the updated-Lagrangian state (x_temp, c1_i/_ip, ef_i/_ip, deffarea ... three generations copied
 * around by loops inside main()), so every call here is STATELESS: the arrays the call receives are
 * uploaded, the device evaluates, the results the reference routine would have written are downloaded.  One
 * handle per element type keeps the calls independent of each other exactly as the reference's are
 * (stiff_tr, stiff_fr, stiff_sh each ADD their elements into ss).  That costs PCIe round trips per call -
 * this file is the drop-in PROOF (same decks, same driver, histories equal to 1e-9), the production
 * integration keeps the state resident (cu-bens_b200/host/cb_newton.c ...).
 *
 * updatc (misc.c:71) is left to the reference: it only refreshes main()'s own copies of the coordinates and
 * triads; the device forms the same quantities inside cb_update_forces (fused), from the pre-update
 * coordinates and dd that the shim remembers from the updatc call.  Bricks / FSI stay on the reference path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/cubens_b200.h"

/* globals of main.c:323-328 */
extern long NJ, NE_TR, NE_FR, NE_SH, NE_SBR, NE_FBR, NEQ;
extern int ANAFLAG, ALGFLAG, SLVFLAG;

/* the reference's own routines under their compile-time aliases */
void ref_prop_tr(double *, double *, double *, double *, double *, double *, double *, double *, double *, long *);
void ref_prop_fr(double *, double *, double *, double *, double *, double *, int *, double *, double *, double *,
                 double *, double *, double *, double *, double *, double *, double *, double *, double *, double *,
                 int *, long *);
void ref_prop_sh(double *, double *, double *, double *, double *, double *, double *, double *, double *, double *,
                 double *, double *, long *);
void ref_codes(long *, long *, long *, int *);
int ref_skylin(long *, long *, long *, long *, long *, long *);
int ref_load(double *, double *, double *, double *, double *, int *, double *, double *, double *, long *, long *,
             long *, long *, double *, double *, double *, double *, double *, double *, double *);
void ref_updatc(double *, double *, double *, double *, double *, double *, double *, double *, int *, double *,
                double *, double *, double *, long *, long *);

/* ---- what main() showed the set-up routines ------------------------------------------------------- */
static struct {
    double *x, *xfr, *emod, *gmod, *dens, *offset, *auxpt, *carea, *llength, *istrong, *iweak, *ipolar, *iwarp,
        *yield, *zstrong, *zweak, *c1, *c2, *c3, *nu, *xlocal, *thick, *farea, *slength, *efFE_ref;
    int *osflag, *mendrel;
    long *minc, *mcode, *jcode, *maxa;
    /* last updatc call: coordinates before the update (its x_ip output) and the increment */
    double *upd_x_ip, *upd_dd;
} S;

static cb_handle *H[3];                 /* truss, frame, shell */
static double *scratch; static long nscratch;

static double *buf(long n)
{
    if (n > nscratch) { free(scratch); scratch = (double *)malloc((size_t)n * sizeof(double)); nscratch = n; }
    if (!scratch) { fprintf(stderr, "cb_ref_shim: out of memory\n"); exit(3); }
    return scratch;
}
static void die(const char *what)
{
    fprintf(stderr, "cb_ref_shim: %s failed: %s\n", what, cb_last_error());
    exit(3);
}
#define CK(call) do { if ((call) != CB_OK) die(#call); } while (0)

void prop_tr(double *px, double *pemod, double *pcarea, double *pdens, double *pllength, double *pyield, double *pc1,
             double *pc2, double *pc3, long *pminc)
{
    ref_prop_tr(px, pemod, pcarea, pdens, pllength, pyield, pc1, pc2, pc3, pminc);
    S.x = px; S.emod = pemod; S.carea = pcarea; S.dens = pdens; S.llength = pllength; S.yield = pyield;
    S.c1 = pc1; S.c2 = pc2; S.c3 = pc3; S.minc = pminc;
}
void prop_fr(double *px, double *pxfr, double *pemod, double *pgmod, double *pdens, double *poffset, int *posflag,
             double *pauxpt, double *pcarea, double *pllength, double *pistrong, double *piweak, double *pipolar,
             double *piwarp, double *pyield, double *pzstrong, double *pzweak, double *pc1, double *pc2, double *pc3,
             int *pmendrel, long *pminc)
{
    ref_prop_fr(px, pxfr, pemod, pgmod, pdens, poffset, posflag, pauxpt, pcarea, pllength, pistrong, piweak, pipolar,
                piwarp, pyield, pzstrong, pzweak, pc1, pc2, pc3, pmendrel, pminc);
    S.x = px; S.xfr = pxfr; S.emod = pemod; S.gmod = pgmod; S.dens = pdens; S.offset = poffset; S.osflag = posflag;
    S.auxpt = pauxpt; S.carea = pcarea; S.llength = pllength; S.istrong = pistrong; S.iweak = piweak;
    S.ipolar = pipolar; S.iwarp = piwarp; S.yield = pyield; S.zstrong = pzstrong; S.zweak = pzweak;
    S.c1 = pc1; S.c2 = pc2; S.c3 = pc3; S.mendrel = pmendrel; S.minc = pminc;
}
void prop_sh(double *px, double *pemod, double *pnu, double *pxlocal, double *pthick, double *pdens, double *pfarea,
             double *pslength, double *pyield, double *pc1, double *pc2, double *pc3, long *pminc)
{
    ref_prop_sh(px, pemod, pnu, pxlocal, pthick, pdens, pfarea, pslength, pyield, pc1, pc2, pc3, pminc);
    S.x = px; S.emod = pemod; S.nu = pnu; S.xlocal = pxlocal; S.thick = pthick; S.dens = pdens; S.farea = pfarea;
    S.slength = pslength; S.yield = pyield; S.c1 = pc1; S.c2 = pc2; S.c3 = pc3; S.minc = pminc;
}
void codes(long *pmcode, long *pjcode, long *pminc, int *pwrpres)
{
    ref_codes(pmcode, pjcode, pminc, pwrpres);
    S.mcode = pmcode; S.jcode = pjcode; S.minc = pminc;
}
int skylin(long *pmaxa, long *pmcode, long *plss, long *pjcode, long *pkht, long *ppmot)
{
    const int rc = ref_skylin(pmaxa, pmcode, plss, pjcode, pkht, ppmot);
    S.maxa = pmaxa; S.mcode = pmcode; S.jcode = pjcode;
    return rc;
}
int load(double *pq, double *pefFE_ref, double *px, double *pllength, double *poffset, int *posflag, double *pc1,
         double *pc2, double *pc3, long *pjnt, long *pmcode, long *pjcode, long *pminc, double *ptinpt, double *ppinpt,
         double *pdinpt, double *ppdisp, double *pum, double *pvm, double *pam)
{
    S.efFE_ref = pefFE_ref;
    return ref_load(pq, pefFE_ref, px, pllength, poffset, posflag, pc1, pc2, pc3, pjnt, pmcode, pjcode, pminc, ptinpt,
                    ppinpt, pdinpt, ppdisp, pum, pvm, pam);
}
void updatc(double *px_temp, double *px_ip, double *pxfr_temp, double *pdd, double *pdefllen_i, double *pdeffarea_i,
            double *pdefslen_i, double *poffset, int *posflag, double *pauxpt, double *pc1_i, double *pc2_i,
            double *pc3_i, long *pminc, long *pjcode)
{
    ref_updatc(px_temp, px_ip, pxfr_temp, pdd, pdefllen_i, pdeffarea_i, pdefslen_i, poffset, posflag, pauxpt, pc1_i,
               pc2_i, pc3_i, pminc, pjcode);
    S.upd_x_ip = px_ip; S.upd_dd = pdd;      /* x_ip now holds the coordinates the update started from */
}

/* ---- one handle per element type: a sub-model with that type's elements only ------------------------- */
enum { T_TR = 0, T_FR = 1, T_SH = 2 };
static cb_handle *handle(int t)
{
    if (H[t]) return H[t];
    if (!S.jcode || !S.mcode || !S.minc || !S.x) { fprintf(stderr, "cb_ref_shim: hot routine called before set-up\n"); exit(3); }
    if (SLVFLAG != 0) { fprintf(stderr, "cb_ref_shim: the proof runs the skyline path (SLVFLAG 0)\n"); exit(3); }
    cb_sizes sz; memset(&sz, 0, sizeof sz);
    sz.NJ = NJ; sz.NEQ = NEQ;
    cb_flags fl; memset(&fl, 0, sizeof fl);
    fl.ANAFLAG = ANAFLAG; fl.ALGFLAG = ALGFLAG; fl.SLVFLAG = 0; fl.matrix_layout = CB_MAT_SKYLINE; fl.device = 0;
    cb_model m; memset(&m, 0, sizeof m);
    m.x = S.x; m.jcode = S.jcode; m.maxa = S.maxa; m.dens = S.dens;   /* dens[n] is indexed per type (App. B.4) */
    const long TR = NE_TR, FR = NE_FR;
    if (t == T_TR) {
        sz.NE_TR = NE_TR;
        m.minc = S.minc; m.mcode = S.mcode; m.emod = S.emod; m.yield = S.yield; m.carea = S.carea; m.llength = S.llength;
        m.c1 = S.c1; m.c2 = S.c2; m.c3 = S.c3;
    } else if (t == T_FR) {
        sz.NE_FR = NE_FR;
        m.minc = S.minc + 2 * TR; m.mcode = S.mcode + 6 * TR; m.emod = S.emod + TR; m.yield = S.yield + TR;
        m.carea = S.carea + TR; m.llength = S.llength + TR; m.c1 = S.c1 + TR; m.c2 = S.c2 + TR; m.c3 = S.c3 + TR;
        m.gmod = S.gmod; m.istrong = S.istrong; m.iweak = S.iweak; m.ipolar = S.ipolar; m.iwarp = S.iwarp;
        m.auxpt = S.auxpt; m.offset = S.offset; m.osflag = S.osflag; m.mendrel = S.mendrel; m.efFE_ref = S.efFE_ref;
        m.zstrong = S.zstrong; m.zweak = S.zweak;
    } else {
        sz.NE_SH = NE_SH;
        m.minc = S.minc + 2 * TR + 2 * FR; m.mcode = S.mcode + 6 * TR + 14 * FR; m.emod = S.emod + TR + FR;
        m.yield = S.yield + TR + FR; m.c1 = S.c1 + TR + 3 * FR; m.c2 = S.c2 + TR + 3 * FR; m.c3 = S.c3 + TR + 3 * FR;
        m.nu = S.nu; m.thick = S.thick; m.farea = S.farea; m.slength = S.slength; m.xlocal = S.xlocal;
    }
    if (cb_create(&sz, &fl, &m, &H[t]) != CB_OK) die("cb_create");
    return H[t];
}

static void up(cb_handle *h, int which, const double *src, long n) { if (cb_upload(h, which, src, n) != CB_OK) die("cb_upload"); }
static void down(cb_handle *h, int which, double *dst, long n) { if (cb_download(h, which, dst, n) != CB_OK) die("cb_download"); }

/* ss += the skyline the device assembled for this element type */
static void add_skyline(cb_handle *h, double *pss)
{
    const long lss = S.maxa[NEQ] - 1;
    double *b = buf(lss);
    CK(cb_stiff(h, CB_GEN_IP));
    CK(cb_get_skyline(h, b, lss));
    for (long i = 0; i < lss; ++i) pss[i] += b[i];
    CK(cb_end_iteration(h));
}
/* updatc + forces_* on the device: f_temp += f, returns through the out-pointers what the routine returns */
static void run_forces(cb_handle *h, const double *pdd, double *pf_temp, double *pdlpf, int itecnt, int *fr, int *sh)
{
    double *f = buf(NEQ);
    if (ANAFLAG == 1) {
        CK(cb_forces_linear(h, pdd, f));            /* main.c:1774-1793, 3278-3297: total displacement in */
        *fr = *sh = 0;
    } else {
        double one = 1.0;
        CK(cb_update_forces(h, pdd, pdlpf ? pdlpf : &one, itecnt, f, fr, sh));
    }
    for (long i = 0; i < NEQ; ++i) pf_temp[i] += f[i];
}

/* ---- trusses ------------------------------------------------------------------------------------------ */
void stiff_tr(double *pss, double *pemod, double *pcarea, double *plength, double *pdefllen_ip, double *pyield,
              double *pc1_ip, double *pc2_ip, double *pc3_ip, double *pef_ip, long *pmaxa, long *pmcode)
{
    (void)pemod; (void)pcarea; (void)pyield; (void)pmaxa; (void)pmcode;
    cb_handle *h = handle(T_TR);
    up(h, CB_ARR_LLENGTH, plength, NE_TR); up(h, CB_ARR_DEFLLEN_IP, pdefllen_ip, NE_TR);
    up(h, CB_ARR_C1_IP, pc1_ip, NE_TR); up(h, CB_ARR_C2_IP, pc2_ip, NE_TR); up(h, CB_ARR_C3_IP, pc3_ip, NE_TR);
    up(h, CB_ARR_EF_IP, pef_ip, 2 * NE_TR);
    add_skyline(h, pss);
}
void forces_tr(double *pf_temp, double *pef_i, double *pd, double *pemod, double *pcarea, double *pllength,
               double *pdefllen_i, double *pyield, double *pc1_i, double *pc2_i, double *pc3_i, long *pmcode)
{
    (void)pemod; (void)pcarea; (void)pdefllen_i; (void)pyield; (void)pc1_i; (void)pc2_i; (void)pc3_i; (void)pmcode;
    cb_handle *h = handle(T_TR);
    int fr, sh;
    up(h, CB_ARR_LLENGTH, pllength, NE_TR);
    if (ANAFLAG == 1) {
        run_forces(h, pd, pf_temp, NULL, 0, &fr, &sh);
        down(h, CB_ARR_EF, pef_i, 2 * NE_TR);
    } else {
        up(h, CB_ARR_X_TEMP, S.upd_x_ip, 3 * NJ);
        run_forces(h, S.upd_dd, pf_temp, NULL, 0, &fr, &sh);
        down(h, CB_ARR_EF_I, pef_i, 2 * NE_TR);
        CK(cb_end_iteration(h));
    }
}
void mass_tr(double *psm, double *pcarea, double *pllength, double *pdens, double *px, long *pminc, long *pmcode,
             double *pjac)
{
    (void)pcarea; (void)pdens; (void)pminc; (void)pmcode; (void)pjac;
    cb_handle *h = handle(T_TR);
    double *b = buf(NEQ);
    up(h, CB_ARR_X, px, 3 * NJ);
    CK(cb_mass(h)); CK(cb_get_mass(h, b));
    for (long i = 0; i < NEQ; ++i) psm[i] += b[i];
    down(h, CB_ARR_LLENGTH, pllength, NE_TR);              /* mass_tr rewrites llength (truss.c:396) */
}

/* ---- frames ------------------------------------------------------------------------------------------- */
static void frame_state_up(cb_handle *h, double *pllength, double *pdefllen_ip, double *pc1_ip, double *pc2_ip,
                           double *pc3_ip, double *pef_ip, double *pefFE_ip, int *pyldflag)
{
    const long TR = NE_TR, FR = NE_FR;
    up(h, CB_ARR_LLENGTH, pllength + TR, FR); up(h, CB_ARR_DEFLLEN_IP, pdefllen_ip + TR, FR);
    up(h, CB_ARR_C1_IP, pc1_ip + TR, 3 * FR); up(h, CB_ARR_C2_IP, pc2_ip + TR, 3 * FR); up(h, CB_ARR_C3_IP, pc3_ip + TR, 3 * FR);
    up(h, CB_ARR_EF_IP, pef_ip + 2 * TR, 14 * FR); up(h, CB_ARR_EFFE_IP, pefFE_ip, 14 * FR);
    if (ANAFLAG == 3) CK(cb_set_yldflag(h, pyldflag, 2 * FR));
}
void stiff_fr(double *pss, double *pemod, double *pgmod, double *pcarea, double *poffset, int *posflag,
              double *pllength, double *pdefllen_ip, double *pistrong, double *piweak, double *pipolar, double *piwarp,
              int *pyldflag, double *pyield, double *pzstrong, double *pzweak, double *pc1_ip, double *pc2_ip,
              double *pc3_ip, double *pef_ip, double *pefFE_ip, int *pmendrel, long *pmaxa, long *pmcode)
{
    (void)pemod; (void)pgmod; (void)pcarea; (void)poffset; (void)posflag; (void)pistrong; (void)piweak; (void)pipolar;
    (void)piwarp; (void)pyield; (void)pzstrong; (void)pzweak; (void)pmendrel; (void)pmaxa; (void)pmcode;
    cb_handle *h = handle(T_FR);
    frame_state_up(h, pllength, pdefllen_ip, pc1_ip, pc2_ip, pc3_ip, pef_ip, pefFE_ip, pyldflag);
    add_skyline(h, pss);
}
int forces_fr(double *pf_temp, double *pef_ip, double *pef_i, double *pefFE_ref, double *pefFE_ip, double *pefFE_i,
              int *pyldflag, double *pdd, double *pemod, double *pgmod, double *pcarea, double *poffset, int *posflag,
              double *pllength, double *pdefllen_ip, double *pistrong, double *piweak, double *pipolar, double *piwarp,
              double *pyield, double *pzstrong, double *pzweak, double *pc1_ip, double *pc2_ip, double *pc3_ip,
              double *pc1_i, double *pc2_i, double *pc3_i, int *pmendrel, long *pmcode, double *pdlpf, int *pitecnt)
{
    (void)pefFE_ref; (void)pemod; (void)pgmod; (void)pcarea; (void)poffset; (void)posflag; (void)pistrong; (void)piweak;
    (void)pipolar; (void)piwarp; (void)pyield; (void)pzstrong; (void)pzweak; (void)pc1_i; (void)pc2_i; (void)pc3_i;
    (void)pmendrel; (void)pmcode;
    cb_handle *h = handle(T_FR);
    const long TR = NE_TR, FR = NE_FR;
    int fr = 0, sh = 0;
    frame_state_up(h, pllength, pdefllen_ip, pc1_ip, pc2_ip, pc3_ip, pef_ip, pefFE_ip, pyldflag);
    if (ANAFLAG == 1) {
        run_forces(h, pdd, pf_temp, NULL, 0, &fr, &sh);
        down(h, CB_ARR_EF, pef_i + 2 * TR, 14 * FR);
        return 0;
    }
    up(h, CB_ARR_X_TEMP, S.upd_x_ip, 3 * NJ);
    run_forces(h, pdd, pf_temp, pdlpf, *pitecnt, &fr, &sh);
    down(h, CB_ARR_EF_I, pef_i + 2 * TR, 14 * FR); down(h, CB_ARR_EFFE_I, pefFE_i, 14 * FR);
    if (ANAFLAG == 3) CK(cb_get_yldflag(h, pyldflag, 2 * FR));
    CK(cb_end_iteration(h));
    return fr;
}
void mass_fr(double *psm, double *pcarea, double *pllength, double *pistrong, double *piweak, double *pipolar,
             double *piwarp, double *pdens, int *posflag, double *poffset, double *px, double *pxfr, long *pminc,
             long *pmcode, double *pjac)
{
    (void)pcarea; (void)pistrong; (void)piweak; (void)pipolar; (void)piwarp; (void)pdens; (void)posflag; (void)poffset;
    (void)pminc; (void)pmcode; (void)pjac;
    cb_handle *h = handle(T_FR);
    double *b = buf(NEQ);
    up(h, CB_ARR_X, px, 3 * NJ);
    CK(cb_mass(h)); CK(cb_get_mass(h, b));
    for (long i = 0; i < NEQ; ++i) psm[i] += b[i];
    down(h, CB_ARR_LLENGTH, pllength + NE_TR, NE_FR); down(h, CB_ARR_XFR, pxfr, 6 * NE_FR);   /* frame.c:1331-1343 */
}

/* ---- DKT shells --------------------------------------------------------------------------------------- */
static void shell_state_up(cb_handle *h, double *pfarea, double *pslength, double *pdeffarea_ip, double *pdefslen_ip,
                           double *pc1_ip, double *pc2_ip, double *pc3_ip, double *pef_ip, double *pchi, double *pefN,
                           double *pefM)
{
    const long o = NE_TR + 3 * NE_FR, SH = NE_SH;
    up(h, CB_ARR_FAREA, pfarea, SH); up(h, CB_ARR_SLENGTH, pslength, 3 * SH);
    up(h, CB_ARR_DEFFAREA_IP, pdeffarea_ip, SH); up(h, CB_ARR_DEFSLEN_IP, pdefslen_ip, 3 * SH);
    up(h, CB_ARR_C1_IP, pc1_ip + o, 3 * SH); up(h, CB_ARR_C2_IP, pc2_ip + o, 3 * SH); up(h, CB_ARR_C3_IP, pc3_ip + o, 3 * SH);
    up(h, CB_ARR_EF_IP, pef_ip + 2 * NE_TR + 14 * NE_FR, 18 * SH);
    if (ANAFLAG == 3) { up(h, CB_ARR_CHI_TEMP, pchi, 3 * SH); up(h, CB_ARR_EFN_TEMP, pefN, 9 * SH); up(h, CB_ARR_EFM_TEMP, pefM, 9 * SH); }
}
void stiff_sh(double *pss, double *pemod, double *pnu, double *px_temp, double *pxlocal, double *pthick, double *pfarea,
              double *pdeffarea_ip, double *pslength, double *pdefslen_ip, double *pyield, double *pc1_ip,
              double *pc2_ip, double *pc3_ip, double *pef_ip, double *pd_temp, double *pchi_temp, double *pefN_temp,
              double *pefM_temp, long *pmaxa, long *pminc, long *pmcode)
{
    (void)pemod; (void)pnu; (void)pxlocal; (void)pthick; (void)pyield; (void)pd_temp; (void)pmaxa; (void)pminc; (void)pmcode;
    cb_handle *h = handle(T_SH);
    up(h, CB_ARR_X_TEMP, px_temp, 3 * NJ);
    shell_state_up(h, pfarea, pslength, pdeffarea_ip, pdefslen_ip, pc1_ip, pc2_ip, pc3_ip, pef_ip, pchi_temp, pefN_temp,
                   pefM_temp);
    add_skyline(h, pss);
}
int forces_sh(double *pf_temp, double *pef_ip, double *pef_i, double *pefN_temp, double *pefM_temp, double *pdd,
              double *pd_temp, double *pchi_temp, double *px_temp, double *px_ip, double *pemod, double *pnu,
              double *pxlocal, double *pthick, double *pfarea, double *pdeffarea_ip, double *pslength,
              double *pdefslen_ip, double *pyield, double *pc1_ip, double *pc2_ip, double *pc3_ip, double *pc1_i,
              double *pc2_i, double *pc3_i, long *pminc, long *pmcode, long *pjcode)
{
    (void)pd_temp; (void)px_temp; (void)pemod; (void)pnu; (void)pxlocal; (void)pthick; (void)pyield; (void)pc1_i;
    (void)pc2_i; (void)pc3_i; (void)pminc; (void)pmcode; (void)pjcode;
    cb_handle *h = handle(T_SH);
    const long SH = NE_SH, oe = 2 * NE_TR + 14 * NE_FR;
    int fr = 0, sh = 0;
    shell_state_up(h, pfarea, pslength, pdeffarea_ip, pdefslen_ip, pc1_ip, pc2_ip, pc3_ip, pef_ip, pchi_temp, pefN_temp,
                   pefM_temp);
    if (ANAFLAG == 1) {
        run_forces(h, pdd, pf_temp, NULL, 0, &fr, &sh);
        down(h, CB_ARR_EF, pef_i + oe, 18 * SH);
        return 0;
    }
    up(h, CB_ARR_X_TEMP, px_ip, 3 * NJ);            /* the coordinates updatc started from */
    run_forces(h, pdd, pf_temp, NULL, 0, &fr, &sh);
    down(h, CB_ARR_EF_I, pef_i + oe, 18 * SH);
    if (ANAFLAG == 3) {
        down(h, CB_ARR_CHI_TEMP, pchi_temp, 3 * SH); down(h, CB_ARR_EFN_TEMP, pefN_temp, 9 * SH);
        down(h, CB_ARR_EFM_TEMP, pefM_temp, 9 * SH);
    }
    CK(cb_end_iteration(h));
    return sh;
}
void mass_sh(double *psm, double *pcarea, double *pdens, double *pthick, double *pfarea, double *pslength, double *px,
             long *pminc, long *pmcode, double *pjac)
{
    (void)pcarea; (void)pdens; (void)pthick; (void)pminc; (void)pmcode; (void)pjac;
    cb_handle *h = handle(T_SH);
    double *b = buf(NEQ);
    up(h, CB_ARR_X, px, 3 * NJ);
    CK(cb_mass(h)); CK(cb_get_mass(h, b));
    for (long i = 0; i < NEQ; ++i) psm[i] += b[i];
    down(h, CB_ARR_FAREA, pfarea, NE_SH); down(h, CB_ARR_SLENGTH, pslength, 3 * NE_SH);      /* shell.c:1533-1539 */
}
