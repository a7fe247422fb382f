/* TEST INFRASTRUCTURE ONLY (see oracle/README.md) - never linked into the product library.
 *
 * In-memory harness around the UNMODIFIED reference sources (compiled where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libcubens_ref.so).  The reference keeps
 * every size and flag in file-scope globals that main.c normally owns (main.c:323-328); this
 * TU owns them instead (main.c is not linked into the library) so that the element routines
 * (stiff_sh, forces_sh, updatc, codes, skylin, prop_sh ... prototypes.h) can be called
 * directly on arrays supplied by the tests through ctypes. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* the globals of main.c:323-328 */
long NJ, NE_TR, NE_FR, NE_SH, NE_SBR, NE_FBR, NE_BR, NEQ, NBC, SNDOF, FNDOF, NTSTPS, ntstpsinpt;
double dt, ttot;
int ANAFLAG = 666, ALGFLAG, OPTFLAG, SLVFLAG, FSIFLAG, FSIINCFLAG, brFSI_FLAG, shFSI_FLAG;
FILE *IFP[4], *OFP[8];
int CHKPT, RFLAG;

static char *input_copy = NULL;

void ref_set_sizes(long nj, long ne_tr, long ne_fr, long ne_sh, long ne_sbr, long ne_fbr,
                   long neq)
{
    NJ = nj; NE_TR = ne_tr; NE_FR = ne_fr; NE_SH = ne_sh; NE_SBR = ne_sbr; NE_FBR = ne_fbr;
    NE_BR = ne_sbr + ne_fbr; NEQ = neq;
}

void ref_set_flags(int anaflag, int algflag, int slvflag, int optflag)
{
    ANAFLAG = anaflag; ALGFLAG = algflag; SLVFLAG = slvflag; OPTFLAG = optflag;
    FSIFLAG = FSIINCFLAG = brFSI_FLAG = shFSI_FLAG = 0; CHKPT = RFLAG = 0;
}

/* acoustic FSI (ANAFLAG 4): equation split of codes() (model.c:977,990) and which solids are wet */
void ref_set_fsi(long sndof, long fndof, int br_flag, int sh_flag)
{
    SNDOF = sndof; FNDOF = fndof; brFSI_FLAG = br_flag; shFSI_FLAG = sh_flag; FSIFLAG = 1;
}

long ref_get_NEQ(void)  { return NEQ; }
long ref_get_NBC(void)  { return NBC; }
void ref_set_NEQ(long n) { NEQ = n; }
void ref_set_NBC(long n) { NBC = n; }   /* skylin() counts it from the deck (model.c:1149-1201) */

/* The reference writes its echo / error text to OFP[0..7] (results1..8.txt, main.c:373-379). */
int ref_open_sinks(void)
{
    int i;
    for (i = 0; i < 8; ++i) {
        if (OFP[i] == NULL) OFP[i] = fopen("/dev/null", "w");
        if (OFP[i] == NULL) return 1;
    }
    return 0;
}

/* prop_* and load() fscanf their numbers from IFP[0]; give them an in-memory deck fragment. */
int ref_set_input_text(const char *text, long n)
{
    if (IFP[0] != NULL) { fclose(IFP[0]); IFP[0] = NULL; }
    free(input_copy);
    input_copy = (char *)malloc((size_t)n + 1);
    if (input_copy == NULL) return 1;
    memcpy(input_copy, text, (size_t)n);
    input_copy[n] = '\0';
    IFP[0] = fmemopen(input_copy, (size_t)n, "r");
    return IFP[0] == NULL;
}

void ref_close_input(void)
{
    if (IFP[0] != NULL) { fclose(IFP[0]); IFP[0] = NULL; }
    free(input_copy); input_copy = NULL;
}
