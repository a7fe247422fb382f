/*
 * cubens_b200.h - C-ABI of the B200-native element / assembly hot path of CU-BENs.
 *
 * The reference has no plugin or FFI layer: its boundary is the set of plain C routines of
 * prototypes.h that main.c calls once per Newton iteration with 12-32 raw pointers each, all
 * sizes and flags living in file-scope globals (main.c:323-328).  This header is the drop-in
 * for exactly those call sites.  Every entry point is `extern "C"`, takes plain pointers and
 * sizes in the reference's own host layout (1-based `long` joint/equation
 * numbers, AoS xyz, per-type offsets inside the shared emod/c1/ef/mcode/minc arrays) and
 * returns an int status; the caller keeps ownership of every host pointer.  INTEGRATION.md
 * shows the edits to main.c / solve.c.
 *
 *   reference call (file:line)                         replacement
 *   -------------------------------------------------  ---------------------------------------
 *   main.c:491-1437 allocation + prop_* results         cb_create (uploads, builds the maps)
 *   main.c:1833-1882 generation copies at increment     cb_begin_increment
 *   main.c:1899-1921 ss=0; stiff_tr; stiff_fr; stiff_sh cb_stiff            (+ stiff_br, brick.c:79)
 *       truss.c:82  frame.c:226  shell.c:110  misc.c:41
 *   main.c:3590-3619 sm=0; mass_tr; mass_fr; mass_sh    cb_mass
 *   main.c:1948-1984 d_temp+=dd; f_temp=0; updatc;      cb_update_forces
 *       forces_tr; forces_fr; forces_sh; ef_ip=ef_i
 *       misc.c:71  truss.c:231  frame.c:902  shell.c:1593
 *   main.c:1774-1793 linear force recovery              cb_forces_linear
 *   main.c:2006-2028 _ip <- _i                          cb_end_iteration
 *   main.c:2074-2134 commit converged increment         cb_commit
 *   solve.c:110-119  dense -> Ap/Ai/Ax scan             cb_csc_pattern / cb_get_csc_values /
 *                                                       cb_csc_compact (device-assembled CSC)
 *   ss[] skyline layout (model.c:1269-1278)             cb_get_skyline
 *   output()/checkPoint() reads (misc.c:345,494)        cb_download / cb_upload
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device and returns
 * CB_ERR_CUDA if none is usable.
 */
#ifndef CUBENS_B200_H
#define CUBENS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CB_ABI_VERSION 5

/* status codes (the reference uses 0 = ok / 1 = error, solve.c:558-562, frame.c:1201) */
enum {
    CB_OK = 0,
    CB_ERR_ARG = 1,          /* bad argument / inconsistent sizes                         */
    CB_ERR_CUDA = 2,         /* no device, allocation or launch failure                   */
    CB_ERR_UNSUPPORTED = 3,  /* feature of the reference outside this build (see DESIGN)  */
    CB_ERR_OVERFLOW = 4      /* a count does not fit the device's 32-bit indices          */
};

/* which generation of the updated-Lagrangian state a call reads */
enum { CB_GEN_IP = 0, CB_GEN_COMMITTED = 1 };

/* layout of the assembled tangent matrix */
enum {
    CB_MAT_CSC = 1,      /* full (unsymmetric-storage) CSC, structural node-block pattern  */
    CB_MAT_SKYLINE = 2,  /* the reference's skyline vector ss[lss] addressed through maxa  */
    CB_MAT_BOTH = 3
};

/* the sizes main.c reads from the deck (main.c:384-427) + NEQ from codes() (model.c:943) */
typedef struct cb_sizes {
    long NJ, NE_TR, NE_FR, NE_SH, NE_SBR, NE_FBR, NEQ;
} cb_sizes;

/* ANAFLAG 1 = first-order elastic, 2 = geometric nonlinear, 3 = geometric + material nonlinear
 * (main.c:63-90): trusses and frames with concentrated plasticity (stiffm_tr, stiffm_fr, yield check /
 * regula_falsi / unload in forces_fr), DKT shells with Ivanov's yield criterion in stress resultants
 * (stiffm_sh, strn_curv, the return mapping of forces_sh).  ANAFLAG 4 (acoustic FSI, fsi.c: linear, assembled
 * once): cb_stiff(CB_GEN_COMMITTED) assembles [K L; 0 H] and cb_mass [M 0; -rho L^T Q] on one CSC pattern
 * (cb_get_csc_values / cb_get_mass_csc_values) - sparse, where the reference holds dense NEQ^2 arrays
 * (stiff_fsi / mass_fsi, fsi.c:333-445); the force pass is not part of that analysis.                      */
typedef struct cb_flags {
    int ANAFLAG, ALGFLAG, SLVFLAG;
    int matrix_layout;      /* CB_MAT_*; 0 picks SKYLINE when SLVFLAG==0 else CSC          */
    int device;             /* CUDA device ordinal                                         */
} cb_flags;

/* Host arrays exactly as main.c holds them after struc/codes/skylin/prop_* (App. A).
 * Pointers for absent element types may be NULL.  All are read during cb_create only.    */
typedef struct cb_model {
    const double *x;         /* [NJ*3]                                main.c:491            */
    const long   *minc;      /* [2TR+2FR+3SH+8BR] 1-based joints      model.c:95-137        */
    const long   *jcode;     /* [NJ*7] 0 = fixed, else equation no.   model.c:941-987       */
    const long   *mcode;     /* [6TR+14FR+18SH+24BR]                  model.c:992-1142      */
    const long   *maxa;      /* [NEQ+1] (skyline layouts only)        model.c:1269-1278     */
    const double *emod;      /* [TR+FR+SH+BR]                                               */
    const double *dens;      /* [TR+FR+SH+BR], indexed [n] per type like the reference      */
    const double *carea;     /* [TR+FR]                                                     */
    const double *llength;   /* [TR+FR]   initial lengths from prop_tr / prop_fr            */
    const double *c1, *c2, *c3;   /* [TR+3FR+3SH] initial direction cosines                 */
    /* shells (prop_sh, shell.c:43-108) */
    const double *nu;        /* [SH+BR]                                                     */
    const double *thick;     /* [SH]                                                        */
    const double *farea;     /* [SH]                                                        */
    const double *slength;   /* [SH*3]                                                      */
    const double *xlocal;    /* [SH*3]                                                      */
    /* frames (prop_fr, frame.c:43-224) */
    const double *gmod, *istrong, *iweak, *ipolar, *iwarp;   /* [FR]                        */
    const double *auxpt;     /* [FR*3]                                                      */
    const double *offset;    /* [FR*6]                                                      */
    const int    *osflag;    /* [FR]                                                        */
    const int    *mendrel;   /* [FR*5]                                                      */
    const double *efFE_ref;  /* [FR*14] reference fixed-end forces from load()              */
    /* material nonlinear analysis (ANAFLAG 3), may be NULL otherwise */
    const double *yield;     /* [TR+FR+SH+BR] yield stress           truss.c:66 frame.c:177 */
    const double *zstrong, *zweak;   /* [FR] plastic section moduli  frame.c:177            */
    /* acoustic fluid-structure interaction (ANAFLAG 4, fsi.c), NULL otherwise: what prop_fsi leaves (fsi.c:44-331)
     * and the fluid density of prop_br (brick.c:60-75).  jcode[j][6] then holds the PRESSURE equation of joint j
     * (numbered after all structural equations, model.c:962-990), the fluid bricks are bricks NE_SBR .. NE_SBR +
     * NE_FBR - 1 with the reference's emod = 1e20, nu = 0.5e20 and dens = 1 / c^2 (brick.c:63-75)             */
    const double *nnorm;     /* [NJ*3] unit normals of the fluid-structure interface at the joints           */
    const double *tarea;     /* [NJ]   tributary interface area of the joints                                */
    const double *fdens;     /* [1]    fluid density                                                          */
} cb_model;

typedef struct cb_handle cb_handle;

/* ---- life cycle ---------------------------------------------------------------------- */
int  cb_abi_version(void);
const char *cb_last_error(void);                 /* text of the last failure on this thread */
int  cb_device_count(void);                      /* 0 when no CUDA device is usable         */

int  cb_create(const cb_sizes *sz, const cb_flags *fl, const cb_model *m, cb_handle **out);
void cb_destroy(cb_handle *h);

/* Multi-GPU: each rank creates its handle from a sub-model that keeps the GLOBAL
 * joint / equation numbering (NJ, NEQ, jcode, x) but only the elements of its partition plus
 * the halo elements that touch its owned joints.  cb_set_owned_joints then restricts the
 * assembled matrix columns, the mass and the f_int gather to joints [j0, j1) (0-based), so the
 * owned column slice is complete without any exchange of matrix entries.  Must be called
 * before the first cb_stiff / cb_update_forces.  Default: all joints.                       */
int  cb_set_owned_joints(cb_handle *h, long j0, long j1);

/* ---- the Newton-iteration hot path ---------------------------------------------------- */
int  cb_begin_increment(cb_handle *h);                                /* main.c:1833-1882  */
int  cb_stiff(cb_handle *h, int gen);                                 /* main.c:1899-1921  */
int  cb_mass(cb_handle *h);                                           /* main.c:3590-3619  */
/* dd: host [NEQ] incremental displacements from solve().  f_temp_out (may be NULL): host
 * [NEQ] internal force vector.  dlpf_inout / itecnt as forces_fr takes them (frame.c:908);
 * frcchk_fr / frcchk_sh receive the routines' return codes (0 for ANAFLAG 1, 2).          */
int  cb_update_forces(cb_handle *h, const double *dd, double *dlpf_inout, int itecnt,
                      double *f_temp_out, int *frcchk_fr, int *frcchk_sh);
/* same with dd / f_temp already resident on the device (no PCIe traffic)                  */
int  cb_update_forces_dev(cb_handle *h, const double *dd_dev, double *dlpf_inout, int itecnt,
                          int *frcchk_fr, int *frcchk_sh);
/* Element-partitioned ANAFLAG 3: forces_fr (frame.c:1184-1268) and
 * forces_sh (shell.c:2044-2046) return at the FIRST element that trips, so the ranks must agree on
 * the lowest global element index before flags are committed and f_temp is gathered.
 *   cb_set_element_ids     global (whole-model, 0-based, ascending) index of every local frame /
 *                          shell; either pointer may be NULL (= local numbering)
 *   cb_update_forces_begin = updatc + the element loops of forces_*; first_fr / first_sh (may be
 *                          NULL) receive the lowest tripping global index on this rank (INT_MAX: none)
 *   ... the caller takes the minimum of first_fr, first_sh over the ranks (one all-reduce) ...
 *   cb_update_forces_end   commits yldflag up to first_fr, gathers f_temp without the elements
 *                          from the trip on; frcchk_fr / *dlpf_inout are set where this rank holds
 *                          that member (0 / unchanged elsewhere: take the max code and the min of
 *                          dlpf over the ranks), frcchk_sh on every rank.  first_* < 0: keep the
 *                          rank-local minima (what cb_update_forces_dev does on one GPU).        */
int  cb_set_element_ids(cb_handle *h, const int *fr_gid, const int *sh_gid);
int  cb_update_forces_begin(cb_handle *h, const double *dd_dev, double dlpf, int itecnt,
                            int *first_fr, int *first_sh);
int  cb_update_forces_end(cb_handle *h, int first_fr, int first_sh, double *dlpf_inout,
                          int *frcchk_fr, int *frcchk_sh);
int  cb_forces_linear(cb_handle *h, const double *d, double *f_out);  /* main.c:1774-1793  */
int  cb_end_iteration(cb_handle *h);                                  /* main.c:2006-2028  */
/* ANAFLAG 3: the member-end yield flags forces_fr mutates (main.c:802, frame.c:1184-1268):
 * 0 elastic, 1 on the yield surface, 2 unloading.  n = 2*NE_FR.  cb_update_forces reproduces the
 * reference's early return: frcchk_fr = 1 (overshoot; *dlpf_inout rescaled by regula_falsi) or
 * 2 (elastic unloading) comes from the lowest-numbered member that trips, the flags of the
 * members up to and including it are updated, and - as in the reference - f_temp then lacks the
 * members from that one on.  cb_commit resets 2 -> 0 (main.c:2105-2113).                     */
int  cb_get_yldflag(cb_handle *h, int *yldflag, long n);
int  cb_set_yldflag(cb_handle *h, const int *yldflag, long n);
int  cb_commit(cb_handle *h);                                         /* main.c:2074-2134  */
/* arc-length only: the reference does not advance *_ip on the iteration that converges
 * (main.c:2925-2940), so the next predictor still sees the second-to-last iterate as *_ip; call
 * this after cb_commit, instead of cb_end_iteration, to reproduce that                      */
int  cb_keep_ip(cb_handle *h);

/* ---- results -------------------------------------------------------------------------- */
/* skyline vector in the reference's layout; n must equal lss = maxa[NEQ]-1                */
int  cb_get_skyline(cb_handle *h, double *ss, long n);
/* structural CSC: nnz, then pattern (int Ap[NEQ+1], Ai[nnz], 0-based as umfpack_di_*      */
/* expects) and values Ax[nnz]                                                             */
long cb_csc_nnz(cb_handle *h);
int  cb_csc_pattern(cb_handle *h, int *Ap, int *Ai);
int  cb_get_csc_values(cb_handle *h, double *Ax);
/* value-thresholded copy with the reference's rule fabs(a) > drop_tol (solve.c:112);      */
/* returns the compacted nnz (Ap/Ai/Ax sized for cb_csc_nnz), or -1                        */
long cb_csc_compact(cb_handle *h, double drop_tol, int *Ap, int *Ai, double *Ax);
/* Symmetric hand-off (K_t is symmetric; the device -> host copy of Ax is what an iteration with a
 * host-side solver waits for): the upper triangle of the owned column slice - a prefix of
 * every column, rows being ascending - is packed on the device and shipped in chunks on a second stream.
 *   cb_csc_upper_nnz / _pattern / cb_get_csc_upper_values   standard upper-triangular CSC (Apu[NEQ+1],
 *       Aiu, Axu) for symmetric solvers; on an element-partitioned handle the rows of neighbour joints
 *       past the owned range are appended to each column (their mirror images live on another rank)
 *   cb_csc_values_begin / _end      the FULL Ax[cb_csc_nnz] that umfpack_di_* takes (solve.c:122-134) from
 *       half the PCIe traffic: nthreads host threads rebuild the columns (copy of the upper part,
 *       transposes of the upper blocks of higher-numbered joints) while later chunks are on the wire.
 *       Axu_staging [cb_csc_upper_nnz] must be page-locked (cb_host_alloc).  _begin only queues the work,
 *       so the caller can run cb_update_forces meanwhile; _end waits for the matrix
 *   cb_get_csc_values_mirrored      = _begin + _end                                                    */
long cb_csc_upper_nnz(cb_handle *h);
int  cb_csc_upper_pattern(cb_handle *h, int *Apu, int *Aiu);
int  cb_get_csc_upper_values(cb_handle *h, double *Axu);
int  cb_csc_values_begin(cb_handle *h, double *Ax, double *Axu_staging, int nthreads);
int  cb_csc_values_end(cb_handle *h);
int  cb_get_csc_values_mirrored(cb_handle *h, double *Ax, double *Axu_staging, int nthreads);
/* The copy engine and the host's memory bandwidth are balanced by sending every k-th chunk of joints as
 * full columns straight into Ax (no host work) - environment CB_SYM_FULL_EVERY = k when the first of these
 * calls is made (default 5; 1 = everything in full, 0 = everything packed); Ax should be page-locked too.
 * cb_csc_values_d2h_bytes: bytes one cb_csc_values_begin moves device -> host                       */
long cb_csc_values_d2h_bytes(cb_handle *h);
int  cb_get_mass(cb_handle *h, double *sm_diag);       /* diagonal [NEQ], SLVFLAG 0 layout */
/* models with bricks: cb_mass also assembles the reference's full-order mass matrix (consistent
 * mass_br brick.c:399-537, lumped mass_sh on the diagonal shell.c:1576-1588; dense [NEQ][NEQ] in
 * the reference) on the CSC pattern of K_t: Mx[nnz], same Ap / Ai                            */
int  cb_get_mass_csc_values(cb_handle *h, double *Mx);
double *cb_dev_Mx(cb_handle *h);
int  cb_get_f(cb_handle *h, double *f_temp);           /* [NEQ]                            */

/* device-resident views for consumers that stay on the GPU (no copies): */
double *cb_dev_Ax(cb_handle *h);
double *cb_dev_skyline(cb_handle *h);
double *cb_dev_f(cb_handle *h);
double *cb_dev_dd(cb_handle *h);
const int *cb_dev_Ap(cb_handle *h);
const int *cb_dev_Ai(cb_handle *h);       /* built lazily on first use                     */

/* ---- convergence sums on the device ("next" row 2: the NEQ-vector work between the kernels) --
 * cb_set_q stages the reference load vector q [NEQ]; cb_residual_sums leaves, in the device buffer
 * cb_dev_sums() (11 doubles), what test() needs (misc.c:187-250) over the equations of the owned joints, with
 * qtot = lpf*q:  [0] unbfi = |qtot - f_temp|^2   [1] deltad = |dd|^2   [2] inteneri = dd.(qtot - f_ip)
 *                [3] totald = |d_temp|^2         [4] unbfp = |qtot - fp|^2   (f_ip: f_temp before the last
 * cb_update_forces, main.c:1941-1943; fp: the committed f), and [5..10] the reaction resultants below.
 * Fixed-order reduction in one launch (bit-reproducible).  With several GPUs cb_residual_allreduce sums
 * them over the ranks (NCCL, below) - the only per-iteration collective.
 * cb_get_sums copies the first five to the host.  cb_convergence_test is test() itself on those sums
 * (same return value and *convchk codes; all-reduce included): dd and f_temp need not cross PCIe for it. */
int     cb_set_q(cb_handle *h, const double *q);
int     cb_residual_sums(cb_handle *h, double lpf);
double *cb_dev_sums(cb_handle *h);
int     cb_get_sums(cb_handle *h, double *sums5);
int     cb_residual_sums_allreduce(cb_handle *h, double lpf);   /* sums + all-reduce; ONE launch over NVLink peer
                                                                 * memory when cb_comm_init could map the peers   */
int     cb_convergence_test(cb_handle *h, double lpf, double intener1, double toldisp, double tolforc,
                            double tolener, int *convchk, double *sums5_out /* may be NULL */);
/* cb_residual_sums also leaves, in cb_dev_sums()[5..10], the REACTION resultants of the owned joints: the
 * element forces arriving at fixed degrees of freedom, which the reference drops at mcode == 0
 * (shell.c:2393-2396, frame.c:1273-1309, truss.c:366-376), summed per direction Fx Fy Fz Mx My Mz          */
int     cb_get_reaction_sums(cb_handle *h, double *r6);

/* ---- the collective of an element-partitioned run, inside the library ----------------
 * One process (or host thread) per GPU creates its handle from its sub-model and joins an NCCL communicator:
 *   cb_comm_unique_id   rank 0 obtains the 128-byte ncclUniqueId and hands it to the others by whatever means
 *                       the host has (file, pipe, MPI_Bcast, ...)
 *   cb_comm_init        collective: every rank calls it with the same id; NCCL is bound at run time
 *                       (libnccl.so.2; a process that already holds one reuses it) - CB_ERR_UNSUPPORTED if absent
 *                       cb_comm_init also tries to map every rank's 2.8 KB mailbox into every other rank (CUDA IPC
 *                       handles through one ncclAllGather, one process per GPU on one NVLink / NVSwitch node,
 *                       CB_COMM_P2P=0 disables): with it the all-reduce below is the library's own exchange
 *                       over peer memory - fused into the sums kernel by cb_residual_sums_allreduce
 *   cb_residual_allreduce   sum over the ranks of the eleven doubles of cb_dev_sums() in place, on the handle's
 *                       stream, no host synchronisation (peer-memory exchange kernel, else ncclAllReduce): call it
 *                       after cb_residual_sums, read with cb_get_sums / cb_get_reaction_sums.  The matrix columns
 *                       and f_int of the owned joints need no exchange (halo elements); these sums are the
 *                       per-iteration traffic over NVLink
 *   cb_trip_allreduce   ANAFLAG 3: minimum over the ranks of the tripping element indices, between
 *                       cb_update_forces_begin and cb_update_forces_end
 * Without a communicator (one rank) the two all-reduce calls return at once.                              */
int  cb_comm_unique_id(void *id128);
int  cb_comm_init(cb_handle *h, const void *id128, int rank, int world);
int  cb_comm_peer_memory(cb_handle *h);      /* 1: the all-reduce runs over mapped peer memory, 0: through NCCL */
int  cb_comm_destroy(cb_handle *h);
int  cb_residual_allreduce(cb_handle *h);
int  cb_trip_allreduce(cb_handle *h, int *first_fr, int *first_sh);

/* ---- state transfer for output()/checkPoint()/restartStep() (misc.c:345,494,605) ------ */
enum {
    CB_ARR_X = 1, CB_ARR_X_TEMP, CB_ARR_X_IP,
    CB_ARR_C1, CB_ARR_C2, CB_ARR_C3,                 /* committed                            */
    CB_ARR_C1_I, CB_ARR_C2_I, CB_ARR_C3_I,
    CB_ARR_C1_IP, CB_ARR_C2_IP, CB_ARR_C3_IP,
    CB_ARR_EF, CB_ARR_EF_I, CB_ARR_EF_IP,
    CB_ARR_DEFLLEN, CB_ARR_DEFLLEN_I, CB_ARR_DEFLLEN_IP,
    CB_ARR_DEFFAREA, CB_ARR_DEFFAREA_I, CB_ARR_DEFFAREA_IP,
    CB_ARR_DEFSLEN, CB_ARR_DEFSLEN_I, CB_ARR_DEFSLEN_IP,
    CB_ARR_XFR, CB_ARR_XFR_TEMP,
    CB_ARR_EFFE, CB_ARR_EFFE_I, CB_ARR_EFFE_IP,
    CB_ARR_D, CB_ARR_D_TEMP, CB_ARR_F, CB_ARR_F_TEMP,
    CB_ARR_LLENGTH, CB_ARR_FAREA, CB_ARR_SLENGTH,    /* mass_* overwrite these (App. B.5)    */
    /* ANAFLAG 3 shells: equivalent plastic curvature [SH*3] and the membrane-force / moment
     * resultants at the three vertices [SH*9] (main.c:889-913), committed and *_temp        */
    CB_ARR_CHI, CB_ARR_CHI_TEMP, CB_ARR_EFN, CB_ARR_EFN_TEMP, CB_ARR_EFM, CB_ARR_EFM_TEMP
};
/* n = number of doubles of the reference array (checked) */
int  cb_download(cb_handle *h, int which, double *dst, long n);
int  cb_upload(cb_handle *h, int which, const double *src, long n);

/* Binary checkpoint of the committed device state (bit for bit, unlike the reference's "%e" text
 * results8.txt, misc.c:494-716): everything an increment / time step starts from, including the
 * reference geometry mass_* rewrites, yldflag and the shell plastic state.  cb_checkpoint_load
 * restores it into a handle created from the same model and ends with cb_begin_increment.  The
 * host's own vectors (uc, vc, ac, load factor, step number) stay the host's to save.            */
int  cb_checkpoint_save(cb_handle *h, const char *path);
int  cb_checkpoint_load(cb_handle *h, const char *path);

/* ---- instrumentation ------------------------------------------------------------------- */
/* kernels launched by this handle since creation (bench.py's gpu_launches)                 */
long cb_launch_count(cb_handle *h);
/* CUDA-event time in ms of the most recent cb_stiff / cb_update_forces* device work.  The hot   */
/* calls only enqueue work on the handle's stream; these getters, cb_sync and every cb_get_* / */
/* cb_download wait for it.                                                                    */
double cb_last_stiff_ms(cb_handle *h);
double cb_last_forces_ms(cb_handle *h);
/* bytes of implementation-only maps read per cb_stiff (reported next to the roofline)      */
long cb_map_bytes(cb_handle *h);
/* where the element-to-nonzero maps / CSC pattern were built (1: on the device, cb_plan_device.cuh - shell-only
 * models with the CSC layout; 0: host builder) and how long the build took in seconds.  Environment CB_PLAN=host
 * forces the host builder (the tests compare the two bit for bit).                                       */
int  cb_plan_info(cb_handle *h, double *seconds, int *on_device);
/* test hook: arrays of the shell stream plan as resident on the device (0 tiles, 1 step records, 2 pair records,
 * 3 shell slots, 4 Ai, 5 Ap); returns the size in bytes (dst may be NULL), -1 if absent                    */
long cb_debug_stream_plan(cb_handle *h, int which, void *dst);
/* equations of the joints this handle's elements touch (= NEQ unless element-partitioned): the part of
 * dd / f_temp that cb_update_forces moves between host and device                            */
long cb_local_equations(cb_handle *h);
/* number of geometry classes in use (shells whose geometry-constant inputs are bit-identical share
 * one cache-resident copy of the DKT matrix), 0 when every shell keeps its own copy: more than
 * 1024 classes, or after cb_mass rewrote the reference geometry.  Setting the
 * environment variable CB_NO_GEOMETRY_CLASSES disables the sharing.                          */
int  cb_geometry_classes(cb_handle *h);
/* CUDA-event time in ms of the block-assembly kernel alone (the dominant kernel; roofline)  */
double cb_last_assemble_ms(cb_handle *h);
/* CUDA events on this handle's stream bracketing any sequence of calls (device timeline)     */
int    cb_timer_start(cb_handle *h);
double cb_timer_stop_ms(cb_handle *h);
/* stage dd in the device buffer cb_dev_dd() for cb_update_forces_dev (host -> device copy)  */
int  cb_set_dd(cb_handle *h, const double *dd);
/* page-locked host buffers so the Ap/Ai/Ax and f_temp transfers run at full PCIe rate        */
void *cb_host_alloc(unsigned long bytes);
void  cb_host_free(void *p);
/* block the host until all work queued on this handle's stream has finished                */
int  cb_sync(cb_handle *h);
/* FP64 DFMA throughput of the device in TFLOP/s (micro-kernel, CUDA events, best of 5): the FP64
 * roofline denominator bench.py reports next to the HBM one; < 0 on failure     */
double cb_measure_fp64_tflops(int device);
/* Host-only consistency check of the element-to-nonzero maps and tile plans a model would get (joint
 * scan, CSC geometry, tile packing, lane assignment), interpreted the way the kernels walk them; needs
 * no device.  j0 = j1 = 0: all joints.  stats (may be NULL) receives nnz, tiles, step rows, pair records,
 * mean steps per tile, plan kind (3 stream, 2 duo, 1 general tiles, 0 block-owner).                 */
int  cb_plan_selfcheck(const cb_sizes *sz, const cb_flags *fl, const cb_model *m, long j0, long j1,
                       long *stats);
/* Host-only self test of the packed upper-triangle layout and of the threaded rebuild of the full
 * matrix (no device): synthetic symmetric values on the model's pattern; seconds (may be NULL) receives the
 * wall time of the rebuild with nthreads threads                                                       */
int  cb_sym_selftest(const cb_sizes *sz, const cb_flags *fl, const cb_model *m, long j0, long j1,
                     int nthreads, double *seconds);
/* the CUDA stream (cudaStream_t cast to void*) all of this handle's kernels run on         */
void *cb_stream(cb_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* CUBENS_B200_H */
