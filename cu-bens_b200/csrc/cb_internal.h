// cb_internal.h - device-side data layout shared by the translation units of libcubens_b200.
//
// Layout in HBM:
//   nodal arrays  AoS xyz  [NJ][3] doubles, three generations (committed / temp / ip)
//   jcode         [NJ][8] int32 (7 used) so one node's equation numbers are two 16-byte loads
//   per-element shell arrays touched by one thread per element are component-major (SoA):
//   constants [12][NE], frame state [10][NE] and end forces [18][NE] per generation, deformed side
//   lengths [3][NE], DKT bending matrix [81][NE]; the stiffness-pass record krec [NE][18] and the
//   DKT blocks in assembly order (kebc): general tile kernel [ncontrib][10] AoS; duo plan
//   work-major SoA, kebc[(T*18 + u*9+i)*128 + t] for work item t of tile T (coalesced).
//   NEQ vectors   dd, f_temp, d_temp, d, f, sm.
//   tangent matrix: CSC values Ax[nnz] (node-block structural pattern) and/or the reference's
//   skyline vector ss[lss].
#ifndef CB_INTERNAL_H
#define CB_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#define CB_T_TRUSS 0
#define CB_T_FRAME 1
#define CB_T_SHELL 2
#define CB_T_BRICK 3
#define CB_T_COUPLE 4    // fluid-structure interface joint (ANAFLAG 4): a structural joint and its pressure twin

#define CB_SH_CONST 12   // E, nu, t, t^3 (libm pow on the host), A0, x2, x3, y3, l12, l23, l31, pad
#define CB_SH_FRAME 10   // c1xyz c2xyz c3xyz, deformed area
#define CB_FR_CONST 16   // E, G, A, L0, L0^2(libm), L0^3(libm), Iz, Iy, J, Cw, aux xyz, pad
#define CB_FR_FRAME 10   // c1xyz c2xyz c3xyz, deformed length
#define CB_TR_CONST 4    // E, A, L0, L0^3(libm)
#define CB_TR_FRAME 4    // c1 c2 c3 deformed length

// one contribution of an element block to a node-pair block of the global matrix
struct CbContrib {
    int32_t e;       // type-local element index
    uint8_t type;    // CB_T_*
    uint8_t a;       // local node of the row block
    uint8_t b;       // local node of the column block
    uint8_t pad;
};

// one node-pair block (row node A, column node B) = one thread of the assembly kernel
struct CbPair {
    int32_t off;      // CSC: index of (first free row of A, first free column of B) in Ax
    int32_t colh;     // CSC: column height of node B's columns
    int32_t cstart;   // first contribution
    int32_t eqA0;     // first equation (1-based) of node A, 0 if none
    int32_t eqB0;     // first equation (1-based) of node B
    uint16_t ccount;  // number of contributions
    uint8_t maskA;    // free-DOF mask of node A (bit r = DOF r free), 7 bits
    uint8_t maskB;
};

// ---- tile assembly (CSC): one CTA owns a run of consecutive joints ------------------------
#ifndef CB_TILE_T
#define CB_TILE_T 128            // threads per CTA = max contributions per tile
#endif
struct CbTile {
    int64_t out0;     // first Ax index of the tile's (contiguous) output range
    int32_t nout;     // number of Ax entries
    int32_t c0;       // first contribution (reference order: the order phase 2 sums in)
    int32_t p0;       // first pair
    int32_t t0;       // first thread slot in tcontribs
    uint16_t nc;      // contributions (<= CB_TILE_T)
    uint16_t np;      // pairs
    uint16_t ns;      // thread slots (nc <= ns <= CB_TILE_T): the contributions regrouped so that the
                      // lanes of a warp evaluate the same kind of block (element type, local joints);
                      // slot record = CbContrib with pad = tile-local contribution index, type 0xff idle
    uint16_t pad;
};
static_assert(sizeof(CbTile) == 32, "tile record");
struct CbTPair {
    int32_t rel;      // offset of (first free row of A, first free column of B) in the tile output
    int32_t colh;     // column height of joint B
    uint16_t cs;      // first contribution, relative to the tile
    uint16_t cnt;     // contributions
    uint8_t maskA, maskB;
    uint16_t pad;
};

// where a thread slot's block lands when it is its joint pair's ONLY contribution and every DOF of
// both joints is free: straight in the tile's output image (frame-only tile kernel), no staging
struct CbTDst {
    int32_t rel;      // CbTPair::rel of the slot's joint pair
    uint16_t colh;    // CbTPair::colh
    uint16_t direct;  // 1: write the block at image[rel + j * colh + i]; 0: stage + reduce
};

// Per-element shell arrays read or written by one thread per element are stored component-major
// ("SoA": component c of element e at [c*NE + e]) so that a warp's accesses are fully coalesced.
#define SOA(p, comp, e, ne) (p)[(long)(comp) * (ne) + (e)]
// staging of element forces in global axes for the joint gather: [local joint][element][6]
#define CB_FG(p, b, e, ne) ((p) + ((long)(b) * (ne) + (e)) * 6)
// DKT bending matrix ke_b[9][9]: component index = nine 3x3 joint-pair blocks, block-major
#define CB_KEB(i, j) ((((i) / 3) * 3 + ((j) / 3)) * 9 + ((i) % 3) * 3 + ((j) % 3))
// ---- shell-only tile assembly ("duo" plan): a thread evaluates up to two consecutive
// contributions of one joint-pair block and sums them in registers; blocks with more than two
// contributions are split into partial sums that are combined through shared memory
#ifndef CB_T2_T
#define CB_T2_T 96               // threads per CTA of the duo kernel = work items per tile: 4 CTAs of 3 warps per SM
                                 // (0.770 ms) beat 3 of 4 warps (0.790) and 2 of 6 (0.78): more independent CTAs
                                 // hide each other's barrier stalls, and 10 joints fill 90 of 96 lanes
#endif
#ifndef CB_T2_OUT
#define CB_T2_OUT 2560           // max doubles of Ax per tile staged in shared memory (10 plate joints x 252)
#endif
#ifndef CB_T2_ELEMS
#define CB_T2_ELEMS 44           // max distinct shells per tile (their krec records are staged)
#endif
#define CB_T2_GROUP 16           // max partial sums of one block (a group of lanes of one warp)
struct CbTile2 {
    int64_t out0;     // first Ax index of the tile's contiguous output range
    int32_t nout;
    int32_t w0, nw;   // work items (<= CB_TILE_T)
    int32_t p0, np;   // pair records
    int32_t nm;       // (unused)
    int32_t e0, ne;   // distinct shells of the tile
};
struct CbWork {       // 16 bytes, self-contained: no dependent load of the contribution records
    int32_t c0;       // first contribution (the second, if any, is c0 + 1)
    uint8_t n;        // 1 or 2 contributions
    uint8_t kind;     // 0: complete block; 2: leader of a group of pad0 partial sums of one block (its
                      // followers, kind 3 with pad1 = 1, 2, ... sit in the next lanes of the same
                      // warp and add to the image in that order); 4: idle lane
    uint16_t dst;     // pair record index inside the tile
    uint8_t a0, b0, s0, pad0;   // local row / column joint and tile-local shell slot, contribution 0
    uint8_t a1, b1, s1, pad1;   // ... contribution 1
};

static_assert(sizeof(CbWork) == 16 && sizeof(CbTPair) == 16, "records are copied as 16-byte units");
static_assert(sizeof(CbTile2) == 40, "tile records are prefetched as five 8-byte words");

// ---- shell-only tile assembly, "stream" plan (k_assemble_shell_stream): one WARP owns a tile (a run of
// consecutive joints = one contiguous slice of Ax).  Every lane walks a short list of contributions
// (steps) that covers WHOLE joint-pair blocks, one after the other: it accumulates a block's 6x6 in
// registers over the block's contributions (reference order) and stores it into the tile image when
// the block's last contribution has been added - a segmented reduction over the sorted
// element-to-nonzero map with the segments aligned to lanes: every shell record is read once per
// contribution, every image entry is written once.  A block with more contributions than a lane has
// steps is cut in two: its second part ("follower") is the last block of another lane, which adds its
// sum to the stored first part once, at the end of the tile.
// Two compiled shapes (the planner is told which): steps per lane S, image / shell-slot / pair capacity.
//   wide   S = 6: 10 plate joints per tile (3 lanes each: 2+2+2 | 6 | 2+2+2), 6 warps per SM
//   narrow S = 4:  6 plate joints per tile (5 lanes each: 2+2 | 2+2 | 2+2 | 3 | 3 follower), 8 warps per SM
// Everything a tile needs sits at a fixed stride of its number, so nothing but the tile number is needed
// to prefetch it: steps[(tile * S + step) * 32 + lane], pairs[tile * PAIRS + k], elems[tile * SLOTS + slot]
// (padded with a valid shell), kebc[((tile * S + step) * 9 + i) * 32 + lane].
struct CbStreamShape {
    int img;          // max doubles of Ax per tile
    int slots;        // max distinct shells per tile
    int steps;        // contributions per lane and tile
    int pairs;        // max joint-pair blocks per tile (multiple of 4)
    int warps;        // warps (= concurrent tiles) per CTA, one CTA per SM
};
#define CB_S_MAXSTEPS_ANY 6       // the larger S of the two shapes
#define CB_S_SHAPE_WIDE   {2560, 44, 6, 80, 6}
#define CB_S_SHAPE_NARROW {1536, 28, 4, 48, 8}
#define CB_S_SHAPE_MINI   {1024, 20, 3, 32, 12}      // 12 warps / SM at 168 registers (no register prefetches)
#define CB_S_SHAPE_NARROW12 {1536, 28, 4, 48, 12}    // the narrow plan, 12 warps with single record buffers
struct CbTileS {
    int64_t out0;     // first Ax index of the tile's contiguous output range
    int32_t nout;
    uint8_t nsteps, np, ne, pad;
};
static_assert(sizeof(CbTileS) == 16, "tile records are read as one 16-byte word");
// step record (uint32): bits 0-5 shell slot (CB_S_IDLE = no work), 6-7 local row joint a, 8-9 local column
// joint b, 10 = last contribution of its block (or block part), 11 = this part is a follower (held in
// registers and added to the image at the end of the tile), 12-18 pair record index, 19 = first contribution
// of its block part (the running sum restarts), 20-31 geometry class of the shell (0 when the classes are
// off).  An idle record is all zero but for the slot field.
#define CB_S_IDLE 63u
#define CB_S_REC(slot, a, b, first, last, follow, dst, cls) \
    ((uint32_t)(slot) | ((uint32_t)(a) << 6) | ((uint32_t)(b) << 8) | ((uint32_t)(last) << 10) | \
     ((uint32_t)(follow) << 11) | ((uint32_t)(dst) << 12) | ((uint32_t)(first) << 19) | ((uint32_t)(cls) << 20))
// pair record (uint32): bits 0-11 offset of (first free row of A, first free column of B) in the tile
// image, 12-19 column height of joint B, 20-25 free-DOF mask of A, 26-31 of B
#define CB_S_PAIR(rel, colh, ma, mb) \
    ((uint32_t)(rel) | ((uint32_t)(colh) << 12) | ((uint32_t)(ma) << 20) | ((uint32_t)(mb) << 26))

#define CB_SH_DER 24
#define CB_KROW 18                // doubles per (class, a, b) row of keb_row
#define CB_SH_KREC 18   // per-shell record for the stiffness pass: R[9], X2,X3,Y3, cm00,cm01,cm22, n0,n1,n2

// one element corner touching a node (node -> corner CSR), used by the f_int / mass gathers
struct CbCorner {
    int32_t e;
    uint8_t type;
    uint8_t b;       // local node index inside the element
    uint8_t pad[2];
};

struct CbGenShell {      // one generation of shell state
    double *frame;       // [NE][CB_SH_FRAME]
    double *dsl;         // [NE][3] deformed side lengths
};
struct CbGenFrame {
    double *frame;       // [NE][CB_FR_FRAME]
    double *xfr;         // [NE][6]
    double *efFE;        // [NE][14]
};
struct CbGenTruss {
    double *frame;       // [NE][CB_TR_FRAME]
};

// everything a kernel needs, passed by value
struct CbDev {
    long NJ, NEQ;
    long NE_TR, NE_FR, NE_SH, NE_BR;
    long NE_SBR;             // bricks [NE_SBR, NE_BR) are FLUID bricks (acoustic, one pressure DOF per joint)
    int ANAFLAG;
    // nodes
    const int32_t *jc;       // [NJ][8]
    // shells
    const int32_t *sh_nodes; // [NE][4] 0-based
    const uint8_t *sh_own;   // [NE] bit a: this shell is the first element at its local joint a (it writes the
                             // joint's updated coordinates when the nodal update is fused into the force kernel)
    const double *sh_const;  // [CB_SH_CONST][NE]
    const double *sh_keb;    // [81][NE], component = CB_KEB(i,j)
    double *sh_der;          // [CB_SH_DER][NE] geometry-constant derived data (k_shell_init_keb):
                             // 0-17 ke_m[i][{2,4,5}] (the membrane columns dm multiplies), 18-20
                             // plane-stress C00,C01,C22, 21-23 t/(4 A0) * C
    double *sh_Nm;           // [NE][CB_SH_KREC] stiffness-pass record written by k_shell_prep
    double *sh_fg;           // [3][NE][6] element force in global axes (staging for the gather)
    // warp-level partial sums of the shell force pass (cb_wsum.cuh; null: every corner is staged in sh_fg)
    const int32_t *ws_start;              // [nwarp + 1] first slot of every warp of 32 consecutive shells
    const unsigned long long *ws_corners; // [nslot] corners summed into the slot: count (3 bits), 7 bits each
    double *ws_fg;                        // [nslot][6] the staged sums
    const int32_t *js_start, *js_slots;   // [NJ + 1], [nslot] the slots of every joint, in warp order
    long ws_nwarp;
    // geometry classes: shells whose geometry-constant inputs (E, nu, t, A0, local coordinates, side
    // lengths) are bit-identical share one copy of the DKT matrix and of sh_der - structured meshes
    // have a handful of classes, so the per-element 105 doubles never leave L1.  nullptr when the
    // model has too many classes (or after mass_* rewrote the reference geometry, App. B.5).
    const int32_t *sh_class; // [NE] class of each shell
    const double *keb_tab;   // [ncls][81] component order of sh_keb (CB_KEB)
    const double *keb_tab10; // [ncls][9][10] the same 3x3 blocks padded to ten doubles (16-byte loads)
    const double *keb_row;   // [ncls][9][CB_KROW] per local joint pair (a, b): DKT block, drilling term, material
                             // membrane block, gradient products of the geometric stiffness (k_class_tables)
    const double *der_tab;   // [ncls][CB_SH_DER]
    // frames
    const int32_t *fr_nodes; // [NE][2]
    const double *fr_const;  // [NE][CB_FR_CONST]
    const double *fr_offset; // [NE][6]
    const int32_t *fr_osflag;
    const int32_t *fr_mendrel; // [NE][5]
    int fr_simple;           // no member has rigid offsets or end releases and ANAFLAG != 3: the force
                             // pass runs its register-only specialisation (k_frame_forces<*, true>)
    const double *fr_efFE_ref; // [NE][14]
    double *fr_fg;           // [NE][14]
    // ANAFLAG 3 (concentrated plasticity), else nullptr
    const double *fr_plast;  // [NE][3] squash load Py, plastic moments Mpy (weak), Mpz (strong)
    int32_t *fr_yldflag;     // [NE][2] 0 elastic, 1 on the yield surface, 2 unloading (frame.c:1184-1268)
    int32_t *fr_ynew;        // [NE][2] flags proposed by the force pass (committed up to the first trip)
    int32_t *fr_code;        // [NE] return code of the force pass for this member: 0, 1, 2
    double *fr_tau;          // [NE] regula-falsi scale of dlpf when code == 1
    // element-partitioned runs (cb_set_element_ids): GLOBAL index of every local member / shell, so
    // that "the first element that trips"  means the same element on every rank;
    // nullptr = local index
    const int32_t *fr_gid, *sh_gid;
    int32_t *fr_trip;        // [4] lowest member index with code != 0 (INT_MAX if none), its code
    const double *tr_py;     // [NE_TR] squash loads of the trusses
    // shells, ANAFLAG 3 (Ivanov yield criterion in stress resultants)
    const double *sh_yield;  // [NE] yield stress
    double *sh_pl;           // [NE][21] chi[3], efN[3][3], efM[3][3]: the *_temp generation
    int32_t *sh_yv;          // [NE] controlling yielded vertex (1..3) for the stiffness pass, 0 elastic
    double *sh_kpl;          // [NE][18*18] local elasto-plastic matrix (stiffm_sh) of yielded shells
    int32_t *sh_trip;        // [4] lowest shell index whose force pass returned 1 (INT_MAX if none)
    // trusses
    const int32_t *tr_nodes; // [NE][2]
    const double *tr_const;  // [NE][CB_TR_CONST]
    double *tr_fg;           // [NE][6]
    // bricks
    const int32_t *br_nodes; // [NE][8]
    const double *br_const;  // [NE][4]  E, nu, rho, pad
    // acoustic FSI (fsi.c): coupling vector of every wet joint, L = tributary area * unit normal (fsi.c:499-525),
    // and the fluid density that scales -L^T in the "mass" matrix (fsi.c:436-443)
    const double *cp_L;      // [ncouple][4]  Lx Ly Lz pad
    double fdens;
};

// Launch configuration that is a property of the DEVICE (cudaFuncSetAttribute opt-ins above 48 KB of
// dynamic shared memory, persistent grid = resident CTAs x SM count) is cached per device ordinal, not
// per process: cb_flags.device makes several devices in one process a supported configuration.
#define CB_MAX_DEVICES 64
struct CbPerDevice {
    int v[CB_MAX_DEVICES];           // 0 = not configured yet on that device
};
static inline int cb_device_slot()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < CB_MAX_DEVICES) ? dev : 0;
}

// ---- host launch wrappers implemented in the .cu files ------------------------------------
struct CbStiffArgs {
    CbDev d;
    const double *x;         // coordinates the stiffness is evaluated at (x_temp or x)
    const double *sh_frame;  // [CB_SH_FRAME][NE] generation read
    const double *sh_ef;     // unused for ANAFLAG<=2 shells
    const double *fr_frame, *fr_ef, *fr_efFE;
    const double *tr_frame, *tr_ef;
    const CbPair *pairs; long npairs;
    const CbContrib *contribs;
    const CbTile *tiles; long ntiles; const CbTPair *tpairs;
    const CbContrib *tcontribs;   // thread slots of the general tile kernel (see CbTile)
    const CbTDst *tdst;           // [slot] direct-write target of the slot (see CbTDst)
    double *br_prep;              // [NE_BR][8 Gauss points][10]: J^-1 and detJ (k_brick_prep)
    const double *kebc;      // DKT 3x3 sub-blocks in assembly order (static; layouts above)
    const CbTile2 *tiles2; long ntiles2; const CbWork *works; const CbTPair *tpairs2;
    const int32_t *tile_elems;
    const CbTileS *tilesS; long ntilesS; const uint32_t *stepsS; const uint32_t *pairsS;   // stream plan
    const int32_t *elemsS; int shapeS;                                                     // 0 wide, 1 narrow
    int tile_smem_out;       // doubles of output staging per tile
    int max_dof;             // 3, 6 or 7: largest DOF count per joint among the model's elements
    int mixed;               // element types with different DOF counts per joint are present
    double *out;             // Ax or ss
    int out_par;             // parity of `out` in 16-byte units (Ax.p + 1 double when most blocks sit on odd indices)
    const long *maxa;        // device copy (skyline mode) or nullptr
    int skyline;
    // mass mode (cb_mass with bricks): the same tiles assemble the full-order mass matrix of the
    // reference's dense layout - consistent brick mass (mass_br, brick.c:399-537), lumped shell
    // mass on the diagonal (mass_sh, shell.c:1576-1588) - onto the CSC pattern of K_t
    int mass_mode;
    const double *sh_dens;   // [NE_SH]
};

struct CbForceArgs {
    long axpy_n;             // d_temp[i] += dd[i] for i < axpy_n, done by the nodal kernel (0: none)
    const double *axpy_x;
    double *axpy_y;
    CbDev d;
    double *x_temp, *x_ip;
    const double *dd;        // [NEQ] device
    // shell-only geometric-nonlinear models: updatc's nodal part (misc.c:83-93) is evaluated inside the force
    // kernel (every shell forms x_temp + dd of its joints anyway; the first shell at a joint writes it) into
    // the buffer that held x_ip, which then becomes x_temp by renaming - and d_temp += dd (main.c:1949)
    // rides in the gather: two launches per force pass instead of three
    int fuse_node;
    double *d_temp;
    // shells
    const double *sh_frame_ip; double *sh_frame_i; double *sh_dsl_i; const double *sh_dsl_ip;
    const double *sh_ef_ip; double *sh_ef_i;
    // frames
    const double *fr_frame_ip; double *fr_frame_i; double *fr_xfr_i;
    const double *fr_ef_ip; double *fr_ef_i;
    const double *fr_efFE_ip; double *fr_efFE_i;
    // trusses
    double *tr_frame_i; double *tr_ef_i;
    double dlpf; int itecnt;
    // joints touched by this rank's elements [jl0, jl1), joints it owns [jo0, jo1), equations of the
    // touched joints [ql0, ql1): the nodal kernels run over these ranges only, so that with the mesh
    // partitioned over several GPUs (global numbering kept) their cost does not grow with the ranks
    long jl0, jl1, jo0, jo1, ql0, ql1;
    // gather
    const int32_t *node_cstart; const CbCorner *corners;
    double *f_temp;
};

#ifdef __cplusplus
extern "C++" {
#endif
// cb_forces.cu  (compiled with -fmad=false: reference operation order, IEEE mul/add)
int cbk_shell_init_keb(const CbDev &d, double *keb_out, cudaStream_t s);
int cbk_shell_init_kebc2(const CbDev &d, const CbTile2 *tiles, long ntiles, const CbWork *works,
                         const CbContrib *contribs, double *kebc, cudaStream_t s);
int cbk_shell_init_kebc(const CbDev &d, const CbContrib *contribs, long ncontrib, double *kebc,
                        cudaStream_t s);
int cbk_shell_init_kebcS(const CbDev &d, const CbTileS *tiles, long ntiles, int steps_per_tile, int slots_per_tile,
                         const uint32_t *steps, const int32_t *elems, double *kebc, cudaStream_t s);
int cbk_shell_prep(const CbDev &d, const double *x, const double *sh_frame, cudaStream_t s);
int cbk_shell_class_tables(const CbDev &d, const int32_t *rep, int ncls, double *keb_tab, double *keb_tab10, double *keb_row, double *der_tab,
                           const CbWork *works, long nworks, const CbContrib *contribs, CbWork *works_cls,
                           cudaStream_t s);
int cbk_shell_plastic_prep(const CbDev &d, const double *sh_frame, const double *sh_dsl,
                           const double *sh_pl, cudaStream_t s);
int cbk_node_update(const CbForceArgs &a, cudaStream_t s);
int cbk_forces(const CbForceArgs &a, cudaStream_t s, long *launches);
int cbk_frame_trip(const CbDev &d, cudaStream_t s, long *launches);
int cbk_forces_linear(const CbForceArgs &a, const double *d_total, cudaStream_t s, long *launches);
int cbk_gather_f(const CbForceArgs &a, cudaStream_t s);
int cbk_mass(const CbDev &d, const double *x, double *sh_const_mut, double *tr_const_mut,
             double *fr_const_mut, double *fr_xfr, const double *dens_tr, const double *dens_fr,
             const double *dens_sh, const int32_t *node_cstart, const CbCorner *corners,
             double *sm, cudaStream_t s, long *launches);
// cb_stiff.cu  (FMA contraction allowed)
int cbk_stiff(const CbStiffArgs &a, cudaStream_t s, long *launches);
#ifdef __cplusplus
}
#endif

#endif
