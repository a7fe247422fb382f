"""Host-side model container + the setup pieces the element path consumes.

The reference keeps its model in ~130 loose arrays allocated in ``main()`` (main.c:491-1389)
and sized by file-scope globals (main.c:323-328).  ``Model`` holds the subset that crosses
the drop-in boundary, in exactly the reference's host layout (Appendix A of SURVEY.md):
1-based ``long`` joint / equation numbers, AoS ``xyz`` coordinates, per-type offsets inside the
shared ``emod / c1 / ef / mcode / minc`` arrays.  Everything here is vectorised numpy so the
BASELINE-sized synthetic meshes (2 M shells, 5 M frames) can be set up in seconds; each routine
states which reference routine it mirrors and keeps that routine's floating-point operation
order, so the arrays are bit-identical to what ``codes / skylin / prop_*`` produce (checked in
tests/test_setup_parity.py against the compiled reference).

This module is host-side set-up only (runs once per model).  It never touches ``oracle/``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

I64 = np.int64
F64 = np.float64


@dataclass
class Model:
    # sizes (main.c:323)
    NJ: int = 0
    NE_TR: int = 0
    NE_FR: int = 0
    NE_SH: int = 0
    NE_SBR: int = 0
    NE_FBR: int = 0
    NEQ: int = 0
    # flags
    ANAFLAG: int = 2
    ALGFLAG: int = 1
    SLVFLAG: int = 0
    # topology / numbering
    x: np.ndarray = None          # [NJ*3]
    minc: np.ndarray = None       # [2TR+2FR+3SH+8BR] 1-based joints
    jcode: np.ndarray = None      # [NJ*7] 0 fixed / 1-based equation
    mcode: np.ndarray = None      # [6TR+14FR+18SH+24BR]
    maxa: np.ndarray = None       # [NEQ+1] 1-based skyline diagonal addresses
    kht: np.ndarray = None        # [NEQ]
    lss: int = 0
    # shared property arrays (type offsets as in the reference)
    emod: np.ndarray = None       # [TR+FR+SH+BR]
    yld: np.ndarray = None        # [TR+FR+SH+BR]
    dens: np.ndarray = None       # [TR+FR+SH+BR]   (reference indexes dens[n] per type, App. B.4)
    carea: np.ndarray = None      # [TR+FR]
    llength: np.ndarray = None    # [TR+FR]
    c1: np.ndarray = None         # [TR+3FR+3SH]
    c2: np.ndarray = None
    c3: np.ndarray = None
    # shell
    nu: np.ndarray = None         # [SH+BR]
    thick: np.ndarray = None      # [SH]
    farea: np.ndarray = None      # [SH]
    slength: np.ndarray = None    # [SH*3]
    xlocal: np.ndarray = None     # [SH*3]
    # frame
    gmod: np.ndarray = None       # [FR]
    istrong: np.ndarray = None
    iweak: np.ndarray = None
    ipolar: np.ndarray = None
    iwarp: np.ndarray = None
    zstrong: np.ndarray = None
    zweak: np.ndarray = None
    auxpt: np.ndarray = None      # [FR*3]
    offset: np.ndarray = None     # [FR*6]
    osflag: np.ndarray = None     # [FR] int32
    mendrel: np.ndarray = None    # [FR*5] int32
    xfr: np.ndarray = None        # [FR*6]
    efFE_ref: np.ndarray = None   # [FR*14]
    # acoustic FSI (ANAFLAG 4): equation split of codes() and what prop_fsi / prop_br leave
    SNDOF: int = 0
    FNDOF: int = 0
    nnorm: np.ndarray = None      # [NJ*3] interface normals
    tarea: np.ndarray = None      # [NJ] tributary interface areas
    fdens: float = 0.0
    # loads
    q: np.ndarray = None          # [NEQ]
    meta: dict = field(default_factory=dict)

    # ---- offsets used all over the reference (shell.c:136-139 etc.) -----------------------
    @property
    def NE_BR(self):
        return self.NE_SBR + self.NE_FBR

    @property
    def n_ef(self):
        return 2 * self.NE_TR + 14 * self.NE_FR + 18 * self.NE_SH

    @property
    def n_c(self):
        return self.NE_TR + 3 * self.NE_FR + 3 * self.NE_SH

    @property
    def n_mcode(self):
        return 6 * self.NE_TR + 14 * self.NE_FR + 18 * self.NE_SH + 24 * self.NE_BR


# ------------------------------------------------------------------------------------------
# codes()  - model.c:937-1143
# ------------------------------------------------------------------------------------------
def initial_jcode(NJ, NE_TR, NE_FR, NE_SH, NE_BR, minc, fixed):
    """jcode before numbering, as struc() leaves it (model.c:215-283).

    ``fixed`` is an iterable of (joint 1-based, direction 1..7).  Returns int64 [NJ,7] holding
    -1 for a free DOF and 0 for a fixed one."""
    jc = np.full((NJ, 7), -1, dtype=I64)
    fixed = np.asarray(list(fixed), dtype=I64).reshape(-1, 2)
    if fixed.size:
        jc[fixed[:, 0] - 1, fixed[:, 1] - 1] = 0
    has_tr = np.zeros(NJ, dtype=bool)
    has_fr = np.zeros(NJ, dtype=bool)
    has_sh = np.zeros(NJ, dtype=bool)
    o = 0
    has_tr[minc[o:o + 2 * NE_TR] - 1] = True
    o += 2 * NE_TR
    has_fr[minc[o:o + 2 * NE_FR] - 1] = True
    o += 2 * NE_FR
    has_sh[minc[o:o + 3 * NE_SH] - 1] = True
    o += 3 * NE_SH
    # struc() counts bricks into the shell column of jflag (model.c:197-205 / App. B.2)
    has_sh[minc[o:o + 8 * NE_BR] - 1] = True
    nofr = ~has_fr
    # no frame, no shell, no truss -> everything fixed ; no frame, no shell, truss -> 4..7 fixed
    floating = nofr & ~has_sh & ~has_tr
    jc[floating, :] = 0
    tr_only = nofr & ~has_sh & has_tr
    jc[tr_only, 3:] = 0
    # no frame but shell/brick -> warping DOF fixed
    jc[nofr & has_sh, 6] = 0
    return jc


def codes(jcode_flags, minc, NE_TR, NE_FR, NE_SH, NE_SBR):
    """Equation numbering + element->DOF maps (model.c:941-1085, ANAFLAG != 4, no released
    warping joints i.e. wrpres[:,0]==0).  Returns (jcode [NJ*7], mcode, NEQ)."""
    jc = np.asarray(jcode_flags, dtype=I64).reshape(-1, 7).copy()
    free = jc != 0
    num = np.cumsum(free.reshape(-1), dtype=I64).reshape(-1, 7)
    jc = np.where(free, num, 0).astype(I64)
    NEQ = int(num[-1, -1]) if jc.size else 0
    parts = []
    o = 0
    if NE_TR:
        e = minc[o:o + 2 * NE_TR].reshape(-1, 2) - 1
        parts.append(jc[e][:, :, :3].reshape(NE_TR, 6).reshape(-1))
    o += 2 * NE_TR
    if NE_FR:
        e = minc[o:o + 2 * NE_FR].reshape(-1, 2) - 1
        parts.append(jc[e].reshape(NE_FR, 14).reshape(-1))
    o += 2 * NE_FR
    if NE_SH:
        e = minc[o:o + 3 * NE_SH].reshape(-1, 3) - 1
        parts.append(jc[e][:, :, :6].reshape(NE_SH, 18).reshape(-1))
    o += 3 * NE_SH
    if NE_SBR:
        e = minc[o:o + 8 * NE_SBR].reshape(-1, 8) - 1
        parts.append(jc[e][:, :, :3].reshape(NE_SBR, 24).reshape(-1))
    mcode = np.concatenate(parts).astype(I64) if parts else np.zeros(0, dtype=I64)
    return jc.reshape(-1), mcode, NEQ


# ------------------------------------------------------------------------------------------
# skylin()  - model.c:1204-1281   (bricks are ignored by the reference, App. B.3)
# ------------------------------------------------------------------------------------------
def skylin(mcode, NEQ, NE_TR, NE_FR, NE_SH):
    kht = np.zeros(NEQ, dtype=I64)
    o = 0
    for ne, nd in ((NE_TR, 6), (NE_FR, 14), (NE_SH, 18)):
        if ne:
            m = mcode[o:o + ne * nd].reshape(ne, nd)
            big = np.where((m > 0) & (m < NEQ), m, NEQ)
            mn = big.min(axis=1, keepdims=True)
            h = np.where(m != 0, m - mn, -1)
            idx = m.reshape(-1)
            hv = h.reshape(-1)
            sel = idx != 0
            np.maximum.at(kht, idx[sel] - 1, hv[sel])
        o += ne * nd
    maxa = np.empty(NEQ + 1, dtype=I64)
    maxa[0] = 1
    np.cumsum(kht + 1, out=maxa[1:])
    maxa[1:] += 1
    lss = int(maxa[NEQ] - 1)
    return kht, maxa, lss


# ------------------------------------------------------------------------------------------
# small vector helpers with the reference's operation order (misc.c:252-282)
# ------------------------------------------------------------------------------------------
def _dot3(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def _cross(a, b, normalise):
    c = np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                  a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                  a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)
    if normalise:
        ln = np.sqrt((c[..., 0] * c[..., 0] + c[..., 1] * c[..., 1]) + c[..., 2] * c[..., 2])
        c = c / ln[..., None]
    return c


def shell_geometry(x, tri):
    """Geometry half of prop_sh (shell.c:63-95) / the shell block of updatc (misc.c:153-184):
    side lengths, face area, triad (c1,c2,c3) and the local membrane coordinates of mem_coord
    (shell.c:2402-2446).  ``tri`` is 0-based [NE,3]."""
    X = x.reshape(-1, 3)
    xj, xk, xl = X[tri[:, 0]], X[tri[:, 1]], X[tri[:, 2]]
    el23 = xl - xk
    el31 = xl - xj
    el12 = xk - xj
    sl = np.stack([np.sqrt(_dot3(el12, el12)), np.sqrt(_dot3(el23, el23)),
                   np.sqrt(_dot3(el31, el31))], axis=-1)
    normal = _cross(el12, el31, False)
    farea = 0.5 * np.sqrt(_dot3(normal, normal))
    lx = el12 / sl[:, 0:1]
    lz = normal / (2 * farea)[:, None]
    ly = _cross(lz, lx, True)
    # mem_coord: x[j][k] = sum_l T[j][l] * X[l][k], X[:,1] = xk-xj, X[:,2] = xl-xj
    Xk = xk - xj
    Xl = xl - xj
    xloc = np.stack([_dot3(lx, Xk), _dot3(lx, Xl), _dot3(ly, Xl)], axis=-1)
    return sl, farea, lx, ly, lz, xloc


def frame_geometry(x, ends, auxpt, offset=None, osflag=None):
    """Geometry half of prop_fr (frame.c:122-158) / frame block of updatc (misc.c:112-147)."""
    X = x.reshape(-1, 3)
    xa = X[ends[:, 0]].copy()
    xb = X[ends[:, 1]].copy()
    if offset is not None and osflag is not None and np.any(osflag):
        m = osflag.astype(bool)
        off = offset.reshape(-1, 6)
        xa[m] = xa[m] + off[m, 0:3]
        xb[m] = xb[m] + off[m, 3:6]
    xfr = np.concatenate([xa, xb], axis=1)
    el = xb - xa
    ll = np.sqrt(_dot3(el, el))
    lx = el / ll[:, None]
    tmp = auxpt.reshape(-1, 3) - xa
    lz = _cross(lx, tmp, True)
    ly = _cross(lz, lx, True)
    return xfr, ll, lx, ly, lz


def truss_geometry(x, ends):
    """prop_tr (truss.c:51-63)."""
    X = x.reshape(-1, 3)
    el = X[ends[:, 1]] - X[ends[:, 0]]
    ll = np.sqrt(_dot3(el, el))
    return ll, el[:, 0] / ll, el[:, 1] / ll, el[:, 2] / ll


# ------------------------------------------------------------------------------------------
# model assembly
# ------------------------------------------------------------------------------------------
def build_model(x, trusses=None, frames=None, shells=None, bricks=None, fixed=(),
                truss_props=None, frame_props=None, shell_props=None, brick_props=None,
                frame_aux=None, loads=(), ANAFLAG=2, ALGFLAG=1, SLVFLAG=0, meta=None,
                jflags=None, want_skyline=None):
    """Assemble a ``Model`` the way main.c:330-1437 does for a deck with these elements.

    trusses/frames [n,2], shells [n,3], bricks [n,8]: 1-based joint numbers.
    truss_props  (E, A, rho, fy)                       truss.c:66
    frame_props  (E, G, rho, A, Iz, Iy, J, Cw, fy, Zz, Zy)   frame.c:62,177
    shell_props  (E, nu, t, rho, fy)                   shell.c:61
    brick_props  (E, nu, rho, fy)                      brick.c:56
    each either one tuple (uniform) or an [n, k] array.
    loads: iterable of (joint, dir 1..7, force)        model.c:1313-1335
    """
    x = np.ascontiguousarray(np.asarray(x, dtype=F64).reshape(-1))
    NJ = x.size // 3

    def arr(a, k):
        if a is None:
            return np.zeros((0, k), dtype=I64)
        return np.ascontiguousarray(np.asarray(a, dtype=I64).reshape(-1, k))

    tr, fr, sh, br = arr(trusses, 2), arr(frames, 2), arr(shells, 3), arr(bricks, 8)
    NE_TR, NE_FR, NE_SH, NE_SBR = len(tr), len(fr), len(sh), len(br)
    minc = np.concatenate([tr.reshape(-1), fr.reshape(-1), sh.reshape(-1), br.reshape(-1)])
    # jflags: the -1/0 jcode of the GLOBAL model when this is an element-partition sub-model
    # (joints without local elements must keep their global equation numbers)
    if jflags is None:
        jflags = initial_jcode(NJ, NE_TR, NE_FR, NE_SH, NE_SBR, minc, fixed)
    jcode, mcode, NEQ = codes(jflags, minc, NE_TR, NE_FR, NE_SH, NE_SBR)
    if want_skyline is None:
        want_skyline = (SLVFLAG == 0)
    if want_skyline:
        kht, maxa, lss = skylin(mcode, NEQ, NE_TR, NE_FR, NE_SH)
    else:
        kht, maxa, lss = None, None, 0

    def props(p, n, k):
        if n == 0:
            return np.zeros((0, k), dtype=F64)
        p = np.asarray(p, dtype=F64)
        if p.ndim == 1:
            p = np.broadcast_to(p, (n, k))
        return np.ascontiguousarray(p)

    ntot = NE_TR + NE_FR + NE_SH + NE_SBR
    m = Model(NJ=NJ, NE_TR=NE_TR, NE_FR=NE_FR, NE_SH=NE_SH, NE_SBR=NE_SBR, NE_FBR=0, NEQ=NEQ,
              ANAFLAG=ANAFLAG, ALGFLAG=ALGFLAG, SLVFLAG=SLVFLAG, x=x, minc=minc, jcode=jcode,
              mcode=mcode, maxa=maxa, kht=kht, lss=lss, meta=dict(meta or {}))
    m.emod = np.zeros(ntot, dtype=F64)
    m.yld = np.zeros(ntot, dtype=F64)
    m.dens = np.zeros(ntot, dtype=F64)
    m.carea = np.zeros(NE_TR + NE_FR, dtype=F64)
    m.llength = np.zeros(NE_TR + NE_FR, dtype=F64)
    nc = NE_TR + 3 * NE_FR + 3 * NE_SH
    m.c1, m.c2, m.c3 = (np.zeros(nc, dtype=F64) for _ in range(3))
    m.nu = np.zeros(NE_SH + NE_SBR, dtype=F64)

    if NE_TR:
        p = props(truss_props, NE_TR, 4)
        m.emod[:NE_TR], m.carea[:NE_TR], m.yld[:NE_TR] = p[:, 0], p[:, 1], p[:, 3]
        m.dens[:NE_TR] = p[:, 2]
        ll, a, b, c = truss_geometry(x, tr - 1)
        m.llength[:NE_TR] = ll
        m.c1[:NE_TR], m.c2[:NE_TR], m.c3[:NE_TR] = a, b, c

    m.gmod = np.zeros(NE_FR); m.istrong = np.zeros(NE_FR); m.iweak = np.zeros(NE_FR)
    m.ipolar = np.zeros(NE_FR); m.iwarp = np.zeros(NE_FR); m.zstrong = np.zeros(NE_FR)
    m.zweak = np.zeros(NE_FR); m.auxpt = np.zeros(NE_FR * 3); m.offset = np.zeros(NE_FR * 6)
    m.osflag = np.zeros(NE_FR, dtype=np.int32); m.mendrel = np.zeros(NE_FR * 5, dtype=np.int32)
    m.xfr = np.zeros(NE_FR * 6); m.efFE_ref = np.zeros(NE_FR * 14)
    if NE_FR:
        p = props(frame_props, NE_FR, 11)
        s = slice(NE_TR, NE_TR + NE_FR)
        m.emod[s], m.gmod[:], m.carea[s] = p[:, 0], p[:, 1], p[:, 3]
        # prop_fr stores density at dens[i] (frame.c:62) - same slot as truss i (App. B.4)
        m.dens[:NE_FR] = p[:, 2]
        m.istrong[:], m.iweak[:], m.ipolar[:], m.iwarp[:] = p[:, 4], p[:, 5], p[:, 6], p[:, 7]
        m.yld[s], m.zstrong[:], m.zweak[:] = p[:, 8], p[:, 9], p[:, 10]
        m.auxpt[:] = np.asarray(frame_aux, dtype=F64).reshape(-1)
        xfr, ll, lx, ly, lz = frame_geometry(x, fr - 1, m.auxpt)
        m.xfr[:] = xfr.reshape(-1)
        m.llength[s] = ll
        cs = slice(NE_TR, NE_TR + 3 * NE_FR)
        m.c1[cs], m.c2[cs], m.c3[cs] = lx.reshape(-1), ly.reshape(-1), lz.reshape(-1)

    m.thick = np.zeros(NE_SH); m.farea = np.zeros(NE_SH)
    m.slength = np.zeros(NE_SH * 3); m.xlocal = np.zeros(NE_SH * 3)
    if NE_SH:
        p = props(shell_props, NE_SH, 5)
        s = slice(NE_TR + NE_FR, NE_TR + NE_FR + NE_SH)
        m.emod[s], m.nu[:NE_SH], m.thick[:], m.yld[s] = p[:, 0], p[:, 1], p[:, 2], p[:, 4]
        m.dens[:NE_SH] = p[:, 3]     # prop_sh: pdens+i (shell.c:61)
        sl, fa, lx, ly, lz, xl = shell_geometry(x, sh - 1)
        m.slength[:], m.farea[:], m.xlocal[:] = sl.reshape(-1), fa, xl.reshape(-1)
        cs = slice(NE_TR + 3 * NE_FR, nc)
        m.c1[cs], m.c2[cs], m.c3[cs] = lx.reshape(-1), ly.reshape(-1), lz.reshape(-1)

    if NE_SBR:
        p = props(brick_props, NE_SBR, 4)
        s = slice(NE_TR + NE_FR + NE_SH, ntot)
        # prop_br writes emod/nu/dens/yield at ptr = NE_TR+NE_FR+NE_SH (brick.c:49,56); nu is only
        # NE_SH+NE_BR long in main.c:594, so models with bricks must have NE_TR = NE_FR = 0.
        if NE_TR or NE_FR:
            raise ValueError("reference overruns nu[] for bricks mixed with truss/frame")
        m.emod[s], m.nu[NE_SH:], m.dens[s], m.yld[s] = p[:, 0], p[:, 1], p[:, 2], p[:, 3]

    m.q = np.zeros(NEQ, dtype=F64)
    jc = jcode.reshape(-1, 7)
    for (jt, dr, val) in loads:
        k = jc[int(jt) - 1, int(dr) - 1]
        if k != 0:
            m.q[k - 1] = val          # load(): *(pq+k-1) = mag (model.c:1328)
    return m


_ARRAY_FIELDS = ["x", "minc", "jcode", "mcode", "maxa", "kht", "emod", "yld", "dens", "carea",
                 "llength", "c1", "c2", "c3", "nu", "thick", "farea", "slength", "xlocal", "gmod",
                 "istrong", "iweak", "ipolar", "iwarp", "zstrong", "zweak", "auxpt", "offset",
                 "osflag", "mendrel", "xfr", "efFE_ref", "q"]
_SCALAR_FIELDS = ["NJ", "NE_TR", "NE_FR", "NE_SH", "NE_SBR", "NE_FBR", "NEQ", "ANAFLAG", "ALGFLAG",
                  "SLVFLAG", "lss"]


def model_to_dict(m, prefix="model_"):
    """flatten a Model into arrays for np.savez (fixtures)"""
    out = {prefix + k: np.asarray(getattr(m, k)) for k in _SCALAR_FIELDS}
    for k in _ARRAY_FIELDS:
        a = getattr(m, k)
        if a is not None:
            out[prefix + k] = a
    return out


def model_from_dict(d, prefix="model_"):
    m = Model()
    for k in _SCALAR_FIELDS:
        setattr(m, k, int(d[prefix + k]))
    for k in _ARRAY_FIELDS:
        if prefix + k in d:
            setattr(m, k, np.ascontiguousarray(d[prefix + k]))
    return m
