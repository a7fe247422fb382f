"""Synthetic meshes of the BASELINE.json shapes.

plate_model    unit-square DKT plate, nx*ny cells, 2 triangles per cell (a,b,c),(a,c,d),
               joints numbered ``nid(i,j) = i*(ny+1)+j+1`` (row-major along the short side),
               all four edges pinned (DOF 1-3), centre point load -1000 in z; material of
               Sample_Input_Files/model_def_5c_shell.txt:33 unless overridden.
lattice_model  cubic n*n*n frame lattice, members along x, y, z, base plane clamped;
               properties of Sample_Input_Files/model_def_5b_frame.txt:36.
truss_model    the same lattice topology with 2-node trusses (pinned base).
brick_model    nx*ny*nz 8-node bricks, base clamped (linear: stiff_br only).
perturbation   the seeded dd ~ U(-1e-4,1e-4) of section 8(d), rng 20261017.
"""
from __future__ import annotations

import numpy as np
from .model import build_model, I64, F64

SHELL_5C = (2.1e11, 0.3, 0.01, 8050.0, 3.45e8)      # E, nu, t, rho, fy
# model_def_5b_frame.txt:36  E,G,rho,A,Iz,Iy,J,Cw ; :57 fy,Zz,Zy
FRAME_5B = (29000.0, 11200.0, 7.34e-7, 9.13, 110.0, 37.1, 0.536, 0.0, 36.0, 24.7, 11.3)
TRUSS_5A = (29000.0, 1.0, 7.34e-7, 36.0)            # E, A, rho, fy
BRICK_DEF = (2.1e11, 0.3, 8050.0, 3.45e8)           # E, nu, rho, fy


def plate_model(nx, ny, lx=1.0, ly=1.0, props=SHELL_5C, load=-1000.0, ANAFLAG=2, ALGFLAG=1,
                SLVFLAG=0, pinned=True, z_bump=0.0, jitter=0.0, jitter_seed=7, unionjack=False):
    """``jitter``: interior joints moved in-plane by up to that fraction of a cell (seeded) - an
    unstructured variant of the same plate in which no two shells share their geometry.
    ``unionjack``: the cell diagonals alternate, so joints have 4 or 8 shells around them instead of 6
    (joint-pair blocks of 4 and 8 contributions)"""
    i, j = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
    xs = (i * (lx / nx)).astype(F64)
    ys = (j * (ly / ny)).astype(F64)
    if jitter:
        rng = np.random.default_rng(jitter_seed)
        inner = (i > 0) & (i < nx) & (j > 0) & (j < ny)
        xs = xs + inner * rng.uniform(-jitter, jitter, xs.shape) * (lx / nx)
        ys = ys + inner * rng.uniform(-jitter, jitter, ys.shape) * (ly / ny)
    zs = np.zeros_like(xs)
    if z_bump:
        zs = z_bump * np.sin(np.pi * xs / lx) * np.sin(np.pi * ys / ly)
    x = np.stack([xs, ys, zs], axis=-1).reshape(-1)
    nid = (i * (ny + 1) + j + 1).astype(I64)
    a = nid[:-1, :-1]; b = nid[1:, :-1]; c = nid[1:, 1:]; d = nid[:-1, 1:]
    t1 = np.stack([a, b, c], axis=-1)
    t2 = np.stack([a, c, d], axis=-1)
    if unionjack:
        odd = ((i[:-1, :-1] + j[:-1, :-1]) % 2 == 1)[..., None]
        t1 = np.where(odd, np.stack([a, b, d], axis=-1), t1)
        t2 = np.where(odd, np.stack([b, c, d], axis=-1), t2)
    shells = np.stack([t1, t2], axis=2).reshape(-1, 3)
    fixed = []
    if pinned:
        edge = (i == 0) | (i == nx) | (j == 0) | (j == ny)
        ej = nid[edge]
        fixed = np.stack([np.repeat(ej, 3), np.tile(np.array([1, 2, 3], dtype=I64), ej.size)],
                         axis=1)
    centre = int(nid[nx // 2, ny // 2])
    return build_model(x, shells=shells, fixed=fixed, shell_props=props,
                       loads=[(centre, 3, load)], ANAFLAG=ANAFLAG, ALGFLAG=ALGFLAG,
                       SLVFLAG=SLVFLAG, meta=dict(kind="plate", nx=nx, ny=ny, centre=centre))


def _lattice(n, h=1.0):
    g = np.arange(n)
    i, j, k = np.meshgrid(g, g, g, indexing="ij")
    nid = (i * n * n + j * n + k + 1).astype(I64)
    x = np.stack([i * h, j * h, k * h], axis=-1).astype(F64).reshape(-1)
    mx = np.stack([nid[:-1], nid[1:]], axis=-1).reshape(-1, 2)          # along x
    my = np.stack([nid[:, :-1], nid[:, 1:]], axis=-1).reshape(-1, 2)    # along y
    mz = np.stack([nid[:, :, :-1], nid[:, :, 1:]], axis=-1).reshape(-1, 2)  # along z
    return x, nid, mx, my, mz


def lattice_model(n, h=100.0, props=FRAME_5B, ANAFLAG=2, ALGFLAG=1, SLVFLAG=0, load=1.0):
    x, nid, mx, my, mz = _lattice(n, h)
    frames = np.concatenate([mx, my, mz])
    X = x.reshape(-1, 3)
    aux = X[frames[:, 0] - 1].copy()
    aux[:len(mx)] += np.array([0.0, 1.0, 0.0]) * h
    aux[len(mx):len(mx) + len(my)] += np.array([1.0, 0.0, 0.0]) * h
    aux[len(mx) + len(my):] += np.array([0.0, 1.0, 0.0]) * h
    base = nid[:, :, 0].reshape(-1)
    fixed = np.stack([np.repeat(base, 7), np.tile(np.arange(1, 8, dtype=I64), base.size)], axis=1)
    top = int(nid[n // 2, n // 2, n - 1])
    return build_model(x, frames=frames, fixed=fixed, frame_props=props, frame_aux=aux,
                       loads=[(top, 1, load)], ANAFLAG=ANAFLAG, ALGFLAG=ALGFLAG, SLVFLAG=SLVFLAG,
                       meta=dict(kind="lattice", n=n, top=top))


def truss_model(n, h=100.0, props=TRUSS_5A, ANAFLAG=2, ALGFLAG=1, SLVFLAG=0, load=1.0):
    x, nid, mx, my, mz = _lattice(n, h)
    # face diagonals in the xz and yz planes make the pin-jointed lattice stable
    dxz = np.stack([nid[:-1, :, :-1], nid[1:, :, 1:]], axis=-1).reshape(-1, 2)
    dyz = np.stack([nid[:, :-1, :-1], nid[:, 1:, 1:]], axis=-1).reshape(-1, 2)
    dxy = np.stack([nid[:-1, :-1, :], nid[1:, 1:, :]], axis=-1).reshape(-1, 2)
    trusses = np.concatenate([mx, my, mz, dxz, dyz, dxy])
    base = nid[:, :, 0].reshape(-1)
    fixed = np.stack([np.repeat(base, 3), np.tile(np.arange(1, 4, dtype=I64), base.size)], axis=1)
    top = int(nid[n // 2, n // 2, n - 1])
    return build_model(x, trusses=trusses, fixed=fixed, truss_props=props,
                       loads=[(top, 1, load)], ANAFLAG=ANAFLAG, ALGFLAG=ALGFLAG, SLVFLAG=SLVFLAG,
                       meta=dict(kind="truss", n=n, top=top))


def brick_model(nx, ny, nz, h=1.0, props=BRICK_DEF, skin=False, shell_props=SHELL_5C,
                SLVFLAG=2, distort=0.0, seed=7):
    """nx*ny*nz hexahedra; node order per brick.c:199-315: local node n has natural
    coordinates (r,s,t) signs (+,+,+),(-,+,+),(-,-,+),(+,-,+),(+,+,-),(-,+,-),(-,-,-),(+,-,-).
    ``skin`` adds DKT shells on the top face sharing the brick joints (config 5 shape)."""
    i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    nid = (i * (ny + 1) * (nz + 1) + j * (nz + 1) + k + 1).astype(I64)
    x = np.stack([i * h, j * h, k * h], axis=-1).astype(F64)
    if distort:
        rng = np.random.default_rng(seed)
        x = x + distort * h * rng.uniform(-1, 1, size=x.shape)
    x = x.reshape(-1)
    lo, hi = slice(0, -1), slice(1, None)
    sel = {(+1): hi, (-1): lo}
    signs = [(1, 1, 1), (-1, 1, 1), (-1, -1, 1), (1, -1, 1), (1, 1, -1), (-1, 1, -1),
             (-1, -1, -1), (1, -1, -1)]
    bricks = np.stack([nid[sel[a], sel[b], sel[c]] for (a, b, c) in signs], axis=-1).reshape(-1, 8)
    base = nid[:, :, 0].reshape(-1)
    fixed = np.stack([np.repeat(base, 3), np.tile(np.arange(1, 4, dtype=I64), base.size)], axis=1)
    shells = None
    if skin:
        top = nid[:, :, nz]
        a = top[:-1, :-1]; b = top[1:, :-1]; c = top[1:, 1:]; d = top[:-1, 1:]
        shells = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], axis=2).reshape(-1, 3)
    corner = int(nid[nx, ny, nz])
    return build_model(x, shells=shells, bricks=bricks, fixed=fixed, brick_props=props,
                       shell_props=shell_props, loads=[(corner, 3, -1000.0)], ANAFLAG=1,
                       ALGFLAG=1, SLVFLAG=SLVFLAG,
                       meta=dict(kind="brick", nx=nx, ny=ny, nz=nz))


def fsi_model(nx, ny, nzs, nzf, h=1.0, props=BRICK_DEF, bulk=2.2e9, fdens=1000.0, skin=False, shell_props=SHELL_5C,
              distort=0.0, seed=7):
    """Acoustic fluid-structure model (ANAFLAG 4, fsi.c; BASELINE.json configs[4] shape): a block of nx*ny*nzf FLUID
    bricks below z = 0 whose top face is wet.  The structure on it is either nx*ny*nzs solid bricks above z = 0
    (brFSI_FLAG) or, with ``skin``, DKT shells on the plane z = 0 (shFSI_FLAG).  Numbering as codes() does for
    ANAFLAG 4 (model.c:962-990): structural equations first, joint by joint, then the pressure equations (jcode
    slot 7); fluid brick mcode keeps the pressure in the z slot of every joint (model.c:1089-1140); fluid brick
    properties as prop_br leaves them (brick.c:60-75): emod 1e20, nu 0.5e20, dens 1 / c^2.  nnorm / tarea (what
    prop_fsi computes from the deck, fsi.c:44-331) are the outward normal (0, 0, -1) of the structure's wet face
    and the tributary area of every interface joint."""
    from .model import Model
    nz = nzs + nzf if not skin else nzf
    i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    nid = (i * (ny + 1) * (nz + 1) + j * (nz + 1) + k + 1).astype(I64)
    X = np.stack([i * h, j * h, (k - nzf) * h], axis=-1).astype(F64)
    if distort:
        rng = np.random.default_rng(seed)
        X = X + distort * h * rng.uniform(-1, 1, size=X.shape) * (k[..., None] != nzf)      # the interface stays plane
    x = X.reshape(-1)
    NJ = x.size // 3
    lo, hi = slice(0, -1), slice(1, None)
    sel = {(+1): hi, (-1): lo}
    signs = [(1, 1, 1), (-1, 1, 1), (-1, -1, 1), (1, -1, 1), (1, 1, -1), (-1, 1, -1), (-1, -1, -1), (1, -1, -1)]
    allb = np.stack([nid[sel[a], sel[b], sel[c]] for (a, b, c) in signs], axis=-1)          # [nx, ny, nz, 8]
    fluid = allb[:, :, :nzf].reshape(-1, 8)
    solid = allb[:, :, nzf:].reshape(-1, 8) if not skin else np.zeros((0, 8), dtype=I64)
    shells = np.zeros((0, 3), dtype=I64)
    if skin:
        top = nid[:, :, nzf]
        a = top[:-1, :-1]; b = top[1:, :-1]; c = top[1:, 1:]; d = top[:-1, 1:]
        shells = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], axis=2).reshape(-1, 3)
    NE_SH, NE_SBR, NE_FBR = len(shells), len(solid), len(fluid)
    # degrees of freedom: translations (+ rotations of the skin) on the structure, pressure on the fluid joints
    kk = k.reshape(-1)
    on_struct = kk >= nzf if not skin else kk == nzf
    on_fluid = kk <= nzf
    jc = np.zeros((NJ, 7), dtype=I64)
    jc[on_struct, :3] = -1
    if skin:
        jc[on_struct, 3:6] = -1
        edge = ((i == 0) | (i == nx) | (j == 0) | (j == ny)).reshape(-1)
        jc[on_struct & edge, 3:6] = 0       # guided edges: L_br (fsi.c:514) needs every wet translation free
    else:
        jc[kk == nz, :3] = 0                                # clamped top face
    jc[on_fluid, 6] = -1
    free = jc[:, :6] != 0
    num = np.cumsum(free.reshape(-1), dtype=I64).reshape(NJ, 6)
    jcode = np.zeros((NJ, 7), dtype=I64)
    jcode[:, :6] = np.where(free, num, 0)
    SNDOF = int(num[-1, -1])
    pf = jc[:, 6] != 0
    jcode[pf, 6] = SNDOF + np.cumsum(pf, dtype=I64)[pf]
    NEQ = int(jcode.max()); FNDOF = NEQ - SNDOF
    minc = np.concatenate([shells.reshape(-1), solid.reshape(-1), fluid.reshape(-1)])
    mc = [jcode[shells - 1][:, :, :6].reshape(-1)] if NE_SH else []
    if NE_SBR:
        mc.append(jcode[solid - 1][:, :, :3].reshape(-1))
    fm = np.zeros((NE_FBR, 8, 3), dtype=I64); fm[:, :, 2] = jcode[fluid - 1][:, :, 6]
    mc.append(fm.reshape(-1))
    mcode = np.concatenate(mc).astype(I64)
    m = Model(NJ=NJ, NE_SH=NE_SH, NE_SBR=NE_SBR, NE_FBR=NE_FBR, NEQ=NEQ, ANAFLAG=4, ALGFLAG=4, SLVFLAG=2, x=x,
              minc=minc, jcode=jcode.reshape(-1), mcode=mcode, lss=0,
              meta=dict(kind="fsi", nx=nx, ny=ny, nzs=nzs, nzf=nzf, skin=skin))
    m.SNDOF, m.FNDOF, m.fdens = SNDOF, FNDOF, float(fdens)
    ntot = NE_SH + NE_SBR + NE_FBR
    m.emod = np.zeros(ntot); m.yld = np.zeros(ntot); m.dens = np.zeros(ntot); m.nu = np.zeros(ntot)
    m.carea = np.zeros(0); m.llength = np.zeros(0)
    nc = 3 * NE_SH
    m.c1, m.c2, m.c3 = (np.zeros(nc) for _ in range(3))
    m.thick = np.zeros(NE_SH); m.farea = np.zeros(NE_SH); m.slength = np.zeros(NE_SH * 3); m.xlocal = np.zeros(NE_SH * 3)
    if NE_SH:
        from .model import shell_geometry
        E, nu, t, rho, fy = shell_props
        m.emod[:NE_SH], m.nu[:NE_SH], m.thick[:], m.yld[:NE_SH], m.dens[:NE_SH] = E, nu, t, fy, rho
        sl, fa, lx, ly, lz, xl = shell_geometry(x, shells - 1)
        m.slength[:], m.farea[:], m.xlocal[:] = sl.reshape(-1), fa, xl.reshape(-1)
        m.c1[:], m.c2[:], m.c3[:] = lx.reshape(-1), ly.reshape(-1), lz.reshape(-1)
    if NE_SBR:
        E, nu, rho, fy = props
        s = slice(NE_SH, NE_SH + NE_SBR)
        m.emod[s], m.nu[s], m.dens[s], m.yld[s] = E, nu, rho, fy
    s = slice(NE_SH + NE_SBR, ntot)
    wvsp = np.sqrt(bulk / fdens)                            # brick.c:63
    m.emod[s], m.nu[s], m.dens[s] = 1e20, .5e20, 1 / wvsp ** 2
    for nm in ("gmod", "istrong", "iweak", "ipolar", "iwarp", "zstrong", "zweak", "auxpt", "offset", "xfr", "efFE_ref"):
        setattr(m, nm, np.zeros(0))
    m.osflag = np.zeros(0, dtype=np.int32); m.mendrel = np.zeros(0, dtype=np.int32)
    # interface data: normal of the structure's wet face, tributary areas (corner 1/4, edge 1/2, interior 1 cell)
    m.nnorm = np.zeros(NJ * 3); m.tarea = np.zeros(NJ)
    wet = (kk == nzf)
    m.nnorm.reshape(-1, 3)[wet] = (0.0, 0.0, -1.0)
    ii, jj = i.reshape(-1), j.reshape(-1)
    wx = np.where((ii == 0) | (ii == nx), 0.5, 1.0); wy = np.where((jj == 0) | (jj == ny), 0.5, 1.0)
    m.tarea[wet] = (wx * wy * h * h)[wet]
    m.q = np.zeros(NEQ)
    return m


def perturbation(model, scale=1e-4, seed=20261017):
    """Seeded incremental displacement of SURVEY.md section 8(d): dd ~ U(-scale, scale) per
    free DOF, numpy default_rng(20261017)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-scale, scale, size=model.NEQ).astype(F64)
