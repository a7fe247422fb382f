"""Reader for the reference's positional `model_def.txt` deck (schema: the comment block of
main.c:63-320, read order main.c:340-427, model.c:95-264, main.c:1394-1396, prop_* in
truss.c:66 / frame.c:62-217 / shell.c:61, load() model.c:1313-1335, main.c:1809-1812).

In a real integration the reference's own C parser stays in place; this
reader exists so the tests and bench can push the SHIPPED sample decks through the device path.
Covers truss / frame / shell decks; the static tail (joint loads + NR/MNR controls) is parsed,
the dynamic tail (time histories) is returned unparsed in ``tail``."""
from __future__ import annotations

import numpy as np

from .model import build_model, frame_geometry


def _rows(text):
    for line in text.replace("\r", "").split("\n"):
        line = line.strip()
        if line:
            yield [t.strip() for t in line.split(",")]


def read_deck(text):
    it = _rows(text)
    nxt = lambda: next(it)
    ANAFLAG = int(nxt()[0]); ALGFLAG = int(nxt()[0])
    if ALGFLAG >= 4:
        nxt()                                   # CHKPT,RFLAG   (main.c:352-356)
    SLVFLAG = int(nxt()[0]); OPTFLAG = int(nxt()[0])
    NJ = int(nxt()[0])
    ne = [int(v) for v in nxt()]
    NE_TR, NE_FR, NE_SH = ne[0], ne[1], ne[2]
    NE_BR = sum(ne[3:])
    trusses = [[int(v) for v in nxt()] for _ in range(NE_TR)]
    frames = [[int(v) for v in nxt()] for _ in range(NE_FR)]
    shells = [[int(v) for v in nxt()] for _ in range(NE_SH)]
    bricks = [[int(v) for v in nxt()] for _ in range(NE_BR)]
    fixed = []
    while True:                                 # joint constraints until 0,0 (model.c:222-262)
        r = nxt()
        if int(r[0]) == 0:
            break
        fixed.append((int(r[0]), int(r[1])))
    moved = []
    if ANAFLAG != 4 and ALGFLAG > 3:            # prescribed-displacement DOFs (model.c:1156-1201)
        while True:
            r = nxt()
            if int(r[0]) == 0:
                break
            moved.append((int(r[0]), int(r[1])))
    x = np.array([[float(v) for v in nxt()] for _ in range(NJ)])
    tp = [[float(v) for v in nxt()] for _ in range(NE_TR)]
    fp = None; aux = None; offsets = {}; releases = {}
    if NE_FR:
        a = [[float(v) for v in nxt()] for _ in range(NE_FR)]       # E,G,rho,A,Iz,Iy,J,Cw
        while True:                                                   # member end offsets
            r = nxt()
            if int(float(r[0])) == 0:
                break
            offsets[int(r[0])] = [float(v) for v in r[1:7]]
        aux = [[float(v) for v in nxt()] for _ in range(NE_FR)]
        y = [[float(v) for v in nxt()] for _ in range(NE_FR)]        # fy,Zz,Zy
        while True:                                                   # member end releases
            r = nxt()
            if int(r[0]) == 0:
                break
            releases[int(r[0])] = [int(v) for v in r[1:5]]
        fp = [ai + yi for ai, yi in zip(a, y)]
    sp = [[float(v) for v in nxt()] for _ in range(NE_SH)]
    loads = []; params = None; tail = None
    if ALGFLAG < 4:
        while True:                             # joint loads until 0,0,0 (model.c:1313-1335)
            r = nxt()
            if int(r[0]) == 0:
                break
            loads.append((int(r[0]), int(r[1]), float(r[2])))
        if ANAFLAG == 1:
            params = dict(lpfmax=float(nxt()[0]))
        elif ALGFLAG in (1, 2):
            a = [float(v) for v in nxt()]; b = [int(v) for v in nxt()]; c = [float(v) for v in nxt()]
            params = dict(lpfmax=a[0], lpf=a[1], dlpf=a[2], dlpfmax=a[3], dlpfmin=a[4], itemax=b[0],
                          submax=b[1], solmin=b[2], toldisp=c[0], tolforc=c[1], tolener=c[2],
                          algflag=ALGFLAG)
    else:
        tail = [r for r in it]
    dyn_tail = tail
    m = build_model(x, trusses=trusses or None, frames=frames or None, shells=shells or None,
                    bricks=bricks or None, fixed=fixed, truss_props=np.array(tp) if tp else None,
                    frame_props=np.array(fp) if fp else None, shell_props=np.array(sp) if sp else None,
                    frame_aux=np.array(aux) if aux else None, loads=loads, ANAFLAG=ANAFLAG,
                    ALGFLAG=ALGFLAG, SLVFLAG=SLVFLAG, want_skyline=True,
                    meta=dict(kind="deck", OPTFLAG=OPTFLAG))
    if offsets or releases:
        for k, v in offsets.items():
            m.osflag[k - 1] = 1; m.offset[(k - 1) * 6:(k - 1) * 6 + 6] = v
        for k, v in releases.items():
            m.mendrel[(k - 1) * 5] = 1; m.mendrel[(k - 1) * 5 + 1:(k - 1) * 5 + 5] = v
        if offsets:
            fr = np.asarray(frames) - 1
            xfr, ll, lx, ly, lz = frame_geometry(m.x, fr, m.auxpt, m.offset, m.osflag)
            s = slice(m.NE_TR, m.NE_TR + m.NE_FR); cs = slice(m.NE_TR, m.NE_TR + 3 * m.NE_FR)
            m.xfr[:] = xfr.reshape(-1); m.llength[s] = ll
            m.c1[cs], m.c2[cs], m.c3[cs] = lx.reshape(-1), ly.reshape(-1), lz.reshape(-1)
    if dyn_tail is not None:
        tail = parse_dynamic_tail(m, dyn_tail, ALGFLAG)
        pmot = np.zeros(m.NEQ, dtype=np.int32)              # skylin(), model.c:1149-1201
        jc = m.jcode.reshape(-1, 7)
        for j, k in moved:
            pmot[jc[j - 1, k - 1] - 1] = 1
        tail["pmot"] = pmot
        tail["nbc"] = int(pmot.sum())
    return m, params, tail


def parse_dynamic_tail(m, rows, ALGFLAG):
    """the transient part of the deck: `ntstpsinpt, ttot` (main.c:1604-1608), the load histories,
    prescribed support motion and initial conditions that load() reads (model.c:1490-1606), the
    Newmark options (main.c:3160-3180 / 3319-3351) and, for ALGFLAG 5, the Newton controls."""
    it = iter(rows)
    r = next(it)
    nin = int(r[0]); ttot = float(r[1])
    dt = ttot / nin
    nt = nin + 1
    jc = m.jcode.reshape(-1, 7)
    pinpt = np.zeros((m.NEQ, nt)); pdisp = np.zeros((m.NEQ, nt))

    def series(dst, zero_to_tiny):
        kps, i = 0, 0
        while True:
            r = next(it)
            jt, dr, mag = int(r[0]), int(r[1]), float(r[2])
            if jt == 0:
                return
            if zero_to_tiny and mag == 0:
                mag = 1e-30                              # model.c:1517-1521, 1570-1572
            ks = int(jc[jt - 1, dr - 1])
            if ks != kps:
                i = 0
            if ks != 0:
                dst[ks - 1, i] = mag
            kps = ks; i += 1
    series(pinpt, ALGFLAG == 5)
    series(pdisp, True)
    um = np.zeros(m.NEQ); vm = np.zeros(m.NEQ); am = np.zeros(m.NEQ)
    while True:
        r = next(it)
        jt = int(r[0])
        if jt == 0:
            break
        k = int(jc[jt - 1, int(r[1]) - 1])
        if k:
            dv, vv, av = float(r[2]), float(r[3]), float(r[4])
            if dv != 0: um[k - 1] = dv
            if vv != 0: vm[k - 1] = vv
            if av != 0: am[k - 1] = av
    r = next(it)
    numopt, rho = float(r[0]), float(r[1])
    if numopt == 0: alpham, alphaf = 0.0, 0.0                      # main.c:3337-3351
    elif numopt == 1: alpham, alphaf = (2 * rho - 1) / (rho + 1), rho / (rho + 1)
    elif numopt == 2: alpham, alphaf = 0.0, (1 - rho) / (1 + rho)
    else: alpham, alphaf = (rho - 1) / (rho + 1), 0.0
    out = dict(ntstps=nt, dt=dt, pinpt=pinpt, pdisp=pdisp, um=um, vm=vm, am=am, alpham=alpham,
               alphaf=alphaf, nbc=int((pdisp != 0).any()))
    if ALGFLAG == 5:
        a = [float(v) for v in next(it)]; b = [int(v) for v in next(it)]; c = [float(v) for v in next(it)]
        out["params"] = dict(lpfmax=a[0], lpf=a[1], dlpf=a[2], dlpfmax=a[3], dlpfmin=a[4], itemax=b[0],
                             submax=b[1], solmin=b[2], toldisp=c[0], tolforc=c[1], tolener=c[2])
    return out


def write_shell_deck(m, props, loads, tail_lines, ALGFLAG=None):
    """text of a `model_def.txt` for a shell-only static model (order fixed by main.c:340-427,
    model.c:95-264, main.c:1394-1396, shell.c:61, model.c:1313-1335): used to run generated meshes
    through the unmodified reference driver.  ``tail_lines``: the solver-control lines that follow
    the joint loads (NR/MNR: main.c:1809-1812; arc-length: arc.c:48 + main.c:2240-2245)."""
    assert m.NE_TR == 0 and m.NE_FR == 0 and m.NE_SBR + m.NE_FBR == 0
    L = [str(m.ANAFLAG), str(m.ALGFLAG if ALGFLAG is None else ALGFLAG), "0", "1", str(m.NJ),
         f"0,0,{m.NE_SH},0,0"]
    L += [",".join(str(int(v)) for v in r) for r in m.minc.reshape(-1, 3)]
    jc = m.jcode.reshape(-1, 7)
    for j in range(m.NJ):
        for r in range(6):
            if jc[j, r] == 0:
                L.append(f"{j + 1},{r + 1}")
    L.append("0,0")
    L += ["%.17g,%.17g,%.17g" % tuple(r) for r in m.x.reshape(-1, 3)]
    L += ["%.17g,%.17g,%.17g,%.17g,%.17g" % tuple(props)] * m.NE_SH
    L += ["%d,%d,%.17g" % (j, r, v) for (j, r, v) in loads] + ["0,0,0"]
    L += list(tail_lines)
    return "\n".join(L) + "\n"
