"""TEST INFRASTRUCTURE ONLY - ctypes binding to oracle/_ref/libcubens_ref.so, i.e. the
UNMODIFIED reference element / assembly routines (prototypes.h) called in memory.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (the product path is cu-bens_b200/ and fails loudly without its CUDA
library).  The library is built by oracle/Makefile from the sources where they lie under
/root/reference; on the GPU box the prebuilt .so travels with the snapshot.

RefState mirrors the generation bookkeeping of main.c (committed / _temp,_i / _ip arrays,
main.c:1833-1882, 1982-1984, 2006-2028, 2074-2134) with plain numpy copies so a test can walk
the reference through exactly the sequence of calls the Newton loop makes.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# CUBENS_REF_O0=1 (set before the first call): the build at the reference's shipped optimisation level (no -O)
REF_SO = os.path.join(HERE, "_ref", "libcubens_ref_O0.so" if os.environ.get("CUBENS_REF_O0") else "libcubens_ref.so")

_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(REF_SO)
        _lib.ref_get_NEQ.restype = C.c_long
        _lib.ref_get_NBC.restype = C.c_long
        _lib.ref_csc_nnz.restype = C.c_long
        _lib.dot.restype = C.c_double
        _lib.ref_open_sinks()
    return _lib


def P(a):
    if a is None:
        return C.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def L(v):
    return C.c_long(int(v))


def set_model(m, SLVFLAG=None, ANAFLAG=None):
    """Point the reference's globals (main.c:323-328) at this model."""
    l = lib()
    l.ref_set_sizes(L(m.NJ), L(m.NE_TR), L(m.NE_FR), L(m.NE_SH), L(m.NE_SBR), L(m.NE_FBR),
                    L(m.NEQ))
    l.ref_set_flags(C.c_int(m.ANAFLAG if ANAFLAG is None else ANAFLAG), C.c_int(m.ALGFLAG),
                    C.c_int(m.SLVFLAG if SLVFLAG is None else SLVFLAG), C.c_int(1))
    return l


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class RefState:
    """All mutable arrays of the hot loop, three generations."""

    def __init__(self, m):
        z = np.zeros
        self.m = m
        self.x = m.x.copy(); self.x_temp = m.x.copy(); self.x_ip = m.x.copy()
        for nm in ("c1", "c2", "c3"):
            base = getattr(m, nm)
            setattr(self, nm, base.copy()); setattr(self, nm + "_i", base.copy())
            setattr(self, nm + "_ip", base.copy())
        nef = m.n_ef
        self.ef = z(nef); self.ef_i = z(nef); self.ef_ip = z(nef)
        nl = m.NE_TR + m.NE_FR
        self.llength = m.llength.copy()
        self.defllen = m.llength.copy(); self.defllen_i = m.llength.copy()
        self.defllen_ip = m.llength.copy()
        self.farea = m.farea.copy(); self.slength = m.slength.copy()
        self.deffarea = m.farea.copy(); self.deffarea_i = m.farea.copy()
        self.deffarea_ip = m.farea.copy()
        self.defslen = m.slength.copy(); self.defslen_i = m.slength.copy()
        self.defslen_ip = m.slength.copy()
        self.chi = z(m.NE_SH * 3); self.chi_temp = z(m.NE_SH * 3)
        self.efN = z(m.NE_SH * 9); self.efN_temp = z(m.NE_SH * 9)
        self.efM = z(m.NE_SH * 9); self.efM_temp = z(m.NE_SH * 9)
        self.xfr = m.xfr.copy(); self.xfr_temp = m.xfr.copy()
        self.efFE = z(m.NE_FR * 14); self.efFE_i = z(m.NE_FR * 14); self.efFE_ip = z(m.NE_FR * 14)
        self.yldflag = np.zeros(m.NE_FR * 2, dtype=np.int32)
        self.d = z(m.NEQ); self.d_temp = z(m.NEQ); self.f = z(m.NEQ); self.f_temp = z(m.NEQ)
        self.Jinv = z(9 * 8 * max(m.NE_BR, 1)); self.jac = z(9)

    # main.c:1833-1882
    def begin_increment(self):
        self.d_temp[:] = self.d; self.f_temp[:] = self.f
        self.x_temp[:] = self.x
        self.ef_i[:] = self.ef; self.ef_ip[:] = self.ef
        for nm in ("c1", "c2", "c3"):
            getattr(self, nm + "_i")[:] = getattr(self, nm)
            getattr(self, nm + "_ip")[:] = getattr(self, nm)
        self.defllen_i[:] = self.defllen; self.defllen_ip[:] = self.defllen
        self.xfr_temp[:] = self.xfr
        self.efFE_i[:] = self.efFE; self.efFE_ip[:] = self.efFE
        self.deffarea_i[:] = self.deffarea; self.deffarea_ip[:] = self.deffarea
        self.defslen_i[:] = self.defslen; self.defslen_ip[:] = self.defslen
        self.chi_temp[:] = self.chi; self.efN_temp[:] = self.efN; self.efM_temp[:] = self.efM

    # main.c:2006-2028
    def end_iteration(self):
        for nm in ("c1", "c2", "c3"):
            getattr(self, nm + "_ip")[:] = getattr(self, nm + "_i")
        self.defllen_ip[:] = self.defllen_i
        self.efFE_ip[:] = self.efFE_i
        self.deffarea_ip[:] = self.deffarea_i
        self.defslen_ip[:] = self.defslen_i

    # main.c:2074-2134
    def commit(self):
        self.d[:] = self.d_temp; self.f[:] = self.f_temp
        self.ef[:] = self.ef_i; self.efFE[:] = self.efFE_i
        self.x[:] = self.x_temp
        for nm in ("c1", "c2", "c3"):
            getattr(self, nm)[:] = getattr(self, nm + "_i")
        self.defllen[:] = self.defllen_i; self.xfr[:] = self.xfr_temp
        self.yldflag[self.yldflag == 2] = 0
        self.deffarea[:] = self.deffarea_i; self.defslen[:] = self.defslen_i
        self.chi[:] = self.chi_temp; self.efN[:] = self.efN_temp; self.efM[:] = self.efM_temp


def stiff(m, s, SLVFLAG=None, gen="ip"):
    """ss <- 0 ; stiff_tr ; stiff_fr ; stiff_sh [; stiff_br]   (main.c:1899-1921 / 1729-1754).
    gen="ip" uses the *_ip generation + x_temp (Newton loop); gen="c" the committed one."""
    slv = m.SLVFLAG if SLVFLAG is None else SLVFLAG
    l = set_model(m, SLVFLAG=slv)
    n = m.lss if slv == 0 else m.NEQ * m.NEQ
    ss = np.zeros(n)
    if gen == "ip":
        c1, c2, c3, ef, dl, dfa, dsl, efFE, x = (s.c1_ip, s.c2_ip, s.c3_ip, s.ef_ip,
                                                 s.defllen_ip, s.deffarea_ip, s.defslen_ip,
                                                 s.efFE_ip, s.x_temp)
        chi, efN, efM, d = s.chi_temp, s.efN_temp, s.efM_temp, s.d_temp
    else:
        c1, c2, c3, ef, dl, dfa, dsl, efFE, x = (s.c1, s.c2, s.c3, s.ef, s.defllen, s.deffarea,
                                                 s.defslen, s.efFE, s.x)
        chi, efN, efM, d = s.chi, s.efN, s.efM, s.d
    if m.NE_TR:
        l.stiff_tr(P(ss), P(m.emod), P(m.carea), P(s.llength), P(dl), P(m.yld), P(c1), P(c2),
                   P(c3), P(ef), P(m.maxa), P(m.mcode))
    if m.NE_FR:
        l.stiff_fr(P(ss), P(m.emod), P(m.gmod), P(m.carea), P(m.offset), P(m.osflag),
                   P(s.llength), P(dl), P(m.istrong), P(m.iweak), P(m.ipolar), P(m.iwarp),
                   P(s.yldflag), P(m.yld), P(m.zstrong), P(m.zweak), P(c1), P(c2), P(c3), P(ef),
                   P(efFE), P(m.mendrel), P(m.maxa), P(m.mcode))
    if m.NE_SH:
        l.stiff_sh(P(ss), P(m.emod), P(m.nu), P(x), P(m.xlocal), P(m.thick), P(s.farea), P(dfa),
                   P(s.slength), P(dsl), P(m.yld), P(c1), P(c2), P(c3), P(ef), P(d), P(chi),
                   P(efN), P(efM), P(m.maxa), P(m.minc), P(m.mcode))
    if m.NE_BR:
        assert slv != 0, "stiff_br only scatters to the dense layout (brick.c:383-395)"
        l.stiff_br(P(ss), P(x), P(m.emod), P(m.nu), P(m.minc), P(m.mcode), P(m.jcode),
                   P(s.Jinv), P(s.jac))
    return ss


def mass(m, s, SLVFLAG=None):
    """sm <- 0 ; mass_tr ; mass_fr ; mass_sh [; mass_br]  (main.c:3590-3619).  NOTE: the
    reference routines overwrite llength / xfr / slength / farea from the committed x
     - RefState carries those arrays so the side effect is visible."""
    slv = m.SLVFLAG if SLVFLAG is None else SLVFLAG
    l = set_model(m, SLVFLAG=slv)
    n = m.NEQ if slv == 0 else m.NEQ * m.NEQ
    sm = np.zeros(max(n, m.lss if slv == 0 else n))
    if m.NE_TR:
        l.mass_tr(P(sm), P(m.carea), P(s.llength), P(m.dens), P(s.x), P(m.minc), P(m.mcode),
                  P(s.jac))
    if m.NE_FR:
        l.mass_fr(P(sm), P(m.carea), P(s.llength), P(m.istrong), P(m.iweak), P(m.ipolar),
                  P(m.iwarp), P(m.dens), P(m.osflag), P(m.offset), P(s.x), P(s.xfr), P(m.minc),
                  P(m.mcode), P(s.jac))
    if m.NE_SH:
        l.mass_sh(P(sm), P(m.carea), P(m.dens), P(m.thick), P(s.farea), P(s.slength), P(s.x),
                  P(m.minc), P(m.mcode), P(s.jac))
    if m.NE_BR:
        assert slv != 0
        l.mass_br(P(sm), P(m.dens), P(s.x), P(m.minc), P(m.mcode), P(s.jac))
    return sm[:n]


def update_forces(m, s, dd, dlpf=1.0, itecnt=0):
    """d_temp += dd ; f_temp <- 0 ; updatc ; forces_tr/fr/sh ; ef_ip <- ef_i
    (main.c:1948-1984).  Returns (frcchk_fr, frcchk_sh, dlpf)."""
    l = set_model(m)
    dd = np.ascontiguousarray(dd, dtype=np.float64)
    s.d_temp += dd
    s.f_temp[:] = 0
    l.updatc(P(s.x_temp), P(s.x_ip), P(s.xfr_temp), P(dd), P(s.defllen_i), P(s.deffarea_i),
             P(s.defslen_i), P(m.offset), P(m.osflag), P(m.auxpt), P(s.c1_i), P(s.c2_i),
             P(s.c3_i), P(m.minc), P(m.jcode))
    fr = sh = 0
    cdl = C.c_double(dlpf)
    cit = C.c_int(itecnt)
    if m.NE_TR:
        l.forces_tr(P(s.f_temp), P(s.ef_i), P(s.d), P(m.emod), P(m.carea), P(s.llength),
                    P(s.defllen_i), P(m.yld), P(s.c1_i), P(s.c2_i), P(s.c3_i), P(m.mcode))
    if m.NE_FR:
        fr = l.forces_fr(P(s.f_temp), P(s.ef_ip), P(s.ef_i), P(m.efFE_ref), P(s.efFE_ip),
                         P(s.efFE_i), P(s.yldflag), P(dd), P(m.emod), P(m.gmod), P(m.carea),
                         P(m.offset), P(m.osflag), P(s.llength), P(s.defllen_ip), P(m.istrong),
                         P(m.iweak), P(m.ipolar), P(m.iwarp), P(m.yld), P(m.zstrong),
                         P(m.zweak), P(s.c1_ip), P(s.c2_ip), P(s.c3_ip), P(s.c1_i), P(s.c2_i),
                         P(s.c3_i), P(m.mendrel), P(m.mcode), C.byref(cdl), C.byref(cit))
    if m.NE_SH:
        sh = l.forces_sh(P(s.f_temp), P(s.ef_ip), P(s.ef_i), P(s.efN_temp), P(s.efM_temp),
                         P(dd), P(s.d_temp), P(s.chi_temp), P(s.x_temp), P(s.x_ip), P(m.emod),
                         P(m.nu), P(m.xlocal), P(m.thick), P(s.farea), P(s.deffarea_ip),
                         P(s.slength), P(s.defslen_ip), P(m.yld), P(s.c1_ip), P(s.c2_ip),
                         P(s.c3_ip), P(s.c1_i), P(s.c2_i), P(s.c3_i), P(m.minc), P(m.mcode),
                         P(m.jcode))
    s.ef_ip[:] = s.ef_i
    return fr, sh, cdl.value


def forces_linear(m, s, d):
    """ANAFLAG==1 force recovery from total displacements (main.c:1774-1793)."""
    l = set_model(m, ANAFLAG=1)
    d = np.ascontiguousarray(d, dtype=np.float64)
    f = np.zeros(m.NEQ)
    cdl = C.c_double(0.0); cit = C.c_int(0)
    if m.NE_TR:
        l.forces_tr(P(f), P(s.ef), P(d), P(m.emod), P(m.carea), P(s.llength), P(s.defllen),
                    P(m.yld), P(s.c1), P(s.c2), P(s.c3), P(m.mcode))
    if m.NE_FR:
        l.forces_fr(P(f), P(s.ef), P(s.ef), P(m.efFE_ref), P(s.efFE), P(s.efFE), P(s.yldflag),
                    P(d), P(m.emod), P(m.gmod), P(m.carea), P(m.offset), P(m.osflag),
                    P(s.llength), P(s.defllen), P(m.istrong), P(m.iweak), P(m.ipolar),
                    P(m.iwarp), P(m.yld), P(m.zstrong), P(m.zweak), P(s.c1), P(s.c2), P(s.c3),
                    P(s.c1), P(s.c2), P(s.c3), P(m.mendrel), P(m.mcode), C.byref(cdl),
                    C.byref(cit))
    if m.NE_SH:
        l.forces_sh(P(f), P(s.ef), P(s.ef), P(s.efN), P(s.efM), P(d), P(d), P(s.chi), P(s.x),
                    P(s.x), P(m.emod), P(m.nu), P(m.xlocal), P(m.thick), P(s.farea),
                    P(s.deffarea), P(s.slength), P(s.defslen), P(m.yld), P(s.c1), P(s.c2),
                    P(s.c3), P(s.c1), P(s.c2), P(s.c3), P(m.minc), P(m.mcode), P(m.jcode))
    return f


def dense_to_csc(m, ss_dense):
    """Run the reference's own dense->CSC compaction (solve.c:110-119) by calling solve() on the
    SLVFLAG=2 static branch and reading back what it handed to umfpack_di_symbolic."""
    l = set_model(m, SLVFLAG=2)
    n = m.NEQ
    Ap = np.zeros(n + 1, dtype=np.int32)
    Ai = np.zeros(n * n, dtype=np.int32)
    Ax = np.zeros(n * n)
    r = np.ones(n); dd = np.zeros(n); uc = np.zeros(n)
    ssd = C.c_double(0); det = C.c_int(0)
    z = C.c_void_p(0)
    # ALGFLAG must be < 4 for the static branch
    l.ref_set_flags(C.c_int(m.ANAFLAG), C.c_int(1), C.c_int(2), C.c_int(1))
    l.solve(P(m.jcode), P(ss_dense), z, z, z, z, P(r), P(dd), P(m.maxa), C.byref(ssd),
            C.byref(det), z, z, z, P(uc), z, z, z, z, z, z, z, P(Ap), P(Ai), P(Ax),
            C.c_double(0), C.c_double(0), z, C.c_int(0), C.c_double(1), z, P(m.kht), z, z, z,
            C.c_int(0))
    nnz = l.ref_csc_nnz()
    Ap2 = np.zeros(n + 1, dtype=np.int32); Ai2 = np.zeros(nnz, dtype=np.int32)
    Ax2 = np.zeros(nnz)
    l.ref_csc_copy(P(Ap2), P(Ai2), P(Ax2))
    return Ap2, Ai2, Ax2, uc


def skyline_solve(m, ss, rhs, fact=0):
    """solve() on the SLVFLAG=0 static branch: skyfact + skysolve (solve.c:78-80, 539-698).
    ss is factorised in place when fact==0.  Returns (dd, ssd, det)."""
    l = set_model(m, SLVFLAG=0)
    l.ref_set_flags(C.c_int(m.ANAFLAG), C.c_int(1), C.c_int(0), C.c_int(1))
    n = m.NEQ
    r = np.ascontiguousarray(rhs, dtype=np.float64).copy(); dd = np.zeros(n)
    ssd = np.zeros(n); det = C.c_int(0)
    z = C.c_void_p(0)
    Ap = np.zeros(2, dtype=np.int32)          # solve() writes *pAp = 0 unconditionally (solve.c:65)
    err = l.solve(P(m.jcode), P(ss), z, z, z, z, P(r), P(dd), P(m.maxa), P(ssd),
                  C.byref(det), z, z, z, z, z, z, z, z, z, z, z, P(Ap), z, z,
                  C.c_double(0), C.c_double(0), z, C.c_int(fact), C.c_double(1), z, P(m.kht),
                  z, z, z, C.c_int(0))
    if err:
        raise RuntimeError("reference skyline solve failed")
    return dd, ssd, det.value


# ---- acoustic FSI (fsi.c) ----------------------------------------------------------------------
def fsi_matrices(m):
    """The reference's dense system matrices of an FSI model (ANAFLAG 4): L_br (fsi.c:447-531), stiff_fsi
    (fsi.c:333-386) and mass_fsi (fsi.c:388-445) called as main.c:1440-1596 does.  Returns (ss_fsi, sm_fsi)
    as [NEQ, NEQ] arrays indexed [row-major as the reference stores them]."""
    l = set_model(m, SLVFLAG=2, ANAFLAG=4)
    br = 1 if m.NE_SH == 0 else 0
    l.ref_set_fsi(L(m.SNDOF), L(m.FNDOF), C.c_int(br), C.c_int(1 - br))
    n, sn, fn = m.NEQ, m.SNDOF, m.FNDOF
    z = np.zeros
    Lm, A, G, LT = z(sn * fn), z(fn * fn), z(sn * fn), z(fn * sn)
    jcode_fsi = np.zeros(m.NJ * 7, dtype=np.int64)
    l.L_br(P(m.minc), P(m.mcode), P(m.jcode), P(jcode_fsi), P(m.nnorm), P(m.tarea), P(Lm), P(A), P(G))
    s = RefState(m)
    ss, ss_fsi, sm, sm_fsi = z(n * n), z(n * n), z(n * n), z(n * n)
    fdens = np.array([m.fdens], dtype=np.float64)
    l.stiff_fsi(P(m.minc), P(m.mcode), P(m.jcode), P(m.nnorm), P(m.tarea), P(s.farea), P(m.thick), P(s.deffarea),
                P(s.slength), P(s.defslen), P(Lm), P(A), P(ss), P(ss_fsi), P(s.x), P(m.xlocal), P(m.emod), P(m.nu),
                P(s.Jinv), P(s.jac), P(m.yld), P(s.c1), P(s.c2), P(s.c3), P(s.ef), P(s.d), P(s.chi), P(s.efN),
                P(s.efM), P(m.maxa if m.maxa is not None else np.zeros(n + 1, dtype=np.int64)))
    l.mass_fsi(P(m.minc), P(m.mcode), P(m.jcode), P(m.nnorm), P(m.tarea), P(m.carea), P(s.farea), P(m.thick),
               P(s.slength), P(Lm), P(LT), P(sm), P(sm_fsi), P(s.x), P(m.dens), P(fdens), P(s.Jinv), P(s.jac))
    l.ref_set_flags(C.c_int(m.ANAFLAG), C.c_int(m.ALGFLAG), C.c_int(m.SLVFLAG), C.c_int(1))
    return ss_fsi.reshape(n, n), sm_fsi.reshape(n, n)
