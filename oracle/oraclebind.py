"""TEST INFRASTRUCTURE ONLY - ctypes binding to oracle/libcubens_oracle.so, the plain-C
restatement of the hot path (oracle/cubens_oracle.c).  Mirrors oracle/refbind.py call for
call so tests can run the same sequence through the restatement, the compiled reference and the
CUDA path.  State bookkeeping (generations) is refbind.RefState - plain numpy."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from .refbind import P, RefState  # noqa: F401  (RefState re-exported)

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libcubens_oracle.so")
_lib = None


class orc_dims(C.Structure):
    _fields_ = [(n, C.c_long) for n in ("NJ", "NE_TR", "NE_FR", "NE_SH", "NE_BR", "NEQ")] + \
               [("ANAFLAG", C.c_int), ("SLVFLAG", C.c_int)]


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        for n in ("orc_codes", "orc_skylin", "orc_dense_to_csc"):
            getattr(_lib, n).restype = C.c_long
    return _lib


def dims(m, SLVFLAG=None, ANAFLAG=None):
    return orc_dims(m.NJ, m.NE_TR, m.NE_FR, m.NE_SH, m.NE_SBR + m.NE_FBR, m.NEQ,
                    m.ANAFLAG if ANAFLAG is None else ANAFLAG,
                    m.SLVFLAG if SLVFLAG is None else SLVFLAG)


def _gen(s, gen):
    if gen == "ip":
        return (s.c1_ip, s.c2_ip, s.c3_ip, s.ef_ip, s.defllen_ip, s.deffarea_ip, s.efFE_ip, s.x_temp)
    return (s.c1, s.c2, s.c3, s.ef, s.defllen, s.deffarea, s.efFE, s.x)


def _plastic(l, m, s, gen="ip"):
    if m.ANAFLAG == 3:
        l.orc_set_plastic(P(m.yld), P(m.zstrong), P(m.zweak), P(s.yldflag))
        if gen == "ip":
            l.orc_set_plastic_sh(P(s.chi_temp), P(s.efN_temp), P(s.efM_temp), P(s.x_ip),
                                 P(s.deffarea_ip), P(s.defslen_ip))
        else:
            l.orc_set_plastic_sh(P(s.chi), P(s.efN), P(s.efM), P(s.x), P(s.deffarea), P(s.defslen))


def stiff(m, s, SLVFLAG=None, gen="ip"):
    slv = m.SLVFLAG if SLVFLAG is None else SLVFLAG
    l = lib(); D = dims(m, SLVFLAG=slv)
    _plastic(l, m, s, gen)
    ss = np.zeros(m.lss if slv == 0 else m.NEQ * m.NEQ)
    c1, c2, c3, ef, dl, dfa, efFE, x = _gen(s, gen)
    if m.NE_TR:
        l.orc_stiff_tr(C.byref(D), P(ss), P(m.emod), P(m.carea), P(s.llength), P(dl), P(c1), P(c2),
                       P(c3), P(ef), P(m.maxa), P(m.mcode))
    if m.NE_FR:
        l.orc_stiff_fr(C.byref(D), P(ss), P(m.emod), P(m.gmod), P(m.carea), P(m.offset),
                       P(m.osflag), P(s.llength), P(dl), P(m.istrong), P(m.iweak), P(m.ipolar),
                       P(m.iwarp), P(c1), P(c2), P(c3), P(ef), P(efFE), P(m.mendrel), P(m.maxa),
                       P(m.mcode))
    if m.NE_SH:
        l.orc_stiff_sh(C.byref(D), P(ss), P(m.emod), P(m.nu), P(x), P(m.xlocal), P(m.thick),
                       P(s.farea), P(dfa), P(s.slength), P(c1), P(c2), P(c3), P(m.maxa), P(m.minc),
                       P(m.mcode))
    if m.NE_BR:
        assert slv != 0
        l.orc_stiff_br(C.byref(D), P(ss), P(x), P(m.emod), P(m.nu), P(m.minc), P(m.mcode))
    return ss


def shell_element_K(m, s, n, gen="ip"):
    l = lib(); D = dims(m)
    c1, c2, c3, ef, dl, dfa, efFE, x = _gen(s, gen)
    K = np.zeros(324)
    l.orc_shell_element_K(C.byref(D), C.c_long(n), P(K), P(m.emod), P(m.nu), P(x), P(m.xlocal),
                          P(m.thick), P(s.farea), P(dfa), P(s.slength), P(c1), P(c2), P(c3),
                          P(m.minc))
    return K.reshape(18, 18)


def update_forces(m, s, dd, dlpf=1.0, itecnt=0):
    l = lib(); D = dims(m)
    _plastic(l, m, s)
    fr = sh = 0
    cdl = C.c_double(dlpf)
    dd = np.ascontiguousarray(dd, dtype=np.float64)
    s.d_temp += dd
    s.f_temp[:] = 0
    l.orc_updatc(C.byref(D), P(s.x_temp), P(s.x_ip), P(s.xfr_temp), P(dd), P(s.defllen_i),
                 P(s.deffarea_i), P(s.defslen_i), P(m.offset), P(m.osflag), P(m.auxpt), P(s.c1_i),
                 P(s.c2_i), P(s.c3_i), P(m.minc), P(m.jcode))
    if m.NE_TR:
        l.orc_forces_tr(C.byref(D), P(s.f_temp), P(s.ef_i), P(s.d), P(m.emod), P(m.carea),
                        P(s.llength), P(s.defllen_i), P(s.c1_i), P(s.c2_i), P(s.c3_i), P(m.mcode))
    if m.NE_FR:
        fr = l.orc_forces_fr(C.byref(D), P(s.f_temp), P(s.ef_ip), P(s.ef_i), P(m.efFE_ref), P(s.efFE_ip),
                        P(s.efFE_i), P(dd), P(m.emod), P(m.gmod), P(m.carea), P(m.offset),
                        P(m.osflag), P(s.llength), P(s.defllen_ip), P(m.istrong), P(m.iweak),
                        P(m.ipolar), P(m.iwarp), P(s.c1_ip), P(s.c2_ip), P(s.c3_ip), P(s.c1_i),
                        P(s.c2_i), P(s.c3_i), P(m.mendrel), P(m.mcode), C.byref(cdl),
                        C.c_int(itecnt))
    if m.NE_SH:
        sh = l.orc_forces_sh(C.byref(D), P(s.f_temp), P(s.ef_ip), P(s.ef_i), P(dd), P(s.d_temp),
                        P(s.x_temp), P(m.emod), P(m.nu), P(m.xlocal), P(m.thick), P(s.farea),
                        P(s.slength), P(s.c1_ip), P(s.c2_ip), P(s.c3_ip), P(s.c1_i), P(s.c2_i),
                        P(s.c3_i), P(m.minc), P(m.mcode))
    s.ef_ip[:] = s.ef_i
    return fr, sh, cdl.value


def forces_linear(m, s, d, dlpf=0.0):
    l = lib(); D = dims(m, ANAFLAG=1)
    d = np.ascontiguousarray(d, dtype=np.float64)
    f = np.zeros(m.NEQ)
    if m.NE_TR:
        l.orc_forces_tr(C.byref(D), P(f), P(s.ef), P(d), P(m.emod), P(m.carea), P(s.llength),
                        P(s.defllen), P(s.c1), P(s.c2), P(s.c3), P(m.mcode))
    if m.NE_FR:
        l.orc_forces_fr(C.byref(D), P(f), P(s.ef), P(s.ef), P(m.efFE_ref), P(s.efFE), P(s.efFE), P(d),
                        P(m.emod), P(m.gmod), P(m.carea), P(m.offset), P(m.osflag), P(s.llength),
                        P(s.defllen), P(m.istrong), P(m.iweak), P(m.ipolar), P(m.iwarp), P(s.c1),
                        P(s.c2), P(s.c3), P(s.c1), P(s.c2), P(s.c3), P(m.mendrel), P(m.mcode),
                        C.byref(C.c_double(dlpf)), C.c_int(0))
    if m.NE_SH:
        l.orc_forces_sh(C.byref(D), P(f), P(s.ef), P(s.ef), P(d), P(d), P(s.x), P(m.emod), P(m.nu),
                        P(m.xlocal), P(m.thick), P(s.farea), P(s.slength), P(s.c1), P(s.c2),
                        P(s.c3), P(s.c1), P(s.c2), P(s.c3), P(m.minc), P(m.mcode))
    return f


def mass(m, s, SLVFLAG=0):
    l = lib(); D = dims(m, SLVFLAG=SLVFLAG)
    sm = np.zeros(m.NEQ if SLVFLAG == 0 else m.NEQ * m.NEQ)

    if m.NE_TR:
        l.orc_mass_tr(C.byref(D), P(sm), P(m.carea), P(s.llength), P(m.dens), P(s.x), P(m.minc),
                      P(m.mcode))
    if m.NE_FR:
        l.orc_mass_fr(C.byref(D), P(sm), P(m.carea), P(s.llength), P(m.dens), P(m.osflag),
                      P(m.offset), P(s.x), P(s.xfr), P(m.minc), P(m.mcode))
    if m.NE_SH:
        l.orc_mass_sh(C.byref(D), P(sm), P(m.dens), P(m.thick), P(s.farea), P(s.slength), P(s.x),
                      P(m.minc), P(m.mcode))
    if m.NE_BR:
        assert SLVFLAG != 0 and not (m.NE_TR or m.NE_FR)
        l.orc_mass_br(C.byref(D), P(sm), P(m.dens), P(s.x), P(m.minc), P(m.mcode))
    return sm


def codes(m, jflags):
    l = lib(); D = dims(m)
    jc = np.ascontiguousarray(jflags, dtype=np.int64).reshape(-1).copy()
    mc = np.zeros(m.n_mcode, dtype=np.int64)
    neq = l.orc_codes(C.byref(D), P(mc), P(jc), P(m.minc))
    return jc, mc, neq


def skylin(m):
    l = lib(); D = dims(m)
    maxa = np.zeros(m.NEQ + 1, dtype=np.int64); kht = np.zeros(m.NEQ, dtype=np.int64)
    lss = l.orc_skylin(C.byref(D), P(maxa), P(kht), P(m.mcode))
    return kht, maxa, lss


def dense_to_csc(neq, ss, tol=1e-10):
    l = lib()
    Ap = np.zeros(neq + 1, dtype=np.int32); Ai = np.zeros(neq * neq, dtype=np.int32)
    Ax = np.zeros(neq * neq)
    nz = l.orc_dense_to_csc(C.c_long(neq), P(np.ascontiguousarray(ss)), C.c_double(tol), P(Ap),
                            P(Ai), P(Ax))
    return Ap, Ai[:nz].copy(), Ax[:nz].copy()
