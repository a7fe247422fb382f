"""cubens_b200 - Python face of the C-ABI in include/cubens_b200.h (ctypes, no torch types).

The reference is C and so is the real host (cu-bens_b200/host/); this module is
what bench.py and the tests drive: it loads the in-tree ``libcubens_b200.so`` (hand-written
sm_100a kernels), mirrors the reference's call sequence with the reference's names, and FAILS
LOUDLY if the library or a CUDA device is missing - there is no CPU path behind it.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from .model import Model  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "libcubens_b200.so"))

CB_GEN_IP, CB_GEN_COMMITTED = 0, 1
CB_MAT_CSC, CB_MAT_SKYLINE, CB_MAT_BOTH = 1, 2, 3

ARR = dict(X=1, X_TEMP=2, X_IP=3, C1=4, C2=5, C3=6, C1_I=7, C2_I=8, C3_I=9, C1_IP=10, C2_IP=11,
           C3_IP=12, EF=13, EF_I=14, EF_IP=15, DEFLLEN=16, DEFLLEN_I=17, DEFLLEN_IP=18,
           DEFFAREA=19, DEFFAREA_I=20, DEFFAREA_IP=21, DEFSLEN=22, DEFSLEN_I=23, DEFSLEN_IP=24,
           XFR=25, XFR_TEMP=26, EFFE=27, EFFE_I=28, EFFE_IP=29, D=30, D_TEMP=31, F=32, F_TEMP=33,
           LLENGTH=34, FAREA=35, SLENGTH=36, CHI=37, CHI_TEMP=38, EFN=39, EFN_TEMP=40, EFM=41,
           EFM_TEMP=42)

EXPORTS = [
    "cb_abi_version", "cb_last_error", "cb_device_count", "cb_create", "cb_destroy",
    "cb_set_owned_joints", "cb_begin_increment", "cb_stiff", "cb_mass", "cb_update_forces",
    "cb_update_forces_dev", "cb_forces_linear", "cb_end_iteration", "cb_commit",
    "cb_get_skyline", "cb_csc_nnz", "cb_csc_pattern", "cb_get_csc_values", "cb_csc_compact",
    "cb_get_mass", "cb_get_f", "cb_dev_Ax", "cb_dev_skyline", "cb_dev_f", "cb_dev_dd",
    "cb_dev_Ap", "cb_dev_Ai", "cb_download", "cb_upload", "cb_launch_count",
    "cb_last_stiff_ms", "cb_last_forces_ms", "cb_last_assemble_ms", "cb_timer_start", "cb_timer_stop_ms", "cb_set_dd", "cb_host_alloc",
    "cb_host_free", "cb_map_bytes", "cb_sync", "cb_stream", "cb_set_q", "cb_residual_sums",
    "cb_dev_sums", "cb_get_sums", "cb_get_yldflag", "cb_set_yldflag", "cb_get_mass_csc_values", "cb_dev_Mx",
    "cb_geometry_classes", "cb_keep_ip", "cb_checkpoint_save", "cb_checkpoint_load",
    "cb_set_element_ids", "cb_update_forces_begin", "cb_update_forces_end", "cb_measure_fp64_tflops",
    "cb_plan_selfcheck", "cb_csc_upper_nnz", "cb_csc_upper_pattern", "cb_get_csc_upper_values",
    "cb_csc_values_begin", "cb_csc_values_end", "cb_get_csc_values_mirrored", "cb_sym_selftest",
    "cb_local_equations", "cb_csc_values_d2h_bytes", "cb_get_reaction_sums", "cb_comm_unique_id", "cb_comm_init",
    "cb_comm_destroy", "cb_residual_allreduce", "cb_residual_sums_allreduce", "cb_comm_peer_memory", "cb_trip_allreduce", "cb_convergence_test", "cb_plan_info",
    "cb_debug_stream_plan",
]


class cb_sizes(C.Structure):
    _fields_ = [(n, C.c_long) for n in ("NJ", "NE_TR", "NE_FR", "NE_SH", "NE_SBR", "NE_FBR", "NEQ")]


class cb_flags(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("ANAFLAG", "ALGFLAG", "SLVFLAG", "matrix_layout", "device")]


_MODEL_FIELDS = ["x", "minc", "jcode", "mcode", "maxa", "emod", "dens", "carea", "llength", "c1",
                 "c2", "c3", "nu", "thick", "farea", "slength", "xlocal", "gmod", "istrong",
                 "iweak", "ipolar", "iwarp", "auxpt", "offset", "osflag", "mendrel", "efFE_ref", "yield",
                 "zstrong", "zweak", "nnorm", "tarea", "fdens"]


class cb_model(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _MODEL_FIELDS]


class CubensError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """dlopen the CUDA library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("CUBENS_LIB") or LIB_PATH      # CUBENS_LIB: kernel-variant builds (development)
    if not os.path.exists(path):
        raise CubensError(f"{path} not found - build it with `make -C cu-bens_b200` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    lib.cb_last_error.restype = C.c_char_p
    lib.cb_csc_nnz.restype = C.c_long
    lib.cb_csc_compact.restype = C.c_long
    lib.cb_csc_upper_nnz.restype = C.c_long
    lib.cb_local_equations.restype = C.c_long
    lib.cb_csc_values_d2h_bytes.restype = C.c_long
    lib.cb_debug_stream_plan.restype = C.c_long
    lib.cb_launch_count.restype = C.c_long
    lib.cb_map_bytes.restype = C.c_long
    lib.cb_last_stiff_ms.restype = C.c_double
    lib.cb_last_forces_ms.restype = C.c_double
    lib.cb_last_assemble_ms.restype = C.c_double
    lib.cb_timer_stop_ms.restype = C.c_double
    lib.cb_measure_fp64_tflops.restype = C.c_double
    lib.cb_host_alloc.restype = C.c_void_p
    lib.cb_host_alloc.argtypes = [C.c_ulong]
    lib.cb_host_free.restype = None
    lib.cb_host_free.argtypes = [C.c_void_p]
    for n in ("cb_dev_Mx", "cb_dev_Ax", "cb_dev_skyline", "cb_dev_f", "cb_dev_dd", "cb_dev_Ap", "cb_dev_Ai",
              "cb_stream", "cb_dev_sums"):
        getattr(lib, n).restype = C.c_void_p
    lib.cb_destroy.restype = None
    _lib = lib
    return lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)


def _model_array(m, n):
    """the array behind cb_model field ``n`` of a Model in the C dtype, or None"""
    a = getattr(m, "yld" if n == "yield" else n, None)
    if n == "fdens":
        a = np.array([a], dtype=np.float64) if getattr(m, "nnorm", None) is not None else None
    if a is None:
        return None
    dt = np.int32 if n in ("osflag", "mendrel") else (np.int64 if n in ("minc", "jcode", "mcode", "maxa") else np.float64)
    return np.ascontiguousarray(a, dtype=dt)


def _c_model(m, layout, device):
    """the C structs of include/cubens_b200.h for a Model (plus the arrays that must stay alive)"""
    sz = cb_sizes(m.NJ, m.NE_TR, m.NE_FR, m.NE_SH, m.NE_SBR, m.NE_FBR, m.NEQ)
    fl = cb_flags(m.ANAFLAG, m.ALGFLAG, m.SLVFLAG, layout, device)
    keep = {}
    cm = cb_model()
    for n in _MODEL_FIELDS:
        a = _model_array(m, n)
        if a is not None:
            keep[n] = a
        setattr(cm, n, _p(a) if a is not None else C.c_void_p(0))
    return sz, fl, cm, keep


def plan_selfcheck(m, j0=0, j1=0, layout=CB_MAT_CSC):
    """cb_plan_selfcheck: build the element-to-nonzero maps / tile plans of ``m`` on the host (no device)
    and interpret them the way the kernels do.  Returns dict(nnz, tiles, rows, pairs, steps, kind);
    raises CubensError when the plan is inconsistent."""
    lib = load_library()
    sz, fl, cm, keep = _c_model(m, layout, 0)
    st = (C.c_long * 6)()
    rc = lib.cb_plan_selfcheck(C.byref(sz), C.byref(fl), C.byref(cm), C.c_long(j0), C.c_long(j1), st)
    if rc != 0:
        raise CubensError(f"cb_plan_selfcheck error {rc}: {lib.cb_last_error().decode()}")
    return dict(zip(("nnz", "tiles", "rows", "pairs", "steps", "kind"), list(st)))


def sym_selftest(m, j0=0, j1=0, nthreads=4, layout=CB_MAT_CSC):
    """cb_sym_selftest: packed upper-triangle layout + threaded rebuild of the full matrix, on the host.
    Returns the wall time of the rebuild in seconds."""
    lib = load_library()
    sz, fl, cm, keep = _c_model(m, layout, 0)
    sec = C.c_double(0.0)
    rc = lib.cb_sym_selftest(C.byref(sz), C.byref(fl), C.byref(cm), C.c_long(j0), C.c_long(j1), C.c_int(nthreads),
                             C.byref(sec))
    if rc != 0:
        raise CubensError(f"cb_sym_selftest error {rc}: {lib.cb_last_error().decode()}")
    return sec.value


class Assembler:
    """One model resident on one B200.  Method names follow the reference call sites they
    replace (see include/cubens_b200.h for the file:line table)."""

    def __init__(self, m, layout=0, device=0):
        self.lib = load_library()
        self.m = m
        if self.lib.cb_device_count() <= 0:
            raise CubensError("no CUDA device visible - the element/assembly path has no CPU fallback")
        sz, fl, cm, keep = _c_model(m, layout, device)
        h = C.c_void_p(0)
        rc = self.lib.cb_create(C.byref(sz), C.byref(fl), C.byref(cm), C.byref(h))
        self._check(rc)
        self.h = h
        self.layout = layout if layout else (CB_MAT_SKYLINE if m.SLVFLAG == 0 else CB_MAT_CSC)

    def _check(self, rc):
        if rc != 0:
            raise CubensError(f"cubens_b200 error {rc}: {self.lib.cb_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            for ptr in getattr(self, "_pinned", []):
                self.lib.cb_host_free(C.c_void_p(ptr))
            self._pinned = []
            self.lib.cb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- main.c loop mirror ---------------------------------------------------------------
    def set_owned_joints(self, j0, j1):
        self._check(self.lib.cb_set_owned_joints(self.h, C.c_long(j0), C.c_long(j1)))

    def begin_increment(self):
        self._check(self.lib.cb_begin_increment(self.h))

    def stiff(self, gen=CB_GEN_IP):
        self._check(self.lib.cb_stiff(self.h, C.c_int(gen)))

    def mass(self):
        self._check(self.lib.cb_mass(self.h))
        sm = np.zeros(self.m.NEQ)
        self._check(self.lib.cb_get_mass(self.h, _p(sm)))
        return sm

    def mass_csc(self):
        """full-order mass matrix values on the CSC pattern of K_t (models with bricks)"""
        self._check(self.lib.cb_mass(self.h))
        Mx = np.zeros(self.lib.cb_csc_nnz(self.h))
        self._check(self.lib.cb_get_mass_csc_values(self.h, _p(Mx)))
        return Mx

    def update_forces(self, dd, dlpf=1.0, itecnt=0, want_f=True):
        dd = np.ascontiguousarray(dd, dtype=np.float64)
        f = np.zeros(self.m.NEQ) if want_f else None
        cdl = C.c_double(dlpf); fr = C.c_int(0); sh = C.c_int(0)
        self._check(self.lib.cb_update_forces(self.h, _p(dd), C.byref(cdl), C.c_int(itecnt),
                                              _p(f) if want_f else C.c_void_p(0), C.byref(fr),
                                              C.byref(sh)))
        return f, fr.value, sh.value, cdl.value

    # ---- element-partitioned ANAFLAG 3: the force pass in two halves around the exchange of the
    # lowest tripping element index (include/cubens_b200.h, cb_update_forces_begin / _end)
    def set_element_ids(self, fr_gid=None, sh_gid=None):
        fr = None if fr_gid is None else np.ascontiguousarray(fr_gid, dtype=np.int32)
        sh = None if sh_gid is None else np.ascontiguousarray(sh_gid, dtype=np.int32)
        self._check(self.lib.cb_set_element_ids(self.h, _p(fr), _p(sh)))

    def update_forces_begin(self, dd, dlpf=1.0, itecnt=0):
        """dd: host [NEQ] (uploaded) or None (already in cb_dev_dd()).  Returns (first_fr, first_sh):
        lowest global index of a frame / shell that trips on this rank (INT_MAX: none)."""
        if dd is not None:
            self.set_dd(dd)
        ffr = C.c_int(0); fsh = C.c_int(0)
        self._check(self.lib.cb_update_forces_begin(self.h, C.c_void_p(0), C.c_double(dlpf), C.c_int(itecnt),
                                                    C.byref(ffr), C.byref(fsh)))
        return ffr.value, fsh.value

    def update_forces_end(self, first_fr, first_sh, dlpf=1.0, want_f=True):
        cdl = C.c_double(dlpf); fr = C.c_int(0); sh = C.c_int(0)
        self._check(self.lib.cb_update_forces_end(self.h, C.c_int(first_fr), C.c_int(first_sh), C.byref(cdl),
                                                  C.byref(fr), C.byref(sh)))
        f = None
        if want_f:
            f = np.zeros(self.m.NEQ)
            self._check(self.lib.cb_get_f(self.h, _p(f)))
        return f, fr.value, sh.value, cdl.value

    def update_forces_dev(self, dlpf=1.0, itecnt=0):
        """dd already resident in the device buffer cb_dev_dd(); f_temp stays on the device."""
        cdl = C.c_double(dlpf); fr = C.c_int(0); sh = C.c_int(0)
        self._check(self.lib.cb_update_forces_dev(self.h, C.c_void_p(self.lib.cb_dev_dd(self.h)),
                                                  C.byref(cdl), C.c_int(itecnt), C.byref(fr),
                                                  C.byref(sh)))
        return fr.value, sh.value

    def forces_linear(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        f = np.zeros(self.m.NEQ)
        self._check(self.lib.cb_forces_linear(self.h, _p(d), _p(f)))
        return f

    def end_iteration(self):
        self._check(self.lib.cb_end_iteration(self.h))

    def checkpoint_save(self, path):
        self._check(self.lib.cb_checkpoint_save(self.h, str(path).encode()))

    def checkpoint_load(self, path):
        self._check(self.lib.cb_checkpoint_load(self.h, str(path).encode()))

    def yldflag(self):
        y = np.zeros(2 * self.m.NE_FR, dtype=np.int32)
        self._check(self.lib.cb_get_yldflag(self.h, _p(y), C.c_long(y.size)))
        return y

    def set_yldflag(self, y):
        y = np.ascontiguousarray(y, dtype=np.int32)
        self._check(self.lib.cb_set_yldflag(self.h, _p(y), C.c_long(y.size)))

    def commit(self):
        self._check(self.lib.cb_commit(self.h))

    # ---- results --------------------------------------------------------------------------
    def skyline(self):
        ss = np.zeros(self.m.lss)
        self._check(self.lib.cb_get_skyline(self.h, _p(ss), C.c_long(ss.size)))
        return ss

    def csc(self):
        nnz = self.lib.cb_csc_nnz(self.h)
        if nnz < 0:
            raise CubensError(self.lib.cb_last_error().decode())
        Ap = np.zeros(self.m.NEQ + 1, dtype=np.int32); Ai = np.zeros(nnz, dtype=np.int32)
        Ax = np.zeros(nnz)
        self._check(self.lib.cb_csc_pattern(self.h, _p(Ap), _p(Ai)))
        self._check(self.lib.cb_get_csc_values(self.h, _p(Ax)))
        return Ap, Ai, Ax

    def csc_values(self):
        Ax = np.zeros(self.lib.cb_csc_nnz(self.h))
        self._check(self.lib.cb_get_csc_values(self.h, _p(Ax)))
        return Ax

    def csc_upper(self):
        """upper-triangular CSC of the owned slice (Apu, Aiu, Axu)"""
        nu = self.lib.cb_csc_upper_nnz(self.h)
        if nu < 0:
            raise CubensError(self.lib.cb_last_error().decode())
        Apu = np.zeros(self.m.NEQ + 1, dtype=np.int32); Aiu = np.zeros(nu, dtype=np.int32); Axu = np.zeros(nu)
        self._check(self.lib.cb_csc_upper_pattern(self.h, _p(Apu), _p(Aiu)))
        self._check(self.lib.cb_get_csc_upper_values(self.h, _p(Axu)))
        return Apu, Aiu, Axu

    def csc_values_mirrored(self, nthreads=8, Ax=None, staging=None):
        """full Ax rebuilt on the host from the packed upper triangle (half the PCIe traffic)"""
        nnz = self.lib.cb_csc_nnz(self.h); nu = self.lib.cb_csc_upper_nnz(self.h)
        Ax = np.zeros(nnz) if Ax is None else Ax
        staging = self.pinned(nu) if staging is None else staging
        self._check(self.lib.cb_get_csc_values_mirrored(self.h, _p(Ax), C.c_void_p(staging.ctypes.data), C.c_int(nthreads)))
        return Ax

    def csc_compact(self, drop_tol=1e-10):
        nnz = self.lib.cb_csc_nnz(self.h)
        Ap = np.zeros(self.m.NEQ + 1, dtype=np.int32); Ai = np.zeros(nnz, dtype=np.int32)
        Ax = np.zeros(nnz)
        nz = self.lib.cb_csc_compact(self.h, C.c_double(drop_tol), _p(Ap), _p(Ai), _p(Ax))
        if nz < 0:
            raise CubensError(self.lib.cb_last_error().decode())
        return Ap, Ai[:nz].copy(), Ax[:nz].copy()

    def download(self, name):
        m = self.m
        n = {"X": m.NJ * 3, "X_TEMP": m.NJ * 3, "X_IP": m.NJ * 3, "D": m.NEQ, "D_TEMP": m.NEQ,
             "F": m.NEQ, "F_TEMP": m.NEQ, "LLENGTH": m.NE_TR + m.NE_FR, "FAREA": m.NE_SH,
             "SLENGTH": 3 * m.NE_SH, "XFR": 6 * m.NE_FR, "XFR_TEMP": 6 * m.NE_FR}.get(name)
        if n is None:
            if name.startswith("C"):
                n = m.n_c
            elif name.startswith("CHI"):
                n = 3 * m.NE_SH
            elif name.startswith("EFN") or name.startswith("EFM"):
                n = 9 * m.NE_SH
            elif name.startswith("EFFE"):
                n = 14 * m.NE_FR
            elif name.startswith("EF"):
                n = m.n_ef
            elif name.startswith("DEFLLEN"):
                n = m.NE_TR + m.NE_FR
            elif name.startswith("DEFFAREA"):
                n = m.NE_SH
            elif name.startswith("DEFSLEN"):
                n = 3 * m.NE_SH
        out = np.zeros(n)
        self._check(self.lib.cb_download(self.h, C.c_int(ARR[name]), _p(out), C.c_long(n)))
        return out

    def upload(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self._check(self.lib.cb_upload(self.h, C.c_int(ARR[name]), _p(a), C.c_long(a.size)))

    # ---- instrumentation -------------------------------------------------------------------
    @property
    def launches(self):
        return self.lib.cb_launch_count(self.h)

    @property
    def last_stiff_ms(self):
        return self.lib.cb_last_stiff_ms(self.h)

    @property
    def last_forces_ms(self):
        return self.lib.cb_last_forces_ms(self.h)

    @property
    def last_assemble_ms(self):
        return self.lib.cb_last_assemble_ms(self.h)

    def timer_start(self):
        self._check(self.lib.cb_timer_start(self.h))

    def timer_stop_ms(self):
        return self.lib.cb_timer_stop_ms(self.h)

    def sync(self):
        self._check(self.lib.cb_sync(self.h))

    def set_q(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        self._check(self.lib.cb_set_q(self.h, _p(q)))

    def residual_sums(self, lpf=1.0, fetch=False):
        self._check(self.lib.cb_residual_sums(self.h, C.c_double(lpf)))
        if fetch:
            s = np.zeros(5)
            self._check(self.lib.cb_get_sums(self.h, _p(s)))
            return s

    def convergence_test(self, lpf, intener1, toldisp, tolforc, tolener):
        """test() of the reference (misc.c:187-250) on the device-resident vectors; returns (err, convchk, sums5)"""
        conv = C.c_int(0); s = np.zeros(5)
        err = self.lib.cb_convergence_test(self.h, C.c_double(lpf), C.c_double(intener1), C.c_double(toldisp),
                                           C.c_double(tolforc), C.c_double(tolener), C.byref(conv), _p(s))
        return err, conv.value, s

    def plan_info(self):
        """(seconds the element-to-nonzero maps took to build, built on the device?)"""
        sec = C.c_double(0); dev = C.c_int(0)
        self._check(self.lib.cb_plan_info(self.h, C.byref(sec), C.byref(dev)))
        return sec.value, bool(dev.value)

    def debug_stream_plan(self, which):
        n = self.lib.cb_debug_stream_plan(self.h, C.c_int(which), C.c_void_p(0))
        if n < 0:
            return None
        buf = np.zeros(n, dtype=np.uint8)
        assert self.lib.cb_debug_stream_plan(self.h, C.c_int(which), _p(buf)) == n
        return buf

    def reaction_sums(self):
        r = np.zeros(6)
        self._check(self.lib.cb_get_reaction_sums(self.h, _p(r)))
        return r

    # ---- the collective inside the library (NCCL bound at run time) ---------------------------
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        self._check(self.lib.cb_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, uid, rank, world):
        assert len(uid) == 128
        self._check(self.lib.cb_comm_init(self.h, C.c_char_p(uid), C.c_int(rank), C.c_int(world)))

    @property
    def comm_peer_memory(self):
        return bool(self.lib.cb_comm_peer_memory(self.h))

    def residual_sums_allreduce(self, lpf=1.0):
        """cb_residual_sums + the all-reduce over the ranks in one launch where the ranks share peer memory"""
        self._check(self.lib.cb_residual_sums_allreduce(self.h, C.c_double(lpf)))

    def residual_allreduce(self):
        self._check(self.lib.cb_residual_allreduce(self.h))

    def trip_allreduce(self, first_fr, first_sh):
        a = C.c_int(first_fr); b = C.c_int(first_sh)
        self._check(self.lib.cb_trip_allreduce(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_dd(self, dd):
        dd = np.ascontiguousarray(dd, dtype=np.float64)
        self._check(self.lib.cb_set_dd(self.h, _p(dd)))

    def pinned(self, n, dtype=np.float64):
        """numpy view of a page-locked host buffer (freed with the Assembler)"""
        nbytes = int(n) * np.dtype(dtype).itemsize
        ptr = self.lib.cb_host_alloc(C.c_ulong(max(nbytes, 8)))
        if not ptr:
            raise CubensError("cudaHostAlloc failed")
        self._pinned = getattr(self, "_pinned", []) + [ptr]
        buf = (C.c_char * max(nbytes, 8)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    @property
    def geometry_classes(self):
        return self.lib.cb_geometry_classes(self.h)

    @property
    def map_bytes(self):
        return self.lib.cb_map_bytes(self.h)


# ---- C host drivers (cu-bens_b200/host) ------------------------------------------------------
HOST_LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "libcubens_host.so"))
_hostlib = None


class cb_nr_params(C.Structure):
    _fields_ = [("lpfmax", C.c_double), ("lpf", C.c_double), ("dlpf", C.c_double),
                ("dlpfmax", C.c_double), ("dlpfmin", C.c_double), ("itemax", C.c_int),
                ("submax", C.c_int), ("solmin", C.c_int), ("toldisp", C.c_double),
                ("tolforc", C.c_double), ("tolener", C.c_double), ("algflag", C.c_int)]


class cb_nr_result(C.Structure):
    _fields_ = [("status", C.c_int), ("increments", C.c_int), ("iterations", C.c_int),
                ("stiff_calls", C.c_int), ("force_calls", C.c_int), ("lpf", C.c_double)]


def load_host_library():
    global _hostlib
    if _hostlib is None:
        load_library()
        if not os.path.exists(HOST_LIB_PATH):
            raise CubensError(f"{HOST_LIB_PATH} not found - run make -C cu-bens_b200")
        _hostlib = C.CDLL(HOST_LIB_PATH)
    return _hostlib


def _solver_args(m, csc):
    """(maxa, lss) of the skyline solver, or (NULL, 0): the drivers then factorise the device-built
    CSC matrix (cb_csc_pattern / cb_get_csc_values -> sparse LDL^T or UMFPACK, host/cb_sparse.c)"""
    if csc:
        return None, C.c_void_p(0), C.c_long(0)
    maxa = np.ascontiguousarray(m.maxa, dtype=np.int64)
    return maxa, _p(maxa), C.c_long(m.lss)


def newton_static(asm, q, lpfmax=1.0, lpf=0.1, dlpf=0.1, dlpfmax=0.5, dlpfmin=1e-4, itemax=20,
                  submax=5, solmin=10, toldisp=1e-8, tolforc=1e-8, tolener=1e-8, algflag=1,
                  hist_dof=-1, csc=False):
    """the C host driver cb_newton_static (main.c:1824-2152 on the device path)"""
    hl = load_host_library()
    m = asm.m
    p = cb_nr_params(lpfmax, lpf, dlpf, dlpfmax, dlpfmin, itemax, submax, solmin, toldisp, tolforc,
                     tolener, algflag)
    res = cb_nr_result()
    d = np.zeros(m.NEQ)
    hist = np.zeros(3 * 4096)
    q = np.ascontiguousarray(q, dtype=np.float64)
    _keep, maxa_p, lss = _solver_args(m, csc)
    hl.cb_newton_static(asm.h, C.c_long(m.NEQ), maxa_p, lss, _p(q), C.byref(p), _p(d),
                        C.byref(res), _p(hist), C.c_int(4096), C.c_long(hist_dof))
    n = res.increments
    return d, res, hist[:3 * n].reshape(-1, 3).copy()


def sky_factor_solve(maxa, ss, rhs):
    """host skyline LDL^T (cb_sky_factor + cb_sky_solve); ss is factorised in place"""
    hl = load_host_library()
    neq = len(rhs)
    v = np.ascontiguousarray(rhs, dtype=np.float64).copy()
    maxa = np.ascontiguousarray(maxa, dtype=np.int64)
    rc = hl.cb_sky_factor(C.c_long(neq), _p(maxa), _p(ss), C.c_void_p(0), C.c_void_p(0), C.c_int(0))
    if rc:
        raise CubensError("non-positive definite stiffness matrix")
    hl.cb_sky_solve(C.c_long(neq), _p(maxa), _p(ss), _p(v))
    return v


def newmark(asm, dyn, nonlinear=True, csc=False):
    """the C host drivers cb_newmark_nonlinear / cb_newmark_linear (main.c:3305-3960 /
    main.c:3143-3303 + solve.c:199-458 on the device path).  dyn = deck.parse_dynamic_tail(...).
    Returns (hist [ntstps][2+NEQ], result)."""
    hl = load_host_library()
    m = asm.m
    if dyn["nbc"] and not nonlinear:
        raise CubensError("prescribed support motion (NBC != 0) is not supported by the linear driver")
    nt = int(dyn["ntstps"])
    hist = np.zeros((nt, m.NEQ + 2))
    res = cb_nr_result()
    pin = np.ascontiguousarray(dyn["pinpt"], dtype=np.float64)
    _keep, maxa_p, lss = _solver_args(m, csc)
    if nonlinear:
        q = dyn["params"]
        p = cb_nr_params(q["lpfmax"], q["lpf"], q["dlpf"], q["dlpfmax"], q["dlpfmin"], q["itemax"],
                         q["submax"], q["solmin"], q["toldisp"], q["tolforc"], q["tolener"], 1)
        pdisp = np.ascontiguousarray(dyn["pdisp"], dtype=np.float64)
        pmot = np.ascontiguousarray(dyn["pmot"], dtype=np.int32)
        hl.cb_newmark_nonlinear_bc(asm.h, C.c_long(m.NEQ), maxa_p, lss, _p(pin), _p(pdisp),
                                   _p(pmot), C.c_long(nt), C.c_double(dyn["dt"]), C.c_double(dyn["alpham"]),
                                   C.c_double(dyn["alphaf"]), C.byref(p), _p(hist), C.byref(res))
    else:
        um, vm, am = (np.ascontiguousarray(dyn[k], dtype=np.float64) for k in ("um", "vm", "am"))
        hl.cb_newmark_linear(asm.h, C.c_long(m.NEQ), maxa_p, lss, _p(pin), C.c_long(nt),
                             C.c_double(dyn["dt"]), C.c_double(dyn["alpham"]), C.c_double(dyn["alphaf"]),
                             _p(um), _p(vm), _p(am), _p(hist), C.byref(res))
    return hist, res


class cb_arc_params(C.Structure):
    _fields_ = [("dk", C.c_double), ("dkdof", C.c_long), ("alpha", C.c_double), ("psi_thresh", C.c_double),
                ("iteopt", C.c_int), ("lpfmax", C.c_double), ("dkimax", C.c_double), ("itemax", C.c_int),
                ("submax", C.c_int), ("imagmax", C.c_int), ("negmax", C.c_int), ("toldisp", C.c_double),
                ("tolforc", C.c_double), ("tolener", C.c_double)]


def arclength_static(asm, q, dk, dkdof, alpha, psi_thresh, iteopt, lpfmax, dkimax, itemax, submax,
                     imagmax, negmax, toldisp, tolforc, tolener, max_rows=4096, csc=False):
    """the C host driver cb_arclength_static (main.c:2158-3141 on the device path).
    Returns (hist [rows][2+NEQ], result)."""
    hl = load_host_library()
    m = asm.m
    p = cb_arc_params(dk, dkdof, alpha, psi_thresh, iteopt, lpfmax, dkimax, itemax, submax, imagmax,
                      negmax, toldisp, tolforc, tolener)
    res = cb_nr_result()
    hist = np.zeros((max_rows, m.NEQ + 2))
    q = np.ascontiguousarray(q, dtype=np.float64)
    _keep, maxa_p, lss = _solver_args(m, csc)
    hl.cb_arclength_static(asm.h, C.c_long(m.NEQ), maxa_p, lss, _p(q), C.byref(p), _p(hist),
                           C.c_int(max_rows), C.byref(res))
    return hist[:res.increments].copy(), res
