"""development aid: end-to-end step (stiff + matrix to the host + update_forces with host dd / f_temp) on the
BASELINE plate for several settings of the hybrid symmetric transfer (CB_SYM_FULL_EVERY)"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cu-bens_b200", "python"))
import numpy as np
import cubens_b200 as cb
from cubens_b200 import meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
settings = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "3", "4", "5", "7"]
threads = int(sys.argv[3]) if len(sys.argv) > 3 else min(32, os.cpu_count() or 1)
m = meshgen.plate_model(n, n, SLVFLAG=2)
dd = meshgen.perturbation(m)
for k in settings:
    os.environ["CB_SYM_FULL_EVERY"] = k
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    lib = a.lib
    a.begin_increment(); a.update_forces(dd, want_f=False); a.end_iteration(); a.stiff()
    nnz = lib.cb_csc_nnz(a.h); nu = lib.cb_csc_upper_nnz(a.h)
    hAx = a.pinned(nnz); hAxu = a.pinned(nu); hdd = a.pinned(m.NEQ); hf = a.pinned(m.NEQ)
    hdd[:] = dd * 1e-3

    def forces():
        cdl = C.c_double(1.0); fr = C.c_int(0); sh = C.c_int(0)
        assert lib.cb_update_forces(a.h, C.c_void_p(hdd.ctypes.data), C.byref(cdl), C.c_int(0), C.c_void_p(hf.ctypes.data),
                                    C.byref(fr), C.byref(sh)) == 0

    def mirror():
        a.stiff()
        assert lib.cb_csc_values_begin(a.h, C.c_void_p(hAx.ctypes.data), C.c_void_p(hAxu.ctypes.data), C.c_int(threads)) == 0
        forces()
        assert lib.cb_csc_values_end(a.h) == 0
        a.end_iteration()

    def plain():
        a.stiff()
        assert lib.cb_get_csc_values(a.h, C.c_void_p(hAx.ctypes.data)) == 0
        forces(); a.end_iteration()

    def upper():
        a.stiff()
        assert lib.cb_get_csc_upper_values(a.h, C.c_void_p(hAxu.ctypes.data)) == 0
        forces(); a.end_iteration()

    out = []
    for fn in (mirror, plain, upper):
        fn(); a.sync()
        t0 = time.perf_counter()
        for _ in range(8):
            fn()
        a.sync()
        out.append(1e3 * (time.perf_counter() - t0) / 8)
    # the rebuilt matrix against the plain copy
    mirror(); a.stiff(); full = np.array(hAx, copy=True)
    ref = a.csc_values()
    err = np.abs(full - ref).max() / np.abs(ref).max()
    print(f"full_every {k:>2s} threads {threads}: mirrored {out[0]:6.2f} ms  plain {out[1]:6.2f} ms  upper-only {out[2]:6.2f} ms"
          f"   max diff vs plain {err:.1e}", flush=True)
    a.close()
