"""ad-hoc timing of the two hot passes on the BASELINE plate (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cu-bens_b200", "python"))
import numpy as np
import cubens_b200 as cb
from cubens_b200 import meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t = time.time(); m = meshgen.plate_model(n, n, SLVFLAG=2); print("mesh", time.time() - t, m.NE_SH, m.NEQ, flush=True)
t = time.time(); a = cb.Assembler(m, layout=cb.CB_MAT_CSC); print("create", time.time() - t, flush=True)
a.begin_increment()
t = time.time(); a.stiff(); print("first stiff (plan build)", time.time() - t, "nnz", a.lib.cb_csc_nnz(a.h), "map MB", a.map_bytes / 1e6, flush=True)
dd = meshgen.perturbation(m)
a.update_forces(dd, want_f=False); a.end_iteration()
for i in range(5):
    a.stiff(); ks = a.last_stiff_ms
    a.update_forces(dd * 0.01, want_f=False); fs = a.last_forces_ms
    a.end_iteration()
    print(f"iter {i}: stiff {ks:.3f} ms  forces {fs:.3f} ms  -> {m.NE_SH / (ks + fs) / 1e3:.1f} M el/s", flush=True)
