"""ad-hoc timing of the frame path on the BASELINE config-4 shape (cubic frame lattice)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cu-bens_b200", "python"))
import numpy as np
import cubens_b200 as cb
from cubens_b200 import meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
t = time.time(); m = meshgen.lattice_model(n, SLVFLAG=2); print("mesh", round(time.time() - t, 2), "frames", m.NE_FR, "NEQ", m.NEQ, flush=True)
t = time.time(); a = cb.Assembler(m, layout=cb.CB_MAT_CSC); print("create", round(time.time() - t, 2), flush=True)
a.begin_increment()
t = time.time(); a.stiff(); a.sync(); print("first stiff (plan build)", round(time.time() - t, 2), "nnz", a.lib.cb_csc_nnz(a.h), flush=True)
dd = meshgen.perturbation(m, scale=1e-3)
a.update_forces(dd, want_f=False); a.end_iteration()
for i in range(4):
    a.stiff(); ks = a.last_stiff_ms
    a.update_forces(dd * 0.01, want_f=False); fs = a.last_forces_ms
    a.end_iteration()
    t0 = time.time(); a.mass(); ms = 1e3 * (time.time() - t0)
    print(f"iter {i}: stiff {ks:.3f} ms  forces {fs:.3f} ms  mass(wall) {ms:.3f} ms -> {m.NE_FR / (ks + fs) / 1e3:.1f} M frames/s", flush=True)
