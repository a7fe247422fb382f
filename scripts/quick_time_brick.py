"""ad-hoc timing of the brick (+ shell skin) assembly on the BASELINE config-5 shape: stiff_br is
linear and assembled once in the reference; reported as bricks/s, not the bench metric."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cu-bens_b200", "python"))
import numpy as np
import cubens_b200 as cb
from cubens_b200 import meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
t = time.time(); m = meshgen.brick_model(n, n, n, skin=True); print("mesh", round(time.time() - t, 2), "bricks", m.NE_SBR, "skin shells", m.NE_SH, "NEQ", m.NEQ, flush=True)
t = time.time(); a = cb.Assembler(m, layout=cb.CB_MAT_CSC); print("create", round(time.time() - t, 2), flush=True)
t = time.time(); a.stiff(cb.CB_GEN_COMMITTED); a.sync(); print("first stiff (plan build)", round(time.time() - t, 2), "nnz", a.lib.cb_csc_nnz(a.h), flush=True)
for i in range(3):
    a.stiff(cb.CB_GEN_COMMITTED); ks = a.last_stiff_ms
    t0 = time.time(); a.lib.cb_mass(a.h); ms = 1e3 * (time.time() - t0)
    print(f"iter {i}: K {ks:.3f} ms ({m.NE_SBR / ks / 1e3:.1f} M bricks/s)  consistent M on the CSC pattern (wall) {ms:.3f} ms", flush=True)
