"""development aid: K_t kernel variants on the BASELINE plate (CB_KT=duo | stream), CUDA-event times of
the assembly kernel alone, and the largest difference between the two matrices"""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cu-bens_b200", "python"))
import numpy as np
import cubens_b200 as cb
from cubens_b200 import meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["duo", "wide", "narrow"]
jit = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
m = meshgen.plate_model(n, n, SLVFLAG=2, jitter=jit)
dd = meshgen.perturbation(m)
ref = None
for kind in kinds:
    os.environ["CB_KT"] = kind
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    a.begin_increment()
    t = time.time(); a.stiff(); a.sync(); tp = time.time() - t
    a.update_forces(dd, want_f=False); a.end_iteration()
    ks, fs = [], []
    for i in range(12):
        a.stiff(); ks.append(a.last_assemble_ms)
        a.update_forces(dd * 0.01, want_f=False); fs.append(a.last_forces_ms)
        a.end_iteration()
    a.stiff()
    Ax = a.csc_values()
    print(f"{kind:7s} jitter {jit}: classes {a.geometry_classes} first stiff {tp:.2f} s  K_t {np.median(ks):.4f} ms "
          f"(min {min(ks):.4f})  f_int {np.median(fs):.4f} ms  map MB {a.map_bytes / 1e6:.1f}", flush=True)
    if ref is None:
        ref = Ax
    else:
        print("   max |Ax - Ax_first| / max |Ax| =", np.abs(Ax - ref).max() / np.abs(ref).max(), flush=True)
    a.close()
