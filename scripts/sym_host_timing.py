"""development aid: wall time of the host-side rebuild of the full CSC matrix from the packed upper triangle
(cb_sym_selftest, no device) on the BASELINE plate, for several thread counts"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cu-bens_b200", "python"))
import cubens_b200 as cb
from cubens_b200 import meshgen
m = meshgen.plate_model(1000, 1000, SLVFLAG=2)
for nt in (1, 2, 4, 8, 16, 32):
    if nt > 2 * (os.cpu_count() or 1):
        break
    t = min(cb.sym_selftest(m, 0, 0, nt) for _ in range(2))
    print(f"{nt:3d} threads: {t * 1e3:7.1f} ms  -> {2.06 / t:6.1f} GB/s of matrix written", flush=True)
