#!/bin/bash
# compute-sanitizer on a small walk through the benchmarked kernels (stream K_t, fused force pass with warp sums,
# fused sums): memcheck and racecheck
out=gpurun_out; mkdir -p $out
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "cu-bens_b200", "python"))
import numpy as np
import cubens_b200 as cb
from cubens_b200 import meshgen
for kw in (dict(), dict(jitter=0.2), dict(unionjack=True, z_bump=0.02)):
    m = meshgen.plate_model(70, 45, SLVFLAG=2, **kw)
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    a.set_q(m.q)
    a.begin_increment()
    dd = meshgen.perturbation(m, scale=1e-4)
    for it in range(2):
        a.stiff(); a.update_forces(dd, itecnt=it); a.residual_sums_allreduce(0.5); a.end_iteration()
    Ax = a.csc_values(); s = a.residual_sums(0.5, fetch=True)
    print(kw, a.geometry_classes, float(np.abs(Ax).max()), s[:2])
    a.close()
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san_case.py > $out/r02f_memcheck.log 2>&1; echo "memcheck rc=$?" >> $out/r02f_memcheck.log
tail -6 $out/r02f_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python /tmp/san_case.py > $out/r02f_racecheck.log 2>&1; echo "racecheck rc=$?" >> $out/r02f_racecheck.log
tail -6 $out/r02f_racecheck.log
