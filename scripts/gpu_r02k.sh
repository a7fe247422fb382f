#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_plan_device_gpu.py -m gpu -x -q > $out/r02k_tests.log 2>&1; echo "tests rc=$?" >> $out/r02k_tests.log
tail -30 $out/r02k_tests.log
timeout 600 python - <<'PY'
import sys, os, time
sys.path.insert(0, "cu-bens_b200/python")
import cubens_b200 as cb
from cubens_b200 import meshgen
m = meshgen.plate_model(1000, 1000, SLVFLAG=2)
for mode in ("device", "host"):
    os.environ["CB_PLAN"] = mode
    t = time.time(); a = cb.Assembler(m, layout=cb.CB_MAT_CSC); a.sync(); tc = time.time() - t
    t = time.time(); a.begin_increment(); a.stiff(); a.sync(); ts = time.time() - t
    print(mode, "create %.2f s  first stiff %.2f s  plan_info" % (tc, ts), a.plan_info(), "K_t ms", a.last_assemble_ms, flush=True)
    for _ in range(3):
        a.stiff()
    print("   K_t", a.last_assemble_ms)
    a.close()
PY
