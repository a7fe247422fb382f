#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shell_forces -s 3 -c 1 \
    -o $out/r02n_prof_forces_jit python scripts/kt_compare.py 1000 narrow 0.2 > $out/r02n_ncu.log 2>&1; tail -2 $out/r02n_ncu.log
