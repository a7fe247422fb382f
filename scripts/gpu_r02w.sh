#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiles -s 1 -c 1 \
    -o $out/r02w_prof_brick python scripts/quick_time_brick.py > $out/r02w_ncu.log 2>&1; tail -3 $out/r02w_ncu.log
