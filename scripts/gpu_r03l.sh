#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiles -s 2 -c 1 \
    -o $out/r03l_prof_frame_tiles python scripts/quick_time_lattice.py 100 > $out/r03l_ncu.log 2>&1; tail -2 $out/r03l_ncu.log
