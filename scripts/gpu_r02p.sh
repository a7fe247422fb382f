#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/kt_compare.py 1000 narrow,mini > $out/r02p_kt.log 2>&1; cat $out/r02p_kt.log
timeout 600 python scripts/kt_compare.py 1000 narrow,mini 0.2 > $out/r02p_kt_jit.log 2>&1; cat $out/r02p_kt_jit.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_shell_stream -s 3 -c 1 \
    -o $out/r02p_prof_mini python scripts/kt_compare.py 1000 mini > $out/r02p_ncu.log 2>&1; tail -2 $out/r02p_ncu.log
