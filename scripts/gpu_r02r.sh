#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_shell_stream -s 3 -c 1 \
    -o $out/r02r_prof_narrow12 python scripts/kt_compare.py 1000 narrow12 > $out/r02r_ncu.log 2>&1; tail -2 $out/r02r_ncu.log
timeout 1500 python -m pytest tests -m gpu -q > $out/r02r_tests.log 2>&1; echo "tests rc=$?" >> $out/r02r_tests.log
tail -5 $out/r02r_tests.log
