#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/kt_compare.py 1000 narrow,narrow12 > $out/r02t_kt.log 2>&1; cat $out/r02t_kt.log
timeout 600 python scripts/kt_compare.py 1000 narrow12 0.2 > $out/r02t_kt_jit.log 2>&1; cat $out/r02t_kt_jit.log
timeout 900 python -m pytest tests/test_midsize_gpu.py tests/test_fullsize_gpu.py tests/test_partition_gpu.py -m gpu -q > $out/r02t_tests.log 2>&1; echo "tests rc=$?" >> $out/r02t_tests.log
tail -5 $out/r02t_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_shell_stream -s 3 -c 1 \
    -o $out/r02t_prof_narrow12 python scripts/kt_compare.py 1000 narrow12 > $out/r02t_ncu.log 2>&1; tail -2 $out/r02t_ncu.log
