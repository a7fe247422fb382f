#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_midsize_gpu.py tests/test_shell_gpu.py -m gpu -x -q > $out/r02d_tests.log 2>&1; echo "tests rc=$?" >> $out/r02d_tests.log
tail -5 $out/r02d_tests.log
timeout 600 python scripts/kt_compare.py 1000 duo,wide,narrow,narrow8 > $out/r02d_kt.log 2>&1; cat $out/r02d_kt.log
timeout 600 python scripts/kt_compare.py 1000 wide,narrow8 0.2 > $out/r02d_kt_jit.log 2>&1; cat $out/r02d_kt_jit.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_shell_stream -s 3 -c 1 \
    -o $out/r02d_prof_wide python scripts/kt_compare.py 1000 wide > $out/r02d_ncu.log 2>&1; tail -2 $out/r02d_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_shell_stream -s 3 -c 1 \
    -o $out/r02d_prof_narrow8 python scripts/kt_compare.py 1000 narrow8 > $out/r02d_ncu8.log 2>&1; tail -2 $out/r02d_ncu8.log
