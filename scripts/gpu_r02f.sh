#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_midsize_gpu.py tests/test_shell_gpu.py -m gpu -x -q > $out/r02f_tests.log 2>&1; echo "tests rc=$?" >> $out/r02f_tests.log
tail -3 $out/r02f_tests.log
echo "== default (8 warps, krec prefetch)"; timeout 600 python scripts/kt_compare.py 1000 wide,narrow 2>&1 | grep -v Ax_first
echo "== 10 warps, no prefetch"; CUBENS_LIB=cu-bens_b200/variants/libcubens_w10p0.so timeout 600 python scripts/kt_compare.py 1000 narrow 2>&1 | grep -v Ax_first
echo "== 8 warps, no prefetch"; CUBENS_LIB=cu-bens_b200/variants/libcubens_w8p0.so timeout 600 python scripts/kt_compare.py 1000 wide,narrow 2>&1 | grep -v Ax_first
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_shell_stream -s 3 -c 1 \
    -o $out/r02f_prof_narrow python scripts/kt_compare.py 1000 narrow > $out/r02f_ncu8.log 2>&1; tail -2 $out/r02f_ncu8.log
