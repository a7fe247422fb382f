#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_midsize_gpu.py tests/test_shell_gpu.py tests/test_c_host_multigpu.py -m gpu -q > $out/r02o_tests.log 2>&1; echo "tests rc=$?" >> $out/r02o_tests.log
tail -5 $out/r02o_tests.log
timeout 600 python scripts/kt_compare.py 1000 narrow 0.2 > $out/r02o_kt_jit.log 2>&1; cat $out/r02o_kt_jit.log
timeout 600 python scripts/kt_compare.py 1000 narrow > $out/r02o_kt.log 2>&1; cat $out/r02o_kt.log
