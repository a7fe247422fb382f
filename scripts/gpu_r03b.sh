#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/kt_compare.py 1000 narrow,narrow12,mini > $out/r03b_kt.log 2>&1; cat $out/r03b_kt.log
timeout 900 python -m pytest tests/test_midsize_gpu.py tests/test_fullsize_gpu.py tests/test_shell_gpu.py -m gpu -q > $out/r03b_tests.log 2>&1; echo "tests rc=$?" >> $out/r03b_tests.log
tail -5 $out/r03b_tests.log
