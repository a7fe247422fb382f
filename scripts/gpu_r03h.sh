#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/kt_compare.py 1000 narrow,narrow12 0.2 > $out/r03h_kt.log 2>&1
CB_KEBC_STREAM=1 timeout 600 python scripts/kt_compare.py 1000 narrow12 0.2 >> $out/r03h_kt.log 2>&1; cat $out/r03h_kt.log
timeout 900 python -m pytest tests/test_midsize_gpu.py tests/test_fullsize_gpu.py tests/test_shell_gpu.py tests/test_plan_device_gpu.py -m gpu -q > $out/r03h_tests.log 2>&1; echo "tests rc=$?" >> $out/r03h_tests.log
tail -5 $out/r03h_tests.log
