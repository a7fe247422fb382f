#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/kt_compare.py 1000 duo,narrow12,narrow,mini,wide > $out/r02s_kt.log 2>&1; cat $out/r02s_kt.log
timeout 600 python scripts/kt_compare.py 1000 narrow12 0.2 > $out/r02s_kt_jit.log 2>&1; cat $out/r02s_kt_jit.log
timeout 900 python -m pytest tests/test_midsize_gpu.py tests/test_fullsize_gpu.py -m gpu -q > $out/r02s_tests.log 2>&1; echo "tests rc=$?" >> $out/r02s_tests.log
tail -5 $out/r02s_tests.log
