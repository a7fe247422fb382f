#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/r02l_tests.log 2>&1; echo "tests rc=$?" >> $out/r02l_tests.log
tail -25 $out/r02l_tests.log
timeout 600 python scripts/kt_compare.py 1000 narrow 0.2 > $out/r02l_kt_jit.log 2>&1; cat $out/r02l_kt_jit.log
