#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/kt_compare.py 1000 narrow12 > $out/r03c_kt.log 2>&1
CB_NO_WARP_SUMS=1 timeout 600 python scripts/kt_compare.py 1000 narrow12 >> $out/r03c_kt.log 2>&1
timeout 600 python scripts/kt_compare.py 1000 narrow12 0.2 >> $out/r03c_kt.log 2>&1; cat $out/r03c_kt.log
timeout 1500 python -m pytest tests -m gpu -q -x > $out/r03c_tests.log 2>&1; echo "tests rc=$?" >> $out/r03c_tests.log
tail -8 $out/r03c_tests.log
