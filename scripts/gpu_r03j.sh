#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/kt_compare.py 1000 narrow12 > $out/r03j_kt.log 2>&1
timeout 600 python scripts/kt_compare.py 1000 narrow12 0.2 >> $out/r03j_kt.log 2>&1; cat $out/r03j_kt.log
timeout 1500 python -m pytest tests -m gpu -q > $out/r03j_tests.log 2>&1; echo "tests rc=$?" >> $out/r03j_tests.log
tail -4 $out/r03j_tests.log
