#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_shell_gpu.py -m gpu -q > $out/r03e_tests.log 2>&1; echo "tests rc=$?" >> $out/r03e_tests.log
tail -12 $out/r03e_tests.log
