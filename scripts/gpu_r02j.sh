#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_reference_driver_gpu.py tests/test_shell_gpu.py -m gpu -q > $out/r02j_tests.log 2>&1; echo "tests rc=$?" >> $out/r02j_tests.log
tail -40 $out/r02j_tests.log
