#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/quick_time_lattice.py 119 > $out/r03k_lattice.log 2>&1; tail -4 $out/r03k_lattice.log
timeout 1500 python -m pytest tests -m gpu -q > $out/r03k_tests.log 2>&1; echo "tests rc=$?" >> $out/r03k_tests.log
tail -4 $out/r03k_tests.log
