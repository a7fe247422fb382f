#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q > $out/r02g_tests.log 2>&1; echo "tests rc=$?" >> $out/r02g_tests.log
tail -4 $out/r02g_tests.log
python __graft_entry__.py smoke > $out/r02g_smoke.log 2>&1; tail -2 $out/r02g_smoke.log
bash scripts/capture_profiles.sh r02g > $out/r02g_capture.log 2>&1
head -c 1500 $out/bench_r02g_n1.json
