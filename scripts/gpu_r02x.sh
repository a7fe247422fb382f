#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -k "brick or fsi or mixed or mass or truss" > $out/r02x_tests.log 2>&1; echo "tests rc=$?" >> $out/r02x_tests.log
tail -5 $out/r02x_tests.log
timeout 600 python scripts/quick_time_brick.py > $out/r02x_brick.log 2>&1; tail -3 $out/r02x_brick.log
python __graft_entry__.py smoke > $out/r02x_smoke.log 2>&1; tail -3 $out/r02x_smoke.log
