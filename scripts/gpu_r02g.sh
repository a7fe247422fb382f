#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r02g_tests.log 2>&1; echo "tests rc=$?" >> $out/r02g_tests.log
tail -5 $out/r02g_tests.log
timeout 600 python scripts/kt_compare.py 1000 duo,narrow 2>&1 | grep -v Ax_first
CB_NO_FUSED_NODE_UPDATE=1 timeout 600 python scripts/kt_compare.py 1000 narrow 2>&1 | grep -v Ax_first
