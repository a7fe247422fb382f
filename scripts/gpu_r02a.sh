#!/bin/bash
# round-2 first GPU call: tests, K_t variant timing, one full ncu capture of the stream kernel
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r02a_tests.log 2>&1; echo "tests rc=$?" >> $out/r02a_tests.log
tail -5 $out/r02a_tests.log
timeout 600 python scripts/kt_compare.py 1000 duo,stream > $out/r02a_kt.log 2>&1; cat $out/r02a_kt.log
timeout 600 python scripts/kt_compare.py 1000 duo,stream 0.2 > $out/r02a_kt_jit.log 2>&1; cat $out/r02a_kt_jit.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_shell_stream -s 3 -c 1 \
    -o $out/r02a_prof_stream python scripts/kt_compare.py 1000 stream > $out/r02a_ncu.log 2>&1; tail -3 $out/r02a_ncu.log
