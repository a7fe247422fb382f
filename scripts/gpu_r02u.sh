#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/kt_compare.py 1000 narrow,narrow12 > $out/r02u_kt.log 2>&1; cat $out/r02u_kt.log
timeout 600 python scripts/kt_compare.py 1000 narrow12 0.2 > $out/r02u_kt_jit.log 2>&1; cat $out/r02u_kt_jit.log
