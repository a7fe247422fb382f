#!/bin/bash
out=gpurun_out; mkdir -p $out
for v in "" build/libcubens_fc5.so build/libcubens_fc6.so; do
  if [ -n "$v" ]; then export CUBENS_LIB=$PWD/cu-bens_b200/$v; fi
  echo "== lib ${v:-default}" >> $out/r03f_kt.log
  timeout 600 python scripts/kt_compare.py 1000 narrow12 >> $out/r03f_kt.log 2>&1
done
cat $out/r03f_kt.log
