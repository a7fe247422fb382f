#!/bin/bash
out=gpurun_out; mkdir -p $out
for v in "" build/libcubens_ft32.so build/libcubens_ft64.so; do
  if [ -n "$v" ]; then export CUBENS_LIB=$PWD/cu-bens_b200/$v; fi
  echo "== lib ${v:-default}" >> $out/r03i_kt.log
  timeout 600 python scripts/kt_compare.py 1000 narrow12 >> $out/r03i_kt.log 2>&1
  timeout 600 python scripts/kt_compare.py 1000 narrow12 0.2 >> $out/r03i_kt.log 2>&1
done
cat $out/r03i_kt.log
