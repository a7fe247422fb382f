#!/bin/bash
out=gpurun_out; mkdir -p $out
for v in "" build/libcubens_n2.so; do
  if [ -n "$v" ]; then export CUBENS_LIB=$PWD/cu-bens_b200/$v; fi
  echo "== lib ${v:-default}" >> $out/r03m_kt.log
  timeout 600 python scripts/kt_compare.py 1000 narrow12 0.2 >> $out/r03m_kt.log 2>&1
done
cat $out/r03m_kt.log
