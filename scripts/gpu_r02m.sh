#!/bin/bash
out=gpurun_out; mkdir -p $out
for v in "" build/libcubens_f3.so build/libcubens_f4.so; do
  if [ -n "$v" ]; then export CUBENS_LIB=$PWD/cu-bens_b200/$v; fi
  echo "== lib ${v:-default}" >> $out/r02m_kt_jit.log
  timeout 600 python scripts/kt_compare.py 1000 narrow 0.2 >> $out/r02m_kt_jit.log 2>&1
done
unset CUBENS_LIB
cat $out/r02m_kt_jit.log
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_midsize_gpu.py tests/test_c_host_multigpu.py -m gpu -q > $out/r02m_tests.log 2>&1; echo "tests rc=$?" >> $out/r02m_tests.log
tail -5 $out/r02m_tests.log
