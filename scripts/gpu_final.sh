#!/bin/bash
out=gpurun_out; mkdir -p $out
git_rev=$(cat .git_rev 2>/dev/null)
timeout 1800 python -m pytest tests -m gpu -q > $out/final_tests.log 2>&1; echo "tests rc=$?" >> $out/final_tests.log
tail -4 $out/final_tests.log
python __graft_entry__.py smoke > $out/final_smoke.log 2>&1; tail -2 $out/final_smoke.log
python bench.py > $out/final_bench_n1.json 2> $out/final_bench_n1.err; head -c 700 $out/final_bench_n1.json
