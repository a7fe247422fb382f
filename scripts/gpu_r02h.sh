#!/bin/bash
out=gpurun_out; mkdir -p $out
nproc; free -g | head -2
timeout 900 python -m pytest tests/test_shell_gpu.py tests/test_partition_gpu.py -m gpu -x -q > $out/r02h_tests.log 2>&1; echo "tests rc=$?" >> $out/r02h_tests.log
tail -4 $out/r02h_tests.log
timeout 900 python bench.py --steps 30 --no-others --no-unstructured --cpu-steps 10 > $out/r02h_bench.json 2> $out/r02h_bench.err; tail -3 $out/r02h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','split_ms','gpu_launches')})
print(d['e2e'])
print(d.get('cpu_baseline'))
PY
