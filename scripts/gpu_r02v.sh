#!/bin/bash
# N = $1 GPUs: bench line (weak + strong + e2e) and, at N=2, the C-host NCCL test
N=$1
out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_c_host_multigpu.py tests/test_partition_gpu.py -m gpu -q > $out/r02v_tests_n2.log 2>&1; echo "tests rc=$?" >> $out/r02v_tests_n2.log
  tail -4 $out/r02v_tests_n2.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 3 > $out/r02v_bench_n$N.json 2> $out/r02v_bench_n$N.err; tail -3 $out/r02v_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02v_bench_n$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')})
    print('e2e', {k:v for k,v in d['e2e'].items() if k!='note'})
    print('strong', d.get('strong'))
except Exception as e:
    print('bench parse failed', e)
PY
