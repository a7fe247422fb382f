#!/bin/bash
out=gpurun_out; mkdir -p $out
CB_EXPECT_PEER_MEMORY=1 timeout 600 python -m pytest tests/test_c_host_multigpu.py tests/test_partition_gpu.py -m gpu -q -s > $out/r02z_tests_n2.log 2>&1; echo "tests rc=$?" >> $out/r02z_tests_n2.log
grep -E "rank|passed|failed|rc=|Error|error" $out/r02z_tests_n2.log | tail -14
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 3 --no-others > $out/r02z_bench_n2.json 2> $out/r02z_bench_n2.err; tail -2 $out/r02z_bench_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02z_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')})
print('strong', {k:v for k,v in d['strong'].items() if k not in ('note','workload')})
PY
