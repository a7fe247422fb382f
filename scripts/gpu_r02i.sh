#!/bin/bash
out=gpurun_out; mkdir -p $out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r02i_tests.log 2>&1; echo "tests rc=$?" >> $out/r02i_tests.log
tail -6 $out/r02i_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 3 > $out/r02i_bench_n2.json 2> $out/r02i_bench_n2.err; tail -3 $out/r02i_bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02i_bench_n2.json'))
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','setup_s')})
    print('e2e', {k:v for k,v in d['e2e'].items() if k!='note'})
    print('strong', d.get('strong'))
except Exception as e:
    print('bench parse failed', e)
PY
