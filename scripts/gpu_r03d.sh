#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_shell_forces|k_gather" -s 6 -c 6 --csv --log-file $out/r03d_launches.csv python scripts/kt_compare.py 1000 narrow12 > $out/r03d_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r03d_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
for r in rows[hi+2:]:
    if len(r)>mv: print(r[kn][:40], r[mn], r[mv], r[mu])
PY
timeout 1500 python -m pytest tests -m gpu -q > $out/r03d_tests.log 2>&1; echo "tests rc=$?" >> $out/r03d_tests.log
tail -5 $out/r03d_tests.log
