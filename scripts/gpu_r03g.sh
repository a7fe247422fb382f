#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_assemble_shell_stream|k_shell_forces|k_gather" -s 6 -c 3 --csv --log-file $out/r03g_jit.csv python scripts/kt_compare.py 1000 narrow12 0.2 > $out/r03g_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r03g_jit.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
for r in rows[hi+2:]:
    if len(r)>mv: print(r[kn][:34], r[mn], r[mv], r[mu])
PY
