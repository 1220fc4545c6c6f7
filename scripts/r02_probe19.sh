#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sw_trace" -c 60 --csv --log-file gpurun_out/p19_long_trace_launches.csv python bench.py --workload long_contigs --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/p19_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/p19_long_trace_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:30]:
    print(r[4][:60], r[-1], r[-2])
PY
