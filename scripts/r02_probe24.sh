#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sw_trace|sw_mid|sw_reverse_extract|sw_binscan|sw_scatter|sw_init|sw_finalize|sw_cigar|scan_" -c 120 --csv --log-file gpurun_out/p24_launches.csv python scripts/probes/one_step.py > gpurun_out/p24_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/p24_launches.csv')) if len(r)>10 and r[0].isdigit()]
half=len(rows)//2
for r in rows[half:]:
    print(r[4][:70], r[-1], r[-2])
PY
