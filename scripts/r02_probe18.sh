#!/bin/bash
# anti-diagonal-major direction storage of the cooperative traceback: parity, merge shapes, launch list of pe_contigs
set -u
mkdir -p gpurun_out
python -m pytest tests/test_sw_gpu.py tests/test_golden_gpu.py tests/test_callers_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/p18_pytest.txt
cat gpurun_out/p18_pytest.txt
summ() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']; s=r['stage_ms_per_step']
print(sys.argv[2], 'GCUPS', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), {k: round(v,2) for k,v in s.items() if v > 0.005})
" "$1" "$2"; }
for w in pe_contigs long_contigs; do
  python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/p18_$w.json 2> gpurun_out/p18_$w.err; summ gpurun_out/p18_$w.json $w
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sw_trace" -c 200 --csv --log-file gpurun_out/p18_pe_trace_launches.csv python bench.py --workload pe_contigs --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/p18_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/p18_pe_trace_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:40]:
    print(r[4][:60], r[-1], r[-2])
PY
