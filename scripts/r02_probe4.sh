#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_small_path_gpu.py tests/test_ed_gpu.py tests/test_golden_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_new_3.txt
cat gpurun_out/r02_pytest_new_3.txt
timeout 600 python scripts/latency_table.py > gpurun_out/r02_latency_2.json 2> gpurun_out/r02_latency_2.err
cat gpurun_out/r02_latency_2.json | head -70; tail -5 gpurun_out/r02_latency_2.err
for w in hifi_nw ont_nw; do timeout 300 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r02_${w}_1.json 2> gpurun_out/r02_${w}_1.err; python - <<PY
import json
d=json.load(open('gpurun_out/r02_${w}_1.json')); r=d['roofline']
print('$w', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'stages', {k: round(v,2) for k,v in r['stage_ms_per_step'].items()}, 'mism', d['cpu_baseline']['parity_mismatches_on_sample'])
PY
done
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1.json 2> gpurun_out/r02_bench_1.err; cat gpurun_out/r02_bench_1.json | head -c 1500
