#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_4.json 2> gpurun_out/r02_bench_4.err; tail -3 gpurun_out/r02_bench_4.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_4.json'))
for k in ('e2e', 'e2e_single_call', 'e2e_ascii', 'python_front_door'):
    print(k, d[k] and round(d[k]['value']), d[k] and d[k].get('ms_per_step'))
print('value', round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_whole_step'], d['roofline']['stage_ms_per_step'])
print('merge', d['merge_shape'])
print('cpu', d['cpu_baseline'], d['clocks'], d['gpu_launches'])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cat gpurun_out/r02_bench_ref.json | head -c 600; echo
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu_4.txt; cat gpurun_out/r02_pytest_gpu_4.txt
