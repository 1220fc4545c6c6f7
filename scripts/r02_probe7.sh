#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_minimizers_gpu.py tests/test_concurrency_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_new_5.txt
cat gpurun_out/r02_pytest_new_5.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_3.json 2> gpurun_out/r02_bench_3.err; tail -3 gpurun_out/r02_bench_3.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_3.json'))
for k in ('e2e', 'e2e_single_call', 'e2e_ascii', 'python_front_door'):
    print(k, d[k] and round(d[k]['value']), d[k] and d[k].get('ms_per_step'))
print('value', round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_whole_step'])
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ed_myers_kernel -c 14 -o gpurun_out/r02_ed_myers_full -f python bench.py --workload hifi_nw --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_ed2.log 2>&1
tail -2 gpurun_out/r02_ncu_ed2.log
