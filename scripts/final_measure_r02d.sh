#!/bin/bash
# Round 2, last kernel change (running maximum as one chain, cheaper query decode): SW tests, parity sweep, bench lines, SW workloads
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sw_gpu.py tests/test_golden_gpu.py tests/test_input_forms_gpu.py tests/test_small_path_gpu.py tests/test_callers_gpu.py tests/test_patch_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/find_pytest_sw.txt; cat gpurun_out/find_pytest_sw.txt
python scripts/parity_sweep.py --pairs 100000 --max-len 600 --seed 31 2>&1 | grep "^SW\|TOTAL" > gpurun_out/find_parity_sweep.txt; grep TOTAL gpurun_out/find_parity_sweep.txt
python scripts/parity_sweep.py --pairs 6000 --max-len 6000 --seed 32 2>&1 | grep "^SW\|TOTAL" >> gpurun_out/find_parity_sweep.txt; tail -1 gpurun_out/find_parity_sweep.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/find_bench_1gpu.json 2> gpurun_out/find_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/find_bench_reference_arm.json 2>> gpurun_out/find_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/find_bench_1gpu.json'))
print('headline', round(d['value']), 'e2e', round(d['e2e']['value']), 'single', round(d['e2e_single_call']['value']), 'ascii', round(d['e2e_ascii']['value']), 'front', round(d['python_front_door']['value']), 'ms', round(d['ms_per_step'],2), 'roof', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['frac_whole_step'],3), 'merge', round(d['merge_shape']['value']), round(d['merge_shape']['e2e']['value']), d['clocks'], 'launches', d['gpu_launches'], 'cpu', round(d['cpu_baseline']['value'],1), d['cpu_baseline']['parity_mismatches_on_sample'], d['cpu_baseline']['cigar_mismatches_on_sample'])
print({k: round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items() if v > 0.005})
r=json.load(open('gpurun_out/find_bench_reference_arm.json')); print('reference arm', round(r['value'],1), r['cpu_baseline']['cores'])
"
STEPS=4 WLS="pe_contigs long_contigs" bash scripts/bench_workloads.sh > gpurun_out/find_workloads_summary.txt 2>&1
cp gpurun_out/bench_workloads.jsonl gpurun_out/find_bench_workloads_sw.jsonl; cat gpurun_out/find_workloads_summary.txt
python scripts/probes/cfg_cost.py 16,1 16,2 16,4 13,4 18,8 10,16 > gpurun_out/find_cfg_cost.txt 2>&1; cat gpurun_out/find_cfg_cost.txt
