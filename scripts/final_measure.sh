#!/bin/bash
# One GPU box: GPU test suite, randomised parity sweeps against oracle/_ref, the headline bench line (+ reference arm),
# every named workload through bench.py.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_pytest_gpu.txt
{ echo "# python scripts/parity_sweep.py --pairs 100000 --max-len 600 --seed 3"; python scripts/parity_sweep.py --pairs 100000 --max-len 600 --seed 3;
  echo; echo "# python scripts/parity_sweep.py --pairs 15000 --max-len 4000 --seed 4"; python scripts/parity_sweep.py --pairs 15000 --max-len 4000 --seed 4; } > gpurun_out/final_parity_sweep.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2>> gpurun_out/final_bench_1gpu.err
STEPS=5 scripts/bench_workloads.sh > gpurun_out/final_workloads_summary.txt 2>&1
cp gpurun_out/bench_workloads.jsonl gpurun_out/final_bench_workloads.jsonl
cat gpurun_out/final_pytest_gpu.txt; grep TOTAL gpurun_out/final_parity_sweep.txt; cat gpurun_out/final_workloads_summary.txt
python -c "
import json
d=json.load(open('gpurun_out/final_bench_1gpu.json'))
print('headline', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'roof', round(d['roofline']['frac'],3), d['clocks'])
r=json.load(open('gpurun_out/final_bench_reference_arm.json')); print('reference arm', round(r['value'],1), r['cpu_baseline']['cores'])
"
