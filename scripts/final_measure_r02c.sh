#!/bin/bash
# Round 2, final artefacts after the fine-configuration threshold change: SW tests, bench lines as JSON, workloads as JSON lines,
# a full ncu capture of the largest forward buckets (report deleted after the text export: 64 MB travel back), DRAM traffic.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sw_gpu.py tests/test_golden_gpu.py tests/test_input_forms_gpu.py tests/test_small_path_gpu.py tests/test_queue_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/finc_pytest_sw.txt; cat gpurun_out/finc_pytest_sw.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/finc_bench_1gpu.json 2> gpurun_out/finc_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/finc_bench_reference_arm.json 2>> gpurun_out/finc_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/finc_bench_1gpu.json'))
print('headline', round(d['value']), 'e2e', round(d['e2e']['value']), 'single', round(d['e2e_single_call']['value']), 'ascii', round(d['e2e_ascii']['value']), 'front', round(d['python_front_door']['value']), 'ms', round(d['ms_per_step'],2), 'roof', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['frac_whole_step'],3), 'merge', round(d['merge_shape']['value']), round(d['merge_shape']['e2e']['value']), d['clocks'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline'])
print({k: round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items() if v > 0.005})
r=json.load(open('gpurun_out/finc_bench_reference_arm.json')); print('reference arm', round(r['value'],1), r['cpu_baseline']['cores'])
"
STEPS=4 WLS="pe_contigs long_contigs remap_chain hifi_nw ont_nw hifi_hw remap_path hifi_path" bash scripts/bench_workloads.sh > gpurun_out/finc_workloads_summary.txt 2>&1
cp gpurun_out/bench_workloads.jsonl gpurun_out/finc_bench_workloads.jsonl; cat gpurun_out/finc_workloads_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_pass_kernel --launch-skip 396 -c 14 -o gpurun_out/finc_fwd_kernels -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-merge-shape --e2e-steps 1 > gpurun_out/finc_ncu_fwd.log 2>&1; tail -2 gpurun_out/finc_ncu_fwd.log
ncu -i gpurun_out/finc_fwd_kernels.ncu-rep --page details > gpurun_out/finc_fwd_kernels_details.txt 2>/dev/null
ncu -i gpurun_out/finc_fwd_kernels.ncu-rep --page raw --csv > gpurun_out/finc_fwd_kernels_raw.csv 2>/dev/null
rm -f gpurun_out/finc_fwd_kernels.ncu-rep
bash scripts/forward_traffic.sh 2>&1 | tail -2
rm -f gpurun_out/r02_fwd_traffic.csv
du -sh gpurun_out
