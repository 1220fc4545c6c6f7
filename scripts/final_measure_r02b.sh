#!/bin/bash
# Round 2, second half: GPU test suite, randomised parity sweeps against oracle/_ref, the headline bench line (+ reference arm),
# every named workload through bench.py, launch list / full capture / DRAM traffic of the scoring kernel, per-configuration
# cost table.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/fin_pytest_gpu.txt; cat gpurun_out/fin_pytest_gpu.txt
{ echo "# python scripts/parity_sweep.py --pairs 100000 --max-len 600 --seed 21"; python scripts/parity_sweep.py --pairs 100000 --max-len 600 --seed 21;
  echo; echo "# python scripts/parity_sweep.py --pairs 15000 --max-len 4000 --seed 22"; python scripts/parity_sweep.py --pairs 15000 --max-len 4000 --seed 22; } > gpurun_out/fin_parity_sweep.txt 2>&1
grep TOTAL gpurun_out/fin_parity_sweep.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/fin_bench_1gpu.json 2> gpurun_out/fin_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_bench_reference_arm.json 2>> gpurun_out/fin_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/fin_bench_1gpu.json'))
print('headline', round(d['value']), 'e2e', round(d['e2e']['value']), 'single', round(d['e2e_single_call']['value']), 'ascii', round(d['e2e_ascii']['value']), 'ms', round(d['ms_per_step'],2), 'roof', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['frac_whole_step'],3), 'merge', round(d['merge_shape']['value']), d['clocks'], 'launches', d['gpu_launches'])
print({k: round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items() if v > 0.005})
r=json.load(open('gpurun_out/fin_bench_reference_arm.json')); print('reference arm', round(r['value'],1), r['cpu_baseline']['cores'])
"
STEPS=4 WLS="pe_contigs long_contigs remap_chain hifi_nw ont_nw hifi_hw remap_path hifi_path" bash scripts/bench_workloads.sh > gpurun_out/fin_workloads_summary.txt 2>&1
cp gpurun_out/bench_workloads.jsonl gpurun_out/fin_bench_workloads.jsonl; cat gpurun_out/fin_workloads_summary.txt
python scripts/probes/cfg_cost.py > gpurun_out/fin_cfg_cost.txt 2>&1; tail -30 gpurun_out/fin_cfg_cost.txt
# profiles (each ncu pass only after the plain command exited 0 above)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-merge-shape --e2e-steps 1 > gpurun_out/fin_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/fin_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(r[ui], 1e-6)
    k = r[ki].split("(")[0].split("<")[0].replace("void ", "").replace("db200::", "")
    tot[k] += v; cnt[k] += 1
s = sum(tot.values())
with open('gpurun_out/fin_launches_summary.txt', 'w') as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-merge-shape --e2e-steps 1\n")
    f.write("(run after the same command had exited 0 without ncu; per-launch times are serialised and cold-cache; device arm and the host arms inside the capture)\n")
    f.write("launches captured: %d, summed duration %.1f ms\n\n" % (sum(cnt.values()), s))
    for k, v in tot.most_common(24):
        f.write("%-44s n=%5d %10.3f ms %6.1f%%\n" % (k, cnt[k], v, 100 * v / s))
print(open('gpurun_out/fin_launches_summary.txt').read())
PY
rm -f gpurun_out/fin_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sw_pass_kernel --launch-skip 400 -c 40 -o gpurun_out/fin_fwd_kernels -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-merge-shape --e2e-steps 1 > gpurun_out/fin_ncu_fwd.log 2>&1; tail -2 gpurun_out/fin_ncu_fwd.log
ncu -i gpurun_out/fin_fwd_kernels.ncu-rep --page details > gpurun_out/fin_fwd_kernels_details.txt 2>/dev/null
ncu -i gpurun_out/fin_fwd_kernels.ncu-rep --page raw --csv > gpurun_out/fin_fwd_kernels_raw.csv 2>/dev/null
bash scripts/forward_traffic.sh 2>&1 | tail -3
