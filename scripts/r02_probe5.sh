#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_input_forms_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_new_4.txt
cat gpurun_out/r02_pytest_new_4.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_3.txt
cat gpurun_out/r02_pytest_gpu_3.txt
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d['roofline']
print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 2),
      'ascii', round(d['e2e_ascii']['value']), 'front', d.get('python_front_door') and round(d['python_front_door']['value']), 'frac', round(r['frac'], 3),
      'whole', r.get('frac_whole_step') and round(r['frac_whole_step'], 3), 'cpu', d['cpu_baseline'])
PY
}
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_2.json 2> gpurun_out/r02_bench_2.err; tail -3 gpurun_out/r02_bench_2.err; show gpurun_out/r02_bench_2.json
for sch in "49152,131072,344064" "32768,65536,131072,294912" "65536,196608,262144"; do
  DB200_HOST_CHUNK_SCHEDULE=$sch timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_sch.json 2> gpurun_out/r02_bench_sch.err; echo "schedule $sch"; show gpurun_out/r02_bench_sch.json
done
timeout 600 bash scripts/forward_traffic.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_hifi_nw.csv python bench.py --workload hifi_nw --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_hifi_nw_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_launches_hifi_nw.csv')) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(r[ui], 1e-6)
    k = r[ki].split("(")[0]
    tot[k] += v; cnt[k] += 1
for k, v in tot.most_common(14): print("%9.3f ms %4d  %s" % (v, cnt[k], k))
PY
