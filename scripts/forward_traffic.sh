#!/bin/bash
# DRAM traffic of the forward scoring launches of ONE step of the default bench command, measured with ncu on this code
# (run only after the plain command exited 0) -> profiles/r02_forward_traffic.json, read by bench.py for roofline.traffic.
set -u
mkdir -p gpurun_out
python scripts/probes/one_step.py > /dev/null 2>&1 || { echo "the step failed without ncu"; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sw_pass_kernel --csv \
    --log-file gpurun_out/r02_fwd_traffic.csv python scripts/probes/one_step.py > gpurun_out/r02_fwd_traffic.log 2>&1
python - <<'PY'
import csv, json
rows = [r for r in csv.reader(open('gpurun_out/r02_fwd_traffic.csv')) if len(r) > 10]
hdr = rows[0]; ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
ii = hdr.index("ID")
per = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    per.setdefault(int(r[ii]), {"kernel": r[ki]})[r[mi]] = v
ids = sorted(per)
# scripts/probes/one_step.py runs two identical device-resident steps of the bench's workload (same generator, same seed);
# every step launches the same list of sw_pass_kernel instances.  Take the second step.
n_per_step = len(ids) // 2
step = ids[n_per_step:]
tot = sum(per[i].get("dram__bytes_read.sum", 0) + per[i].get("dram__bytes_write.sum", 0) for i in step)
json.dump({"pairs_per_step": 524288, "dram_bytes_per_step": tot, "launches_per_step": n_per_step,
           "note": "all sw_pass_kernel launches of one step (forward buckets, reverse buckets and re-runs); ncu --clock-control none"},
          open('gpurun_out/r02_forward_traffic.json', 'w'))
print(open('gpurun_out/r02_forward_traffic.json').read())
PY
