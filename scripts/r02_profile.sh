#!/bin/bash
# round 2 profile artefacts: launch list of the default bench command, --set full of the largest forward bucket kernel and of
# the banded NW kernels after the load batching (each only after the plain command exited 0)
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-merge-shape > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || { echo "plain bench failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-merge-shape > gpurun_out/r02_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(r[ui], 1e-6)
    k = r[ki].split("(")[0].split("<")[0].replace("void ", "").replace("db200::", "")
    tot[k] += v; cnt[k] += 1
s = sum(tot.values())
with open('gpurun_out/r02_launches_summary.txt', 'w') as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-merge-shape\n")
    f.write("(run after the same command had exited 0 without ncu; per-launch times are serialised and cold-cache; device arm, the three\n host arms and the Python front door all inside the capture)\n")
    f.write("launches captured: %d, summed duration %.1f ms\n\n" % (sum(cnt.values()), s))
    for k, v in tot.most_common(24):
        f.write("%-44s n=%5d %10.3f ms %6.1f%%\n" % (k, cnt[k], v, 100 * v / s))
print(open('gpurun_out/r02_launches_summary.txt').read())
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sw_pass_kernel --launch-skip 200 -c 30 -o gpurun_out/r02_fwd_kernels -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-merge-shape > gpurun_out/r02_ncu_fwd.log 2>&1; tail -2 gpurun_out/r02_ncu_fwd.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ed_banded_nw_kernel -c 5 -o gpurun_out/r02_ed_banded_nw_after -f python bench.py --workload hifi_nw --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_nw2.log 2>&1; tail -2 gpurun_out/r02_ncu_nw2.log
./dysgu_b200/int_peak > gpurun_out/r02_int_peak_2.json
STEPS=4 WLS="pe_contigs long_contigs remap_chain hifi_nw ont_nw hifi_hw remap_path hifi_path" bash scripts/bench_workloads.sh 2>&1 | tail -12
