#!/bin/bash
# --set full captures of this round's code, turned into text on the box (the reports themselves exceed what travels back)
set -u
mkdir -p gpurun_out
KEEP='Duration|Registers Per Thread|Achieved Occupancy|Theoretical Occupancy|Executed Ipc Active|ALU|Issue Slots Busy|L1/TEX Hit Rate|L2 Hit Rate|Avg. Active Threads Per Warp|No Eligible|Warp Cycles Per Issued Instruction|DRAM Throughput|Memory Throughput|Compute \(SM\) Throughput|Dynamic Shared Memory|Block Limit|Eligible Warps|Active Warps Per Scheduler|dram__bytes|Stall'
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-merge-shape > /dev/null 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sw_pass_kernel --launch-skip 72 -c 24 -o /tmp/r02_fwd -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-merge-shape > gpurun_out/r02_ncu_fwd.log 2>&1; tail -1 gpurun_out/r02_ncu_fwd.log
{ echo "# ncu --set full --clock-control none --import-source on -k regex:sw_pass_kernel --launch-skip 72 -c 24 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-merge-shape"
  echo "# (the 24 forward bucket launches of the fourth device-resident step; round-2 code; extract of the details page per launch, longest first)"
  ncu -i /tmp/r02_fwd.ncu-rep --page details 2>/dev/null | grep -E "sw_pass_kernel|$KEEP" ; } > gpurun_out/r02_fwd_kernel_details_all.txt
ncu -i /tmp/r02_fwd.ncu-rep --page raw --csv 2>/dev/null > /tmp/r02_fwd_raw.csv
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/r02_fwd_raw.csv')))
hdr = rows[0]
g = lambda r, k: r[hdr.index(k)] if k in hdr else ''
best = None
out = []
for r in rows[2:]:
    d = float(g(r, 'gpu__time_duration.sum') or 0)
    out.append((d, g(r, 'Kernel Name'), g(r, 'ID')))
out.sort(reverse=True)
with open('gpurun_out/r02_fwd_kernel_launch_times.txt', 'w') as f:
    for d, k, i in out:
        f.write("%10.3f ms  id %s  %s\n" % (d, i, k))
print(open('gpurun_out/r02_fwd_kernel_launch_times.txt').read()[:1500])
PY
timeout 900 ncu --set full --clock-control none -k regex:ed_banded_nw_kernel -c 5 -o /tmp/r02_nw -f python bench.py --workload hifi_nw --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_nw2.log 2>&1; tail -1 gpurun_out/r02_ncu_nw2.log
{ echo "# ncu --set full --clock-control none, ed_banded_nw_kernel on bench.py --workload hifi_nw after the load batching (2-8 columns fetched before they are stepped)"
  ncu -i /tmp/r02_nw.ncu-rep --page details 2>/dev/null | grep -E "ed_banded_nw_kernel|$KEEP" ; } > gpurun_out/r02_ed_banded_nw_after.txt
head -40 gpurun_out/r02_ed_banded_nw_after.txt
rm -f gpurun_out/*.ncu-rep
timeout 600 python scripts/latency_table.py > gpurun_out/r02_latency_3.json 2> gpurun_out/r02_latency_3.err; python -c "
import json; d=json.load(open('gpurun_out/r02_latency_3.json')); print(d['shapes']['micro_150x100']); print(d['shapes']['remap_window_vs_clip'])"
timeout 600 python -m pytest tests/test_small_path_gpu.py tests/test_golden_gpu.py -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
du -sh gpurun_out
