#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sw_trace_group_kernel|sw_trace_coop_kernel" -c 7 -o gpurun_out/p27_trace -f python bench.py --workload long_contigs --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/p27_ncu.log 2>&1
tail -3 gpurun_out/p27_ncu.log
ncu -i gpurun_out/p27_trace.ncu-rep --page raw --csv > gpurun_out/p27_trace_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/p27_trace_raw.csv')))
hdr=rows[0]
want=["Kernel Name","gpu__time_duration.sum","sm__inst_executed.avg.per_cycle_elapsed","smsp__inst_executed.sum","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","smsp__average_warp_latency_issue_stalled_long_scoreboard_pipe_l1tex.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_membar_per_issue_active.ratio","smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__warps_eligible.avg.per_cycle_active","dram__bytes_write.sum","dram__bytes_read.sum","lts__t_sectors_op_write.sum","launch__grid_size","launch__occupancy_limit_shared_mem","smsp__warps_active.avg.per_cycle_active"]
idx=[hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print("----")
    for i in idx: print("  ",hdr[i][:70], r[i])
PY
