#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/probes/e2e_timeline.py > gpurun_out/r02_timeline.txt 2>&1; cat gpurun_out/r02_timeline.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ed_myers_kernel<5, 32|ed_myers_kernel<4, 32|ed_banded_nw_kernel<2|ed_banded_nw_kernel<17" -c 8 -o gpurun_out/r02_ed_kernels_1 -f python bench.py --workload hifi_nw --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_ed1.log 2>&1
tail -3 gpurun_out/r02_ncu_ed1.log
