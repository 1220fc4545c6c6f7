#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sw_pass_kernel<\(int\)13, \(int\)4, \(bool\)0, \(bool\)0, \(bool\)0, \(bool\)1" --launch-skip 1 -c 1 -o gpurun_out/src16 -f python scripts/probes/one_step.py > gpurun_out/src16.log 2>&1
tail -2 gpurun_out/src16.log
ncu -i gpurun_out/src16.ncu-rep --page source --csv --print-source sass > gpurun_out/src16_sass.csv 2>/dev/null
ncu -i gpurun_out/src16.ncu-rep --page details 2>/dev/null | grep -E "Duration|highest-utilized|No Eligible|Warp Cycles Per Issued|Issue Slots Busy|Achieved Occupancy|Registers Per" 
rm -f gpurun_out/src16.ncu-rep
wc -l gpurun_out/src16_sass.csv; du -sh gpurun_out
