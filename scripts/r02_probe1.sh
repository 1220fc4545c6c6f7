#!/bin/bash
# round 2, first GPU call: full GPU suite on the fixed host pipeline, the never-run long-query test, latency probe,
# integer peak, ncu --set full of the banded NW kernels (hifi_nw / ont_nw)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_0.txt
timeout 900 python -m pytest tests/pending_long_queries_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_long_queries_0.txt
timeout 300 python scratch/latency_probe.py > gpurun_out/r02_latency_0.txt 2>&1
./dysgu_b200/int_peak > gpurun_out/r02_int_peak_0.json 2>&1
timeout 300 python bench.py --workload hifi_nw --steps 2 --warmup 1 > gpurun_out/r02_hifi_nw_0.json 2> gpurun_out/r02_hifi_nw_0.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ed_banded_nw_kernel -c 8 -o gpurun_out/r02_ed_banded_nw_0 -f python bench.py --workload hifi_nw --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_nw.log 2>&1
cat gpurun_out/r02_pytest_gpu_0.txt gpurun_out/r02_long_queries_0.txt gpurun_out/r02_latency_0.txt
