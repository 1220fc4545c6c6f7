#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_small_path_gpu.py tests/test_patch_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_new_2.txt
cat gpurun_out/r02_pytest_new_2.txt
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_2.txt
cat gpurun_out/r02_pytest_gpu_2.txt
timeout 600 python scripts/latency_table.py > gpurun_out/r02_latency_1.json 2> gpurun_out/r02_latency_1.err
cat gpurun_out/r02_latency_1.json | head -60; tail -5 gpurun_out/r02_latency_1.err
