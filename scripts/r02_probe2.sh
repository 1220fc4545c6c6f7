#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_compat_gpu.py tests/test_concurrency_gpu.py tests/test_long_queries_gpu.py tests/test_ed_path_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_new_1.txt
./dysgu_b200/int_peak > gpurun_out/r02_int_peak_1.json 2>&1
cat gpurun_out/r02_pytest_new_1.txt gpurun_out/r02_int_peak_1.json
