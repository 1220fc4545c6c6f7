#!/bin/bash
# two GPUs: the multi-device tests with real devices, then the bench under torchrun at N = 2
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_golden_gpu.py -q -k "multi" 2>&1 | tail -5 > gpurun_out/r02_pytest_2gpu.txt; cat gpurun_out/r02_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -3 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_2gpu.json'))
for k in ('e2e', 'e2e_single_call', 'e2e_ascii'):
    print(k, d[k] and round(d[k]['value']), d[k] and d[k].get('ms_per_step'))
print('value', round(d['value']), d['ms_per_step'], d['n_gpus'], d['config']['host'])
print('merge', d['merge_shape'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_2gpu.json 2> gpurun_out/r02_bench_ref_2gpu.err; head -c 300 gpurun_out/r02_bench_ref_2gpu.json; echo
