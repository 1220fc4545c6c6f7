#!/bin/bash
# eight GPUs: the bench under torchrun at N = 8 and N = 4 (weak scaling, pairs sharded by rank)
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; numactl -H 2>/dev/null | head -4
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err; tail -2 gpurun_out/r02_bench_${n}gpu.err
python - $n <<'PY'
import json, sys
d = json.load(open('gpurun_out/r02_bench_%sgpu.json' % sys.argv[1]))
print('N', d['n_gpus'], ' '.join('%s %s %.2f' % (k, round(d[k]['value']), d[k]['ms_per_step']) for k in ('e2e', 'e2e_single_call', 'e2e_ascii')), 'value', round(d['value']), round(d['ms_per_step'], 2), d['config']['host'])
print('merge', round(d['merge_shape']['value']), round(d['merge_shape']['e2e']['value']))
PY
done
