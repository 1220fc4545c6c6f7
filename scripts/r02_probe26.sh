#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sw_gpu.py tests/test_golden_gpu.py tests/test_callers_gpu.py -m gpu -x -q 2>&1 | tail -3
summ() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']; s=r['stage_ms_per_step']
print(sys.argv[2], 'GCUPS', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), 'whole', round(r.get('frac_whole_step') or 0,3), 'e2e', round(d['e2e']['value']), {k: round(v,2) for k,v in s.items() if v > 0.005})
" "$1" "$2"; }
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-merge-shape --e2e-steps 2"
$B > gpurun_out/p26_head.json 2> gpurun_out/p26_head.err; summ gpurun_out/p26_head.json headline

for w in pe_contigs long_contigs; do
python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/p26_$w.json 2> gpurun_out/p26_$w.err; summ gpurun_out/p26_$w.json $w
done

python scripts/parity_sweep.py --pairs 30000 --max-len 3000 --seed 14 2>&1 | grep "TOTAL" | tail -2
