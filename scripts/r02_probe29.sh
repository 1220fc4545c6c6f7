#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sw_gpu.py tests/test_golden_gpu.py tests/test_input_forms_gpu.py tests/test_callers_gpu.py -m gpu -x -q 2>&1 | tail -3
summ() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']; s=r['stage_ms_per_step']
print(sys.argv[2], 'GCUPS', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), 'whole', round(r.get('frac_whole_step') or 0,3), 'e2e', round(d['e2e']['value']), 'single', round(d['e2e_single_call']['value']), {k: round(v,2) for k,v in s.items() if v > 0.005})
" "$1" "$2"; }
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-merge-shape --e2e-steps 3"
$B > gpurun_out/p29_head.json 2> gpurun_out/p29_head.err; summ gpurun_out/p29_head.json headline_ovh8
for o in 4 0; do
DB200_SW_ROW_OVERHEAD=$o $B > gpurun_out/p29_head_$o.json 2> gpurun_out/p29_head_$o.err; summ gpurun_out/p29_head_$o.json headline_ovh$o
done
python scripts/parity_sweep.py --pairs 40000 --max-len 600 --seed 15 2>&1 | grep "TOTAL" | tail -1
cp dysgu_b200/libdysgu_b200_align.so /tmp/lib_unr4.so
DB200_NVCC_FLAGS="-DSW_UNR=8" python -m dysgu_b200.build --force > gpurun_out/p29_build8.log 2>&1 && { $B > gpurun_out/p29_head_unr8.json 2> gpurun_out/p29_head_unr8.err; summ gpurun_out/p29_head_unr8.json headline_unr8; }
cp /tmp/lib_unr4.so dysgu_b200/libdysgu_b200_align.so
