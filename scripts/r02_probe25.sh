#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_sw_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3
summ() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']; s=r['stage_ms_per_step']
print(sys.argv[2], 'GCUPS', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), 'whole', round(r.get('frac_whole_step') or 0,3), 'e2e', round(d['e2e']['value']), {k: round(v,2) for k,v in s.items() if v > 0.005})
" "$1" "$2"; }
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-merge-shape --e2e-steps 2"
for mj in 16384 4096 512 0; do
DB200_TRACE_MEDIUM_MIN_JOBS=$mj $B > gpurun_out/p25_head_$mj.json 2> gpurun_out/p25_head_$mj.err; summ gpurun_out/p25_head_$mj.json headline_minjobs_$mj
done
python bench.py --workload pe_contigs --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/p25_pe.json 2> gpurun_out/p25_pe.err; summ gpurun_out/p25_pe.json pe_contigs
