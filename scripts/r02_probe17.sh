#!/bin/bash
# merge shapes after the fine configurations: stage split of pe_contigs / long_contigs, with and without them
set -u
mkdir -p gpurun_out
summ() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']; s=r['stage_ms_per_step']
print(sys.argv[2], 'GCUPS', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), {k: round(v,2) for k,v in s.items() if v > 0.005})
" "$1" "$2"; }
for w in pe_contigs long_contigs; do
  python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/p17_$w.json 2> gpurun_out/p17_$w.err; summ gpurun_out/p17_$w.json $w
  DB200_SW_NO_FINE=1 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/p17_${w}_classic.json 2> gpurun_out/p17_${w}_classic.err; summ gpurun_out/p17_${w}_classic.json ${w}_classic
done
