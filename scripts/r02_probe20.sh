#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_sw_gpu.py tests/test_golden_gpu.py tests/test_callers_gpu.py tests/test_small_path_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/p20_pytest.txt
cat gpurun_out/p20_pytest.txt
summ() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']; s=r['stage_ms_per_step']
print(sys.argv[2], 'GCUPS', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), {k: round(v,2) for k,v in s.items() if v > 0.005})
" "$1" "$2"; }
for w in pe_contigs long_contigs remap_windows; do
  python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-merge-shape --e2e-steps 2 > gpurun_out/p20_$w.json 2> gpurun_out/p20_$w.err; summ gpurun_out/p20_$w.json $w
done
python scripts/parity_sweep.py --pairs 40000 --max-len 3000 --seed 11 2>&1 | tail -5
