#!/bin/bash
# Fine-grained (K, G) configurations of the forward kernel: GPU suite, then the headline bench with and without them
# and with different row-overhead terms of the cost model.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_sw_gpu.py tests/test_golden_gpu.py tests/test_input_forms_gpu.py tests/test_small_path_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/p15_pytest.txt
cat gpurun_out/p15_pytest.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-merge-shape --e2e-steps 2"
summ() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']; s=r['stage_ms_per_step']
print(sys.argv[2], 'GCUPS', round(d['value']), 'ms', round(d['ms_per_step'],2), 'fwd', round(s['forward'],2), 'rev', round(s['reverse'],2), 'trace', round(s['traceback'],2), 'bucket', round(s['bucket'],2), 'frac', round(r['frac'],3), 'whole', round(r['frac_whole_step'],3), 'e2e', round(d['e2e']['value']))
" "$1" "$2"; }
DB200_SW_NO_FINE=1 $B > gpurun_out/p15_classic.json 2> gpurun_out/p15_classic.err; summ gpurun_out/p15_classic.json classic
$B > gpurun_out/p15_fine21.json 2> gpurun_out/p15_fine21.err; summ gpurun_out/p15_fine21.json fine_ovh21
for o in 8 14 30 40; do
DB200_SW_ROW_OVERHEAD=$o $B > gpurun_out/p15_fine$o.json 2> gpurun_out/p15_fine$o.err; summ gpurun_out/p15_fine$o.json fine_ovh$o
done
