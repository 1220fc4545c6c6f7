#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(' '.join('%s %s %.2f' % (k, round(d[k]['value']), d[k]['ms_per_step']) for k in ('e2e', 'e2e_single_call', 'e2e_ascii')), 'value', round(d['value']), round(d['ms_per_step'], 2))
PY
}
for cfg in "2 524288" "2 524288" "3 524288"; do
  set -- $cfg
  DB200_BENCH_STREAM_THREADS=$1 DB200_BENCH_STREAM_CHUNK=$2 timeout 600 python bench.py --steps 6 --warmup 3 --e2e-steps 6 --no-cpu-baseline > gpurun_out/r02_bench_stream.json 2> gpurun_out/r02_bench_stream.err
  echo "threads $1 chunk $2"; tail -2 gpurun_out/r02_bench_stream.err; show gpurun_out/r02_bench_stream.json
done
DB200_TIMELINE=1 timeout 300 python bench.py --steps 1 --warmup 3 --e2e-steps 2 --no-cpu-baseline 2>&1 >/dev/null | grep -A2 "timeline" | tail -12
