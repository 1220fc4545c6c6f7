#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(' '.join('%s %s %.2f' % (k, round(d[k]['value']), d[k]['ms_per_step']) for k in ('e2e', 'e2e_single_call', 'e2e_ascii')), 'value', round(d['value']), round(d['ms_per_step'], 2), {k: round(v, 2) for k, v in d['roofline']['stage_ms_per_step'].items()})
PY
}
for mc in 8 16 32; do
for cfg in "2 524288" "2 262144" "2 131072"; do
  set -- $cfg
  CUDA_DEVICE_MAX_CONNECTIONS=$mc DB200_BENCH_STREAM_THREADS=$1 DB200_BENCH_STREAM_CHUNK=$2 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_stream.json 2> gpurun_out/r02_bench_stream.err
  echo "maxconn $mc threads $1 chunk $2"; tail -2 gpurun_out/r02_bench_stream.err; show gpurun_out/r02_bench_stream.json
done
done
