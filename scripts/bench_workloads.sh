#!/bin/bash
# every named shape through bench.py (one JSON line each) -> gpurun_out/bench_workloads.jsonl
out=gpurun_out/bench_workloads.jsonl
: > $out
for w in ${WLS:-remap_windows pe_contigs long_contigs remap_chain hifi_nw ont_nw hifi_hw remap_path hifi_path}; do
  timeout 600 python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 >> $out 2> gpurun_out/bench_$w.err || echo "{\"workload\": \"$w\", \"failed\": true}" >> $out
done
python - <<'PY'
import json
for ln in open('gpurun_out/bench_workloads.jsonl'):
    d=json.loads(ln)
    if d.get('failed'): print(d); continue
    r=d['roofline']; c=d['cpu_baseline'] or {}
    print("%-14s value %8.0f  e2e %8.0f  ms %7.2f  roof %.3f  kms %7.2f  cpu %7.1f (mism %s)" % (d['config']['workload'].split(':')[0], d['value'], d['e2e']['value'], d['ms_per_step'], r['frac'] or 0, r['kernel_ms_per_step'], c.get('value') or 0, c.get('parity_mismatches_on_sample')))
PY
