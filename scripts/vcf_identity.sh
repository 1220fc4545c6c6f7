#!/usr/bin/env bash
# `dysgu call` twice on the same input - stock, and with the alignment path routed to the GPU engine with caller-side
# batching (python -m dysgu_b200.run_dysgu) - then compare the VCFs and the wall-clock seconds (BASELINE metric part 2,
# north_star: "the resulting VCF from dysgu call must be identical").
#
#   scripts/vcf_identity.sh ref.fa input.bam [extra dysgu call options, e.g. -p 8 --mode pe]
#
# Needs an importable dysgu (pysam, htslib: neither this build container nor the GPU box image has them - run it where
# dysgu is installed, with this repository on PYTHONPATH and libdysgu_b200_align.so built).  The only run-dependent
# header line is ##command="..." (io_funcs.pyx:360,453: sys.argv), which is dropped before diffing; the records carry
# SVMETHOD=DYSGUv<version> (io_funcs.pyx:291), identical because the same dysgu package writes both files.
# Also records every alignment of the stock run (DB200_CAPTURE) so that scripts/replay_capture.py can replay them.
set -euo pipefail
REF="$1"; BAM="$2"; shift 2
OUT="${DB200_VCF_OUT:-vcf_identity_out}"
mkdir -p "$OUT"
t0=$(date +%s.%N)
DB200_CAPTURE="$OUT/pairs.jsonl.gz" python -m dysgu_b200.run_dysgu --stock run --overwrite -x "$@" "$REF" "$OUT/wd_stock" "$BAM" > "$OUT/stock.vcf"
t1=$(date +%s.%N)
python -m dysgu_b200.run_dysgu run --overwrite -x "$@" "$REF" "$OUT/wd_gpu" "$BAM" > "$OUT/gpu.vcf"
t2=$(date +%s.%N)
grep -v '^##command=' "$OUT/stock.vcf" > "$OUT/stock.cmp"
grep -v '^##command=' "$OUT/gpu.vcf" > "$OUT/gpu.cmp"
python - "$OUT" "$t0" "$t1" "$t2" <<'PY'
import json, sys
out, t0, t1, t2 = sys.argv[1], *map(float, sys.argv[2:5])
a = open(out + "/stock.cmp").read().splitlines(); b = open(out + "/gpu.cmp").read().splitlines()
rec = sum(1 for l in a if not l.startswith("#"))
print(json.dumps({"records": rec, "identical": a == b, "stock_wall_s": round(t1 - t0, 2), "gpu_wall_s": round(t2 - t1, 2)}))
sys.exit(0 if a == b else 1)
PY
python scripts/replay_capture.py "$OUT/pairs.jsonl.gz"
