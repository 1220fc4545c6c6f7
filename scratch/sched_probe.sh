#!/bin/bash
# e2e of the headline step under different host chunk schedules
run() { echo -n "$1: "; DB200_HOST_CHUNK_SCHEDULE="$1" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))"; }
echo -n "default: "; python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))"
run 32768,65536,131072,131072,98304,65536
run 32768,65536,98304,131072,98304,65536,32768
run 32768,65536,65536
run 32768,65536,98304,98304
run 16384,32768,65536,131072,131072,98304,49152
run 65536
run 49152,98304,131072
