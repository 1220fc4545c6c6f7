#!/bin/bash
set -u
mkdir -p gpurun_out
{ echo "# python scripts/parity_sweep.py --pairs 60000 --max-len 600 --seed 21 --forms --tasks nopath"; timeout 1500 python scripts/parity_sweep.py --pairs 60000 --max-len 600 --seed 21 --forms --tasks nopath;
  echo; echo "# python scripts/parity_sweep.py --pairs 8000 --max-len 4000 --seed 22 --forms --tasks nopath"; timeout 1500 python scripts/parity_sweep.py --pairs 8000 --max-len 4000 --seed 22 --forms --tasks nopath;
  echo; echo "# python scripts/parity_sweep.py --pairs 24 --max-len 65536 --seed 23 --tasks ed   (edit distance incl. task path, queries up to 65,536 symbols)"; timeout 2400 python scripts/parity_sweep.py --pairs 24 --max-len 65536 --seed 23 --tasks ed; } > gpurun_out/r02_parity_sweep.txt 2>&1
grep -E "TOTAL|Traceback|Error|diffs=[1-9]" gpurun_out/r02_parity_sweep.txt | head -20
