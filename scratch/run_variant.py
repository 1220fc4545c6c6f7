import sys, os, shutil
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from dysgu_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
import pairgen
from oracle import oracle as O
from dysgu_b200 import batch
for mx in (250, 1000):
    qs, ts = pairgen.sw_pairs(mx, 600, mx)
    try:
        r = batch.sw_align_batch(qs, ts, 2, -8, 6, 1)
        ref = O.oracle_sw(qs, ts, 2, -8, 6, 1)
        bad = sum(1 for i in range(len(qs)) if r.status[i] != 0 or r.as_tuple(i) != ref.as_tuple(i))
        print(sys.argv[1], mx, "bad", bad)
    except Exception as e:
        print(sys.argv[1], mx, "EXC", str(e)[:150]); break
