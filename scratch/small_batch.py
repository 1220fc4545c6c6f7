import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from dysgu_b200 import workloads, batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
q, t, p = workloads.remap_windows(n, seed=1)
for _ in range(2):
    r = batch.sw_align_batch(q, t, **p)
print("ok", int(r.score1.sum()))
