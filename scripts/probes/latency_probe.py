import sys, time
sys.path.insert(0,'/root/repo')
from dysgu_b200.scikitbio import StripedSmithWaterman
from dysgu_b200.edlib import align
from dysgu_b200 import batch, workloads
import numpy as np
q,t,p = workloads.remap_windows(64, seed=1)
qs=[x.decode() for x in q.tolist()]; ts=[x.decode() for x in t.tolist()]
s = StripedSmithWaterman(qs[0], gap_open_penalty=6, gap_extend_penalty=1, match_score=2, mismatch_score=-8)
s(ts[0])
t0=time.perf_counter()
for i in range(200): s(ts[i%64])
print("SSW drop-in call (window 1-2kb vs clip): %.1f us/call" % ((time.perf_counter()-t0)/200*1e6))
t0=time.perf_counter()
for i in range(200): StripedSmithWaterman(qs[i%64], gap_open_penalty=6, gap_extend_penalty=1, match_score=2, mismatch_score=-8)(ts[i%64])
print("SSW init+call: %.1f us/call" % ((time.perf_counter()-t0)/200*1e6))
align(ts[0].upper(), qs[0], "HW", "locations")
t0=time.perf_counter()
for i in range(200): align(ts[i%64].upper(), qs[i%64], "HW", "locations")
print("edlib drop-in call: %.1f us/call" % ((time.perf_counter()-t0)/200*1e6))
for n in (1, 8, 64, 512, 4096):
    q,t,p = workloads.remap_windows(n, seed=2)
    batch.sw_align_batch(q,t,**p)
    t0=time.perf_counter()
    for _ in range(20): batch.sw_align_batch(q,t,**p)
    dt=(time.perf_counter()-t0)/20
    print("sw_align_batch n=%d: %.1f us/batch, %.2f us/pair" % (n, dt*1e6, dt*1e6/n))
