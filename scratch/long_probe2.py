import sys, os, time, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np
from dysgu_b200 import _lib, workloads, batch
L=_lib.lib(); L.db200_init(0)
names=["begin","pack","bucket","forward","fwd_rerun","rev_prep","reverse","rev_rerun","finalize","traceback","cigar"]
def probe(tag, fn, cells):
    fn()
    L.db200_profile_enable(0,1)
    t0=time.perf_counter(); r=fn(); dt=time.perf_counter()-t0
    sm=(C.c_float*16)(); L.db200_profile_collect(0,sm,16); L.db200_profile_enable(0,0)
    print(tag, "wall %.1f ms  %.0f GCUPS"%(dt*1e3, cells/dt/1e9), {n:round(x,2) for n,x in zip(names,sm) if x>0.005}, flush=True)
    return r
for n in (1000, 4000, 16000):
    q,t,p = workloads.contig_pairs(n, seed=5, platform="ont")
    cells=float((q.lengths.astype(np.float64)*t.lengths).sum())
    q,t=batch.pinned_copy(q),batch.pinned_copy(t)
    a=probe("contigs ont %d flag8 piped"%n, lambda: batch.sw_align_batch(q,t,flag=8,**p), cells)
    os.environ["DB200_SW_NO_PIPE"]="1"
    b=probe("contigs ont %d flag8 seq  "%n, lambda: batch.sw_align_batch(q,t,flag=8,**p), cells)
    del os.environ["DB200_SW_NO_PIPE"]
    print("  identical:", bool(np.array_equal(a.fixed,b.fixed)))
