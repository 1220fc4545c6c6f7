import sys, time, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np
from dysgu_b200 import _lib, workloads, batch
L=_lib.lib(); L.db200_init(0)
names=["begin","pack","bucket","forward","fwd_rerun","rev_prep","reverse","rev_rerun","finalize","traceback","cigar","ed_prep","ed_fwd","ed_starts","ed_out"]
def probe(tag, fn):
    fn(); 
    L.db200_profile_enable(0,1)
    t0=time.perf_counter(); fn(); dt=time.perf_counter()-t0
    sm=(C.c_float*16)(); L.db200_profile_collect(0,sm,16); L.db200_profile_enable(0,0)
    print(tag, "wall %.1f ms"%(dt*1e3), {n:round(x,2) for n,x in zip(names,sm) if x>0.005})
which = sys.argv[1] if len(sys.argv)>1 else "all"
if which in ("all","contigs"):
    q,t,p = workloads.contig_pairs(4000, seed=5, platform="ont")
    probe("contigs ont 4000", lambda: batch.sw_align_batch(q,t,**p))
    probe("contigs ont 4000 flag8", lambda: batch.sw_align_batch(q,t,flag=8,**p))
    q,t,p = workloads.contig_pairs(4000, seed=5, platform="hifi")
    probe("contigs hifi 4000", lambda: batch.sw_align_batch(q,t,**p))
if which in ("all","pe"):
    q,t,p = workloads.pe_contig_pairs(200000, seed=2)
    probe("pe contigs 200k", lambda: batch.sw_align_batch(q,t,**p))
if which in ("all","ed"):
    q,t,p = workloads.remap_chain(400000, seed=1)
    probe("ed remap chain 400k", lambda: batch.ed_align_batch(q,t,**p))
    q,t,p = workloads.insertion_linking(20000, seed=3, platform="hifi")
    probe("ed hifi NW 20k", lambda: batch.ed_align_batch(q,t,**p))
    probe("ed hifi HW loc 20k", lambda: batch.ed_align_batch(q,t,mode="HW",task="locations"))
if which in ("remap",):
    import os
    for n in (8192, 32768, 65536, 131072, 524288):
        q,t,p = workloads.remap_windows(n, seed=1)
        os.environ["DB200_HOST_CHUNK_PAIRS"] = str(1<<20)
        probe("remap %d" % n, lambda: batch.sw_align_batch(q,t,**p))
