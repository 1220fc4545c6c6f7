import sys, os, time, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np
from dysgu_b200 import _lib, batch
import bench
L=_lib.lib(); L.db200_init(0)
q,t,prm,cells = bench.make_workload(524288, 20261021)
q,t = batch.pinned_copy(q), batch.pinned_copy(t)
n=len(q)
prm_c=_lib.SwParams(); prm_c.match,prm_c.mismatch,prm_c.gap_open,prm_c.gap_extend=2,-8,6,1; prm_c.score_size,prm_c.flag=2,1
cap=4*n+1024
hres=bench.pinned_array(L,(n,8),np.int32); hcig=bench.pinned_array(L,(cap,),np.uint32); hco=bench.pinned_array(L,(n+1,),np.int64); hst=bench.pinned_array(L,(n,),np.int32)
def step():
    rc=L.db200_sw_batch_host(0,C.byref(prm_c),q.buf.ctypes.data,q.off.ctypes.data,t.buf.ctypes.data,t.off.ctypes.data,n,None,hres.ctypes.data,hcig.ctypes.data,cap,hco.ctypes.data,hst.ctypes.data)
    assert rc==0
def timeit(tag, reps=5):
    for _ in range(2): step()
    t0=time.perf_counter()
    for _ in range(reps): step()
    dt=(time.perf_counter()-t0)/reps
    print(tag, "%.2f ms  %.0f GCUPS"%(dt*1e3, cells/dt/1e9), flush=True)
timeit("default")
os.environ["DB200_TIMELINE"]="1"; step(); del os.environ["DB200_TIMELINE"]
for sched in ("65536,131072,163840,163840", "131072", "174763", "262144", "87382", "32768,65536,131072,147456,147456", "49152,98304,188416,188416"):
    os.environ["DB200_HOST_CHUNK_SCHEDULE"]=sched
    timeit("schedule "+sched)
del os.environ["DB200_HOST_CHUNK_SCHEDULE"]
