"""usage: variant_bench.py <lib.so> : parity sample vs oracle + forward-stage time of the headline step"""
import sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from dysgu_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
import numpy as np, torch
import pairgen
from oracle import oracle as O
from dysgu_b200 import batch, workloads
bad = 0
for mx, sc in ((250, (2,-8,6,1)), (1000, (2,-3,10,1)), (400, (1,-4,6,1))):
    qs, ts = pairgen.sw_pairs(mx, 1500, mx)
    r = batch.sw_align_batch(qs, ts, *sc)
    ref = O.oracle_sw(qs, ts, *sc)
    bad += sum(1 for i in range(len(qs)) if r.status[i] != 0 or r.as_tuple(i) != ref.as_tuple(i))
q, t, p = workloads.remap_windows(20000, seed=3)
r = batch.sw_align_batch(q, t, **p); ref = O.oracle_sw(q.tolist(), t.tolist(), 2, -8, 6, 1)
bad += sum(1 for i in range(len(q)) if r.status[i] != 0 or r.as_tuple(i) != ref.as_tuple(i))
import bench
L=_lib.lib(); L.db200_init(0)
q,t,prm,cells = bench.make_workload(524288, 20261021)
n=len(q); qb=int(q.off[-1]); tb=int(t.off[-1])
prm_c=_lib.SwParams(); prm_c.match,prm_c.mismatch,prm_c.gap_open,prm_c.gap_extend=2,-8,6,1; prm_c.score_size,prm_c.flag=2,1
dev=torch.device('cuda',0)
d_q=torch.from_numpy(q.buf).to(dev); d_t=torch.from_numpy(t.buf).to(dev); d_qo=torch.from_numpy(q.off).to(dev); d_to=torch.from_numpy(t.off).to(dev)
cap=4*n+1024
d_res=torch.zeros((n,8),dtype=torch.int32,device=dev); d_cig=torch.zeros(cap,dtype=torch.int32,device=dev); d_co=torch.zeros(n+1,dtype=torch.int64,device=dev); d_st=torch.zeros(n,dtype=torch.int32,device=dev)
st=torch.cuda.current_stream(dev)
def step():
    rc=L.db200_sw_batch_device(0,C.byref(prm_c),d_q.data_ptr(),d_qo.data_ptr(),qb,d_t.data_ptr(),d_to.data_ptr(),tb,n,int(q.lengths.max()),int(t.lengths.max()),None,d_res.data_ptr(),d_cig.data_ptr(),cap,d_co.data_ptr(),d_st.data_ptr(),None,C.c_void_p(st.cuda_stream))
    assert rc==0
    torch.cuda.synchronize()
for _ in range(3): step()
L.db200_profile_enable(0,1)
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4): step()
e1.record(); torch.cuda.synchronize()
sm=(C.c_float*16)(); L.db200_profile_collect(0,sm,16)
print("%-28s bad=%d  step %.2f ms  forward %.2f  reverse %.2f" % (sys.argv[1].split('/')[-1], bad, e0.elapsed_time(e1)/4, sm[3]/4, sm[6]/4))
