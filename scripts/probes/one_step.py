"""Two device-resident steps of the headline workload and nothing else (for ncu passes: scripts/forward_traffic.sh)."""
import sys, ctypes as C
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from dysgu_b200 import _lib
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
for _ in range(2):
    rc=L.db200_sw_batch_device(0,C.byref(prm_c),d_q.data_ptr(),d_qo.data_ptr(),qb,d_t.data_ptr(),d_to.data_ptr(),tb,n,int(q.lengths.max()),int(t.lengths.max()),None,d_res.data_ptr(),d_cig.data_ptr(),cap,d_co.data_ptr(),d_st.data_ptr(),None,C.c_void_p(st.cuda_stream))
    assert rc==0
    torch.cuda.synchronize()
print("cells", cells, "qb", qb, "tb", tb)
