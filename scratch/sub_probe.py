import sys, os, time, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from dysgu_b200 import _lib
import bench
L=_lib.lib(); L.db200_init(0)
pairs=int(sys.argv[1]) if len(sys.argv)>1 else 524288
q,t,prm,cells = bench.make_workload(pairs, 20261021)
n=len(q); qb=int(q.off[-1]); tb=int(t.off[-1])
prm_c=_lib.SwParams(); prm_c.match,prm_c.mismatch,prm_c.gap_open,prm_c.gap_extend=2,-8,6,1; prm_c.score_size,prm_c.flag=2,1
dev=torch.device('cuda',0)
d_q=torch.from_numpy(q.buf).to(dev); d_t=torch.from_numpy(t.buf).to(dev); d_qo=torch.from_numpy(q.off).to(dev); d_to=torch.from_numpy(t.off).to(dev)
cap=4*n+1024
def bufs():
    return (torch.zeros((n,8),dtype=torch.int32,device=dev), torch.zeros(cap,dtype=torch.int32,device=dev), torch.zeros(n+1,dtype=torch.int64,device=dev), torch.zeros(n,dtype=torch.int32,device=dev))
st=torch.cuda.current_stream(dev)
mq,mt=int(q.lengths.max()),int(t.lengths.max())
def step(b):
    rc=L.db200_sw_batch_device(0,C.byref(prm_c),d_q.data_ptr(),d_qo.data_ptr(),qb,d_t.data_ptr(),d_to.data_ptr(),tb,n,mq,mt,None,b[0].data_ptr(),b[1].data_ptr(),cap,b[2].data_ptr(),b[3].data_ptr(),None,C.c_void_p(st.cuda_stream))
    assert rc==0, _lib.last_error() if hasattr(_lib,'last_error') else rc
os.environ["DB200_DEVICE_SUBBATCHES"]="1"
ref=bufs(); step(ref); torch.cuda.synchronize()
for ns in (1,2,3,4,6,8,12):
    os.environ["DB200_DEVICE_SUBBATCHES"]=str(ns)
    b=bufs()
    for _ in range(3): step(b)
    torch.cuda.synchronize()
    ok = all(torch.equal(x,y) for x,y in zip(ref,b))
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5): step(b)
    e1.record(st); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print("nsub",ns,"ms/step %.2f"%ms,"GCUPS %.0f"%(cells/ms/1e6),"identical" if ok else "DIFFERENT", flush=True)
del os.environ["DB200_DEVICE_SUBBATCHES"]
# stage profile with nsub=1
os.environ["DB200_DEVICE_SUBBATCHES"]="1"
L.db200_profile_enable(0,1)
for _ in range(3): step(ref)
sm=(C.c_float*16)(); L.db200_profile_collect(0,sm,16); L.db200_profile_enable(0,0)
print("stages nsub=1", [round(x/3,2) for x in sm[:11]])
