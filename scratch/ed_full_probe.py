import sys
sys.path.insert(0,'/root/repo')
import numpy as np
from dysgu_b200 import _lib, workloads, batch
q,t,p = workloads.insertion_linking(6000, seed=3, platform="hifi", lo=2000, hi=10000)
for _ in range(2):
    r = batch.ed_align_batch(q,t,mode="HW",task="locations")
print("ok", int(r.fixed[:,1].sum()))
