import sys
sys.path.insert(0,'/root/repo')
from dysgu_b200 import workloads, batch
q,t,p = workloads.contig_pairs(4000, seed=5, platform="ont")
r = batch.sw_align_batch(q,t,**p)
r = batch.sw_align_batch(q,t,**p)
print("ok", int(r.cigar_len.sum()))
