import sys
sys.path.insert(0,'/root/repo')
from dysgu_b200 import workloads, batch
q,t,p = workloads.contig_pairs(600, seed=5, platform="ont", lo=6000, hi=9000)
r = batch.sw_align_batch(q,t,**p)
r = batch.sw_align_batch(q,t,**p)
print("ok", int(r.cigar_len.sum()))
