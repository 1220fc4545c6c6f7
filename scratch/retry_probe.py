import sys, collections
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from dysgu_b200 import batch, workloads
from oracle import oracle as O
a, b, prm = workloads.contig_pairs(300, seed=17, platform="ont", lo=1500, hi=4000)
qs, ts = a.tolist(), b.tolist()
res = batch.sw_align_batch(qs, ts, **prm)
print("status hist", collections.Counter(res.status.tolist()))
ref = O.oracle_sw(qs, ts, prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"])
bad = [i for i in range(len(qs)) if res.status[i] != 0 or res.as_tuple(i) != ref.as_tuple(i)]
print("BAD", len(bad))
for i in bad[:3]: print(res.status[i], res.as_tuple(i)[:7], ref.as_tuple(i)[:7], res.as_tuple(i)[7][:60], ref.as_tuple(i)[7][:60])
