"""Host pipeline timeline of one headline step (DB200_TIMELINE=1): packed and ASCII host entry points, per-chunk copy / compute
intervals as the device saw them plus the host-side time of the call."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from dysgu_b200 import _lib, batch as B  # noqa: E402

L = _lib.lib(); L.db200_init(0)
q, t, prm, cells = bench.make_workload(524288, 20261021)
n = len(q)
prm_c = _lib.SwParams(); prm_c.match, prm_c.mismatch, prm_c.gap_open, prm_c.gap_extend = 2, -8, 6, 1; prm_c.score_size, prm_c.flag = 2, 1
pa = lambda shape, dt: bench.pinned_array(L, shape, dt)
pq, pt = B.pack_seqs(q, pinned=True), B.pack_seqs(t, pinned=True)
hql = pa((n,), np.int32); hql[:] = pq.lengths
htl = pa((n,), np.int32); htl[:] = pt.lengths
hquo = pa((n + 1,), np.int64); hquo[:] = pq.unit_off
htuo = pa((n + 1,), np.int64); htuo[:] = pt.unit_off
hq = pa((int(q.off[-1]),), np.uint8); hq[:] = q.buf
ht = pa((int(t.off[-1]),), np.uint8); ht[:] = t.buf
hqo = pa((n + 1,), np.int64); hqo[:] = q.off
hto = pa((n + 1,), np.int64); hto[:] = t.off
cap = 4 * n + 1024
hres = pa((n, 8), np.int32); hvar = pa((cap,), np.uint32); hvoff = pa((n + 1,), np.int64); hstat = pa((n,), np.int32)


def packed():
    return L.db200_sw_batch_host_packed(0, C.byref(prm_c), pq.units.ctypes.data, hquo.ctypes.data, hql.ctypes.data, pt.units.ctypes.data,
                                        htuo.ctypes.data, htl.ctypes.data, n, None, hres.ctypes.data, hvar.ctypes.data, cap, hvoff.ctypes.data,
                                        hstat.ctypes.data)


def ascii_():
    return L.db200_sw_batch_host(0, C.byref(prm_c), hq.ctypes.data, hqo.ctypes.data, ht.ctypes.data, hto.ctypes.data, n, None, hres.ctypes.data,
                                 hvar.ctypes.data, cap, hvoff.ctypes.data, hstat.ctypes.data)


for name, fn in (("packed", packed), ("ascii", ascii_)):
    for _ in range(3):
        assert fn() == 0
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, "ms per call:", " ".join("%.2f" % x for x in ts), flush=True)
    os.environ["DB200_TIMELINE"] = "1"
    sys.stderr.write("---- timeline %s\n" % name); sys.stderr.flush()
    fn()
    del os.environ["DB200_TIMELINE"]
