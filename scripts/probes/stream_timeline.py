"""Streaming end-to-end arm of bench.py (NT host threads, one engine instance each, whole-batch calls) with DB200_TIMELINE=1:
per call, when its copy and its kernels started / ended on the device and when the host had everything enqueued."""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from dysgu_b200 import _lib, batch as B  # noqa: E402

NT = int(os.environ.get("NT", "3"))
STEPS = int(os.environ.get("STEPS", "8"))
L = _lib.lib(); L.db200_init(0)
q, t, prm, cells = bench.make_workload(524288, 20261021)
order = np.argsort((t.lengths + 31) // 32, kind="stable")
q, t = q.take(order), t.take(order)
n = len(q)
prm_c = _lib.SwParams(); prm_c.match, prm_c.mismatch, prm_c.gap_open, prm_c.gap_extend = 2, -8, 6, 1; prm_c.score_size, prm_c.flag = 2, 1
pa = lambda shape, dt: bench.pinned_array(L, shape, dt)
pq, pt = B.pack_seqs(q, pinned=True), B.pack_seqs(t, pinned=True)
hql = pa((n,), np.int32); hql[:] = pq.lengths
htl = pa((n,), np.int32); htl[:] = pt.lengths
hquo = pa((n + 1,), np.int64); hquo[:] = pq.unit_off
htuo = pa((n + 1,), np.int64); htuo[:] = pt.unit_off
cap = 4 * n + 1024
outs = [(pa((n, 8), np.int32), pa((cap,), np.uint32), pa((n + 1,), np.int64), pa((n,), np.int32)) for _ in range(NT)]


def step(j):
    r_, v_, o_, s_ = outs[j]
    rc = L.db200_sw_batch_host_packed(j << 8, C.byref(prm_c), pq.units.ctypes.data, hquo.ctypes.data, hql.ctypes.data, pt.units.ctypes.data,
                                      htuo.ctypes.data, htl.ctypes.data, n, None, r_.ctypes.data, v_.ctypes.data, cap, o_.ctypes.data, s_.ctypes.data)
    assert rc == 0


for j in range(NT):
    L.db200_set_host_chunk_pairs(j << 8, C.c_int64(n))
    step(j); step(j)


def run(j, steps, stamps):
    for k in range(j, steps, NT):
        t0 = time.perf_counter(); step(j); stamps.append((k, j, t0, time.perf_counter()))


for tl in (False, True):
    if tl:
        os.environ["DB200_TIMELINE"] = "1"
    stamps = []
    ths = [threading.Thread(target=run, args=(j, NT * STEPS, stamps)) for j in range(NT)]
    T0 = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    dt = time.perf_counter() - T0
    print("NT=%d steps=%d timeline=%s: %.2f ms per step, %.0f GCUPS" % (NT, NT * STEPS, tl, 1e3 * dt / (NT * STEPS), cells * NT * STEPS / dt / 1e9), flush=True)
    stamps.sort()
    print(" call: start-end (ms since T0) " + "  ".join("%d:t%d %.1f-%.1f" % (k, j, 1e3 * (a - T0), 1e3 * (b - T0)) for k, j, a, b in stamps[:12]), flush=True)
