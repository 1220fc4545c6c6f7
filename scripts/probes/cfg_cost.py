"""Forward-pass cost of single (K, G) configurations of sw_pass_kernel: targets of exactly 2*G*K columns (no padding),
windows of 1500 rows, score only.  DB200_SW_FORCE_CFG pins the configuration; prints cells per second of the forward stage
and the cost per warp row this implies.  Run on a GPU box: python scripts/probes/cfg_cost.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from dysgu_b200 import _lib, batch, workloads  # noqa: E402

os.environ["DB200_SW_FINE_MIN_PAIRS"] = "0"
L = _lib.lib()
L.db200_set_host_chunk_pairs(0, 1 << 22)   # one chunk per call: every bucket kernel runs once over the whole batch
rng = np.random.default_rng(5)
CFGS = [(16, 1), (8, 1), (12, 1), (18, 1), (16, 2), (10, 2), (13, 2), (18, 2), (16, 3), (16, 4), (10, 4), (13, 4), (18, 4), (16, 5), (16, 6), (16, 7),
        (16, 8), (10, 8), (12, 8), (14, 8), (18, 8), (16, 10), (16, 16), (10, 16)]
if len(sys.argv) > 1:
    CFGS = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for K, G in CFGS:
    n = 2 * K * G
    m = 1500
    pairs = int(min(2.0e10 / (m * n), 600000))
    wl = np.full(pairs, m, dtype=np.int64)
    windows = workloads.random_dna(rng, wl)
    start = (rng.random(pairs) * (m - n)).astype(np.int64)
    clips = workloads.mutate(rng, workloads.slices(windows, start, np.full(pairs, n, dtype=np.int64)), 0.01, 0.0, 0.0)
    assert (clips.lengths == n).all()
    os.environ["DB200_SW_FORCE_CFG"] = "%d,%d" % (K, G)
    kw = dict(match=2, mismatch=-8, gap_open=6, gap_extend=1, flag=0, device=0)
    batch.sw_align_batch(windows, clips, **kw)
    L.db200_profile_enable(0, 1)
    reps = 3
    for _ in range(reps):
        batch.sw_align_batch(windows, clips, **kw)
    st = (C.c_float * 16)()
    L.db200_profile_collect(0, st, 16)
    L.db200_profile_enable(0, 0)
    fwd_ms = float(st[3]) / reps
    cells = float(pairs) * m * n
    tcups = cells / (fwd_ms * 1e-3) / 1e12
    # warp rows per second: pairs * (m + 2G - 1) / (32 // G) per launch
    warp_rows = pairs * (m + 2 * G - 1) / (32 // G)
    ns_per_warp_row = fwd_ms * 1e6 / warp_rows
    # per SM: 148 SMs, slots: ns * 1.965 GHz * 2 ALU warp-instructions per clock per SM
    slots = ns_per_warp_row * 148 * 1.965 * 2
    print("K=%2d G=%2d n=%4d pairs=%6d fwd %.3f ms  %.2f TCUPS  frac %.3f  ALU slots per warp row %.1f (recurrence %d)" %
          (K, G, n, pairs, fwd_ms, tcups, tcups * 3 / 18.136, slots, 6 * K), flush=True)
