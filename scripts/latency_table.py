#!/usr/bin/env python
"""Latency of the synchronous call -> one JSON object.

us per call for batches of n = 1, 8, 64, 512, 4096 pairs of the reference's 150 x 100 micro-benchmark shape
(134 us per Python-level call) and of the headline shape (window 1-2 kb vs clip), through
  * the Python drop-in (StripedSmithWaterman(q)(t) in a loop for n = 1; batch.sw_align_batch for n pairs),
  * the C ABI (db200_sw_batch_host with preallocated numpy buffers, ctypes call only),
beside the reference's own Python-level call timed on this box's CPU (oracle/_ref/pyref: _ssw_wrapper.pyx:613-648).
Also edlib.align (HW / locations, clip vs 1 kb window)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dysgu_b200 import _lib, batch, workloads  # noqa: E402
from dysgu_b200.edlib import align as gpu_align  # noqa: E402
from dysgu_b200.scikitbio import StripedSmithWaterman  # noqa: E402


def timeit(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


def micro_pairs(n, seed):
    rng = np.random.default_rng(seed)
    q = workloads.random_dna(rng, np.full(n, 150))
    t = workloads.random_dna(rng, np.full(n, 100))
    # related: the target is a slice of the query with a few substitutions
    tb = []
    for i in range(n):
        s = bytearray(q[i][20:120])
        for k in rng.integers(0, 100, 3):
            s[k] = b"ACGT"[rng.integers(0, 4)]
        tb.append(bytes(s))
    return q, workloads.SeqBatch.from_list(tb)


def c_abi_call(q, t, prm):
    L = _lib.lib()
    n = len(q)
    p = _lib.SwParams()
    p.match, p.mismatch, p.gap_open, p.gap_extend = prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"]
    p.score_size, p.flag = 2, 1
    fixed = np.zeros((n, 8), np.int32); status = np.zeros(n, np.int32); coff = np.zeros(n + 1, np.int64)
    cap = int(q.off[-1] + t.off[-1] + 2 * n + 16)
    cig = np.zeros(cap, np.uint32)
    args = (batch.default_device(), C.byref(p), q.buf.ctypes.data, q.off.ctypes.data, t.buf.ctypes.data, t.off.ctypes.data, n, None,
            fixed.ctypes.data, cig.ctypes.data, cap, coff.ctypes.data, status.ctypes.data)
    keep = (p, fixed, status, coff, cig, q, t)   # the raw addresses in `args` must outlive this function
    return lambda keep=keep: L.db200_sw_batch_host(*args)


def main():
    out = {"unit": "us per call (whole batch)", "shapes": {}}
    ref_cls = None
    try:
        import importlib.util, sysconfig
        path = os.path.join(ROOT, "oracle", "_ref", "pyref", "dysgu", "scikitbio", "_ssw_wrapper" + sysconfig.get_config_var("EXT_SUFFIX"))
        spec = importlib.util.spec_from_file_location("pyref_ssw._ssw_wrapper", path)
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
        ref_cls = m.StripedSmithWaterman
    except Exception as exc:   # noqa: BLE001
        out["reference_wrapper"] = "unavailable: %s" % exc
    for shape, gen, prm in (("micro_150x100", micro_pairs, dict(match=2, mismatch=-3, gap_open=5, gap_extend=2)),
                            ("remap_window_vs_clip", lambda n, s: workloads.remap_windows(n, seed=s)[:2], dict(match=2, mismatch=-8, gap_open=6, gap_extend=1))):
        rows = {}
        q1, t1 = gen(64, 1)
        qs = [x.decode() for x in q1.tolist()]; ts = [x.decode() for x in t1.tolist()]
        kw = dict(match_score=prm["match"], mismatch_score=prm["mismatch"], gap_open_penalty=prm["gap_open"], gap_extend_penalty=prm["gap_extend"])
        it = [0]

        def dropin():
            i = it[0] = (it[0] + 1) % 64
            return StripedSmithWaterman(qs[i], **kw)(ts[i]).optimal_alignment_score
        rows["python_dropin_init_plus_call_n1"] = round(timeit(dropin, 300), 1)
        if ref_cls is not None:
            def refcall():
                i = it[0] = (it[0] + 1) % 64
                return ref_cls(qs[i], **kw)(ts[i]).optimal_alignment_score
            rows["reference_python_call_n1_cpu"] = round(timeit(refcall, 300), 1)
        for n in (1, 8, 64, 512, 4096):
            q, t = gen(n, 2)
            reps = 200 if n <= 64 else 30
            rows["python_batch_n%d" % n] = round(timeit(lambda: batch.sw_align_batch(q, t, **prm), reps), 1)
            rows["c_abi_n%d" % n] = round(timeit(c_abi_call(q, t, prm), reps), 1)
        if ref_cls is not None:
            ref_us = rows["reference_python_call_n1_cpu"]
            be = next((n for n in (1, 8, 64, 512, 4096) if rows["python_batch_n%d" % n] / n < ref_us), None)
            rows["break_even_batch_size_vs_reference_python_call"] = be
        out["shapes"][shape] = rows
    c, w, eprm = workloads.remap_chain(64, seed=3)
    cs = [x.decode() for x in c.tolist()]; ws = [x.decode() for x in w.tolist()]
    it = [0]

    def ed():
        i = it[0] = (it[0] + 1) % 64
        return gpu_align(cs[i], ws[i], "HW", "locations")["editDistance"]
    out["edlib_HW_locations_clip_vs_window_n1_python_dropin"] = round(timeit(ed, 200), 1)
    for n in (1, 64, 4096):
        c, w, eprm = workloads.remap_chain(n, seed=4)
        out["edlib_python_batch_n%d" % n] = round(timeit(lambda: batch.ed_align_batch(c, w, **eprm), 50), 1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
