#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json workload shapes through the host C ABI (not bench lines;
bench.py carries the headline).  Prints one JSON object per shape: GPU time / throughput, and the reference CPU
cores (oracle/_ref) on all host threads over a bounded sample, with a parity count on that sample.

    python scripts/bench_configs.py [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dysgu_b200 import batch, workloads  # noqa: E402


def cells(q, t):
    return float((q.lengths.astype(np.float64) * t.lengths.astype(np.float64)).sum())


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t0)
    return best, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    try:
        from oracle import oracle as O
        O.driver_lib()
    except Exception:
        O = None
    cores = os.cpu_count() or 1
    S = a.scale
    shapes = []
    q, t, p = workloads.remap_chain(int(400000 * S), seed=1)
    shapes.append(("2a edlib HW/locations clip vs 1000 bp window (re_map.py:209)", "ed", q, t, p))
    q, t, p = workloads.pe_contig_pairs(int(200000 * S), seed=2)
    shapes.append(("PE contigs 100-500 bp check_contig_match SSW (2,-3,10,1)", "sw", q, t, p))
    q, t, p = workloads.insertion_linking(int(20000 * S), seed=3, platform="hifi")
    shapes.append(("3 HiFi insertion sequences 50 bp-10 kb edit distance NW", "ed", q, t, p))
    shapes.append(("3 HiFi insertion sequences HW/locations", "ed", q, t, dict(mode="HW", task="locations")))
    q, t, p = workloads.insertion_linking(int(20000 * S), seed=4, platform="ont")
    shapes.append(("4 ONT r10 insertion sequences edit distance NW (~5% error)", "ed", q, t, p))
    q, t, p = workloads.contig_pairs(int(4000 * S), seed=5, platform="ont")
    shapes.append(("4/5 long-read contigs 1-10 kb check_contig_match SSW (2,-3,10,1), cigar", "sw", q, t, p))
    for name, kind, q, t, p in shapes:
        c = cells(q, t)
        q, t = batch.pinned_copy(q), batch.pinned_copy(t)
        if kind == "sw":
            dt, r = timed(lambda: batch.sw_align_batch(q, t, **p))
        else:
            dt, r = timed(lambda: batch.ed_align_batch(q, t, **p))
        out = {"shape": name, "pairs": len(q), "gcells": c / 1e9, "gpu_s": dt, "gpu_gcups_e2e": c / dt / 1e9, "pairs_per_s": len(q) / dt}
        if O is not None and not a.no_cpu:
            n = min(len(q), max(200, int(len(q) * 0.1)))
            qs, ts = q.take(np.arange(n)), t.take(np.arange(n))
            cs = cells(qs, ts)
            if kind == "sw":
                rr = O.ref_sw((qs.buf, qs.off), (ts.buf, ts.off), p["match"], p["mismatch"], p["gap_open"], p["gap_extend"], nthreads=cores, csr=True)
                mism = int((rr.fixed != r.fixed[:n]).any(axis=1).sum())
            else:
                rr = O.ref_ed((qs.buf, qs.off), (ts.buf, ts.off), p["mode"], p["task"], nthreads=cores, csr=True)
                mism = int((rr.fixed != r.fixed[:n]).any(axis=1).sum())
            out.update({"cpu_cores": cores, "cpu_sample_pairs": n, "cpu_s": rr.seconds, "cpu_gcups": cs / rr.seconds / 1e9,
                        "speedup_e2e": (c / dt) / (cs / rr.seconds), "parity_mismatches_on_sample": mism})
        print(json.dumps(out))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
