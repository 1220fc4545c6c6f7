"""GPU: the call patterns dysgu's processes and threads produce at the boundary (SURVEY 8b "Threading"): several host
threads calling the batch entry points at once (edlib.pyx:129 releases the GIL, ctypes does too), and fork()ed pool workers
(merge_svs.pyx:396) that create their engine lazily in the child."""
import os
import subprocess
import sys
import textwrap
import threading

import numpy as np
import pytest

from dysgu_b200 import batch, workloads

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_threads_share_one_device():
    jobs = []
    for s in range(6):
        if s % 2 == 0:
            q, t, prm = workloads.remap_windows(3000, seed=100 + s)
            jobs.append(("sw", q, t, prm))
        else:
            c, w, prm = workloads.remap_chain(3000, seed=100 + s)
            jobs.append(("ed", c, w, dict(prm, task="path" if s == 3 else prm["task"])))

    def run(job):
        kind, a, b, prm = job
        return batch.sw_align_batch(a, b, **prm) if kind == "sw" else batch.ed_align_batch(a, b, **prm)

    serial = [run(j) for j in jobs]
    out = [None] * len(jobs)
    errs = []

    def worker(i):
        try:
            for _ in range(3):
                out[i] = run(jobs[i])
        except Exception as exc:   # noqa: BLE001
            errs.append(exc)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs
    for (kind, *_), a, b in zip(jobs, serial, out):
        assert np.array_equal(a.fixed, b.fixed)
        if kind == "sw":
            assert np.array_equal(a.cigar, b.cigar) and np.array_equal(a.cigar_off, b.cigar_off)
        else:
            assert np.array_equal(a.loc, b.loc) and np.array_equal(a.loc_off, b.loc_off)
            if a.path is not None:
                assert np.array_equal(a.path, b.path) and np.array_equal(a.path_off, b.path_off)


FORK_SCRIPT = textwrap.dedent("""
    import multiprocessing as mp, sys
    sys.path.insert(0, %r)
    from dysgu_b200.scikitbio import StripedSmithWaterman
    from dysgu_b200.edlib import align

    def work(i):
        a = StripedSmithWaterman("ACGTTGCAACGTACGTTTGACCA", match_score=2, mismatch_score=-8, gap_open_penalty=6, gap_extend_penalty=1)("ttgcaacgtaggtttga")
        e = align("ACGTTGCA", "TTACGTAGCATT", "HW", "locations")
        return (a.optimal_alignment_score, a.cigar, e["editDistance"], e["locations"])

    if __name__ == "__main__":
        with mp.get_context("fork").Pool(2) as pool:      # the parent has not touched the GPU: each child creates its engine
            res = pool.map(work, range(4))
        assert res == [(24, "17M", 1, [(2, 9)])] * 4, res
        print("fork ok")
""")


def test_forked_pool_workers_create_their_own_engine():
    r = subprocess.run([sys.executable, "-c", FORK_SCRIPT % ROOT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "fork ok" in r.stdout, (r.stdout, r.stderr[-2000:])
