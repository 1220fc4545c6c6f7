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


# The order `dysgu call -p N` really produces: the PARENT aligns first (re_map.remap_soft_clips, cluster.pyx:475-478), then
# forks the pool of merge_svs.pyx:396 whose workers align again.  CUDA cannot be used in such children; the parent's
# alignment server (dysgu_b200/server.py) runs their batches on the parent's engine.
FORK_AFTER_INIT_SCRIPT = textwrap.dedent("""
    import multiprocessing as mp, os, sys
    sys.path.insert(0, %r)
    from dysgu_b200.scikitbio import StripedSmithWaterman
    from dysgu_b200.edlib import align
    from dysgu_b200 import _lib, batch, workloads

    def work(i):
        a = StripedSmithWaterman("ACGTTGCAACGTACGTTTGACCA", match_score=2, mismatch_score=-8, gap_open_penalty=6, gap_extend_penalty=1)("ttgcaacgtaggtttga")
        e = align("ACGTTGCA", "TTACGTAGCATT", "HW", "locations")
        q, t, prm = workloads.remap_windows(300, seed=40 + i)
        r = batch.sw_align_batch(q, t, **prm)
        return (a.optimal_alignment_score, a.cigar, e["editDistance"], e["locations"], int(r.fixed.sum()), int(r.cigar.sum()),
                int(_lib.lib().db200_forked_after_init()), os.getpid())

    if __name__ == "__main__":
        q, t, prm = workloads.remap_windows(2000, seed=1)
        batch.sw_align_batch(q, t, **prm)                     # the parent touches the GPU first
        want = []
        for i in range(4):
            q, t, prm = workloads.remap_windows(300, seed=40 + i)
            r = batch.sw_align_batch(q, t, **prm)
            want.append((int(r.fixed.sum()), int(r.cigar.sum())))
        with mp.get_context("fork").Pool(2) as pool:
            res = pool.map(work, range(4))
        for i, r in enumerate(res):
            assert r[:4] == (24, "17M", 1, [(2, 9)]), r
            assert r[4:6] == want[i], (r, want[i])
            assert r[6] == 1 and r[7] != os.getpid(), r          # really a forked child that cannot use CUDA itself
        print("fork after init ok")
""")


def test_pool_forked_after_the_parent_used_the_gpu_is_served_by_the_parent():
    r = subprocess.run([sys.executable, "-c", FORK_AFTER_INIT_SCRIPT % ROOT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "fork after init ok" in r.stdout, (r.stdout, r.stderr[-3000:])


FORK_NO_SERVER_SCRIPT = textwrap.dedent("""
    import multiprocessing as mp, sys
    sys.path.insert(0, %r)
    from dysgu_b200 import _lib, batch

    def work(i):
        try:
            batch.sw_align_batch([b"ACGTACGTACGT"], [b"ACGTACGTACGT"])
        except _lib.AlignEngineError as exc:
            return str(exc)
        return "no error"

    if __name__ == "__main__":
        batch.sw_align_batch([b"ACGTACGTACGT"], [b"ACGTACGTACGT"])
        with mp.get_context("fork").Pool(1) as pool:
            msg = pool.map(work, [0])[0]
        assert "fork()ed after its parent initialised CUDA" in msg and "code -6" in msg, msg
        print("clear error ok")
""")


def test_without_the_server_a_forked_child_gets_a_clear_error():
    env = dict(os.environ, DB200_NO_SERVER="1")
    r = subprocess.run([sys.executable, "-c", FORK_NO_SERVER_SCRIPT % ROOT], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "clear error ok" in r.stdout, (r.stdout, r.stderr[-3000:])


def test_engine_instances_run_side_by_side_on_one_gpu():
    """DB200_INSTANCE(dev, k): independent engines on one GPU, one host thread each (bench.py's streaming e2e arm)."""
    dev = batch.default_device()
    q, t, prm = workloads.remap_windows(40000, seed=77)
    pq, pt = batch.pack_seqs(q), batch.pack_seqs(t)
    want = batch.sw_align_batch(q, t, device=dev, **prm)
    out, errs = [None, None, None], []

    def work(j):
        try:
            for _ in range(3):
                out[j] = batch.sw_align_packed(pq, pt, device=dev | (j << 8), **prm) if j else batch.sw_align_batch(q, t, device=dev | (2 << 8), **prm)
        except Exception as exc:   # noqa: BLE001
            errs.append(exc)

    ths = [threading.Thread(target=work, args=(j,)) for j in range(3)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errs, errs
    for r in out:
        assert np.array_equal(r.fixed, want.fixed) and np.array_equal(r.cigar, want.cigar) and np.array_equal(r.cigar_off, want.cigar_off)
