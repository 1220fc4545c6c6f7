"""GPU: the other host input forms of the Smith-Waterman call give the results of the plain CSR call, field by field and
cigar by cigar.
  * slices (db200_sw_batch_host_slices): windows cut out of one shared reference region, overlapping freely
    (re_map.py:429-434), targets as CSR or as slices too;
  * pre-packed (db200_sw_batch_host_packed): both sides in the engine's 4-bit pool format written by db200_sw_pack_host.
Large batches (chunked host pipeline), batches of a few pairs (the one-launch path materialises them) and odd cases."""
import numpy as np
import pytest

from dysgu_b200 import batch, workloads
from dysgu_b200.workloads import SeqBatch

pytestmark = pytest.mark.gpu


def same(a, b):
    assert np.array_equal(a.status, b.status)
    assert np.array_equal(a.fixed, b.fixed)
    assert np.array_equal(a.cigar_off, b.cigar_off) and np.array_equal(a.cigar, b.cigar)


def region_workload(n, seed, region_len=200000):
    rng = np.random.default_rng(seed)
    region = workloads.BASES[rng.integers(0, 4, size=region_len, dtype=np.uint8)]
    wl = rng.choice(np.array([1000, 1500, 2000]), size=n)
    start = rng.integers(0, region_len - 2000, size=n)
    cl = rng.integers(15, 200, size=n)
    cstart = start + (rng.random(n) * (wl - cl)).astype(np.int64)
    win = SeqBatch(np.concatenate([region[s:s + w] for s, w in zip(start, wl)]), np.concatenate(([0], np.cumsum(wl))))
    clips = workloads.lower(workloads.mutate(rng, SeqBatch(np.concatenate([region[s:s + c] for s, c in zip(cstart, cl)]),
                                                          np.concatenate(([0], np.cumsum(cl)))), 0.02, 0.004, 0.004))
    return region, start.astype(np.int64), wl.astype(np.int32), win, clips


@pytest.mark.parametrize("n", [1, 9, 64, 65, 5000, 70000])
def test_slices_equal_csr(n):
    region, start, wl, win, clips = region_workload(n, 100 + n)
    prm = dict(match=2, mismatch=-8, gap_open=6, gap_extend=1)
    want = batch.sw_align_batch(win, clips, **prm)
    got = batch.sw_align_slices(region, start, wl, clips, **prm)
    same(got, want)
    assert (want.fixed[:, 0] > 30).mean() > 0.8


def test_both_sides_as_slices_and_masks():
    region, start, wl, win, clips = region_workload(3000, 7)
    rng = np.random.default_rng(8)
    tl = rng.integers(20, 400, size=3000).astype(np.int32)
    ts = (start + rng.integers(0, 600, size=3000)).astype(np.int64)
    tg = SeqBatch(np.concatenate([region[s:s + c] for s, c in zip(ts, tl)]), np.concatenate(([0], np.cumsum(tl))))
    prm = dict(match=2, mismatch=-3, gap_open=5, gap_extend=2, mask_len=15)
    want = batch.sw_align_batch(win, tg, **prm)
    got = batch.sw_align_slices(region, start, wl, None, tsrc=region, tstart=ts, tlen=tl, **prm)
    same(got, want)


@pytest.mark.parametrize("n", [1, 33, 64, 3000, 90000])
def test_packed_equals_csr(n):
    q, t, prm = workloads.remap_windows(n, seed=200 + n)
    want = batch.sw_align_batch(q, t, **prm)
    pq, pt = batch.pack_seqs(q), batch.pack_seqs(t, pinned=True)
    assert pq.nbytes < 0.56 * q.off[-1]
    got = batch.sw_align_packed(pq, pt, **prm)
    same(got, want)


def test_packed_odd_sequences():
    qs = [b"ACGTNNNNacgtUu-*xyz", b"A", b"ACGT" * 8, b"ACGT" * 8 + b"A", b"G" * 31, b"T" * 33, b"ACGTACGTTGCA" * 30]
    ts = [b"ACGTNNNNACGTTT", b"A", b"ACGT" * 8, b"TACG" * 9, b"G" * 64, b"T" * 32, b"ACGTACGATGCA" * 28]
    want = batch.sw_align_batch(qs, ts, match=2, mismatch=-3, gap_open=5, gap_extend=2)
    got = batch.sw_align_packed(batch.pack_seqs(qs), batch.pack_seqs(ts), match=2, mismatch=-3, gap_open=5, gap_extend=2)
    same(got, want)
    long_q, long_t, prm = workloads.contig_pairs(40, seed=5, platform="ont")      # beyond the one-launch path: pipeline with column tiles
    same(batch.sw_align_packed(batch.pack_seqs(long_q), batch.pack_seqs(long_t), **prm), batch.sw_align_batch(long_q, long_t, **prm))


def test_bad_packed_offsets_are_refused():
    from dysgu_b200 import _lib
    pq, pt = batch.pack_seqs([b"ACGT" * 20]), batch.pack_seqs([b"ACGT" * 20])
    pq.unit_off[1] += 1
    with pytest.raises(_lib.AlignEngineError, match="do not match"):
        batch.sw_align_packed(pq, pt)
