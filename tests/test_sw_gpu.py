"""GPU parity of the Smith-Waterman path against the CPU oracle (bit-exact, all fields + cigars).

Calls go through the C ABI (db200_sw_batch_host) via dysgu_b200.batch.
"""
import numpy as np
import pytest

import pairgen
from oracle import oracle as O
from dysgu_b200 import batch, workloads

pytestmark = pytest.mark.gpu

SCORINGS = [(2, -8, 6, 1), (2, -3, 10, 1), (2, -3, 5, 1), (2, -3, 4, 1), (2, -3, 5, 2), (3, -2, 7, 3)]


def compare(qs, ts, res, ref, tag, limit=5):
    bad = []
    for i in range(len(qs)):
        a = res.as_tuple(i) if res.status[i] == 0 else ("status", int(res.status[i]))
        b = ref.as_tuple(i) if ref.status[i] == 0 else ("status", int(ref.status[i]))
        if a != b:
            bad.append((i, a, b))
    msg = "\n".join("%s pair %d q=%r t=%r\n   gpu=%r\n   ref=%r" % (tag, i, qs[i][:80], ts[i][:80], a, b) for i, a, b in bad[:limit])
    assert not bad, "%d/%d differ\n%s" % (len(bad), len(qs), msg)


@pytest.mark.parametrize("scoring", SCORINGS)
def test_random_classes(scoring):
    qs, ts = pairgen.sw_pairs(1000 + sum(scoring), 3000, 420)
    m, x, go, ge = scoring
    res = batch.sw_align_batch(qs, ts, m, x, go, ge)
    ref = O.oracle_sw(qs, ts, m, x, go, ge)
    compare(qs, ts, res, ref, str(scoring))


@pytest.mark.parametrize("score_size,flag", [(1, 1), (2, 0), (2, 8), (0, 1), (2, 1)])
def test_flags_and_score_size(score_size, flag):
    qs, ts = pairgen.sw_pairs(77 + score_size * 10 + flag, 1500, 200)
    res = batch.sw_align_batch(qs, ts, 2, -3, 5, 2, score_size=score_size, flag=flag)
    ref = O.oracle_sw(qs, ts, 2, -3, 5, 2, score_size=score_size, flag=flag)
    # score_size 0: the reference returns NULL on 8-bit overflow (oracle status -2, ours DB200_PAIR_OVERFLOW = 2)
    for i in range(len(qs)):
        if ref.status[i] == -2:
            assert res.status[i] == 2
            ref.status[i] = 2
            res.fixed[i] = 0
            ref.fixed[i] = 0
    compare(qs, ts, res, ref, "ss%d flag%d" % (score_size, flag))


def test_fixed_mask_len():
    qs, ts = pairgen.sw_pairs(4242, 1500, 300)
    for ml in (15, 3, 40):
        res = batch.sw_align_batch(qs, ts, 2, -3, 5, 2, mask_len=ml)
        ref = O.oracle_sw(qs, ts, 2, -3, 5, 2, mask_len=np.full(len(qs), ml, np.int32))
        compare(qs, ts, res, ref, "mask%d" % ml)


def test_long_pairs_column_tiling():
    # targets longer than 1024 / 2048 columns exercise the column-tiled kernel and its boundary array
    a, b, prm = workloads.contig_pairs(24, seed=5, platform="ont", lo=50, hi=5000)
    qs, ts = a.tolist(), b.tolist()
    res = batch.sw_align_batch(qs, ts, **prm)
    ref = O.oracle_sw(qs, ts, prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"])
    compare(qs, ts, res, ref, "contigs")


def test_remap_window_shape():
    q, t, prm = workloads.remap_windows(4000, seed=11)
    res = batch.sw_align_batch(q, t, **prm)
    ref = O.oracle_sw((q.buf, q.off), (t.buf, t.off), prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"], csr=True)
    compare(q.tolist(), t.tolist(), res, ref, "remap")


def test_empty_sequences_are_reported():
    res = batch.sw_align_batch([b"", b"ACGT", b"ACGT"], [b"ACGT", b"", b"ACGT"], 2, -3, 5, 2)
    assert list(res.status) == [1, 1, 0]
    assert res.as_tuple(2) == (8, 0, 0, 3, 0, 3, 0, "4M")


def test_scratch_starved_pairs_are_retried():
    # with tiny traceback pools a batch of long, indel-rich contigs cannot hold every direction matrix at once:
    # the host entry point must re-run the starved pairs in small groups and still return bit-exact cigars
    import subprocess
    import sys
    import os
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from dysgu_b200 import batch, workloads
from oracle import oracle as O
a, b, prm = workloads.contig_pairs(300, seed=17, platform="ont", lo=1500, hi=4000)
qs, ts = a.tolist(), b.tolist()
res = batch.sw_align_batch(qs, ts, **prm)
ref = O.oracle_sw(qs, ts, prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"])
bad = sum(1 for i in range(len(qs)) if res.status[i] != 0 or res.as_tuple(i) != ref.as_tuple(i))
print("BAD", bad)
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DB200_TRACE_SLAB_MB="64")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert "BAD 0" in out.stdout, out.stdout[-500:] + out.stderr[-1500:]


def test_column_tiled_pairs_with_tied_ends():
    # targets of more than 2048 columns run the column-tiled kernel; low-complexity tails make several columns of
    # one virtual lane reach the maximum, which exercises the single-pass re-run from the kept boundary columns
    import random
    rng = random.Random(2024)
    qs, ts = [], []
    for _ in range(160):
        core = pairgen.rand_dna(rng, rng.randint(2080, 4300))
        unit = rng.choice(["A", "AC", "T", "GT", "CAG"])
        tail_q = (unit * 40)[:rng.randint(0, 30)]
        tail_t = (unit * 40)[:rng.randint(0, 30)]
        q = core + tail_q
        t = pairgen.mutate(rng, core, 0.01, 0.004, 0.004) + tail_t
        if rng.random() < 0.5:
            q, t = t, q
        qs.append(q.encode()); ts.append(t.encode())
    res = batch.sw_align_batch(qs, ts, 2, -3, 10, 1)
    ref = O.oracle_sw(qs, ts, 2, -3, 10, 1)
    compare(qs, ts, res, ref, "tiled ties")


def test_pipelined_column_tiles_equal_sequential_tiles():
    """Targets beyond 2048 columns: the passes of one pair run as a pipeline of warps (sw_pass_kernel PIPE);
    DB200_SW_NO_PIPE=1 runs them one after the other in one warp.  Both must agree with each other on every field
    (with and without column maxima, i.e. small and automatic mask lengths) and with the oracle."""
    import os
    a, b, prm = workloads.contig_pairs(160, seed=23, platform="ont", lo=1200, hi=9000)
    qs, ts = a.tolist(), b.tolist()
    # some pairs swapped so that long targets meet short queries and vice versa, some unrelated
    for i in range(0, 160, 5):
        qs[i], ts[i] = ts[i], qs[i]
    for i in range(3, 160, 16):
        ts[i] = ts[(i * 7 + 1) % 160]
    for mask in (None, 15):
        ml = None if mask is None else np.full(len(qs), mask, np.int32)
        piped = batch.sw_align_batch(qs, ts, mask_len=ml, **prm)
        os.environ["DB200_SW_NO_PIPE"] = "1"
        try:
            seq = batch.sw_align_batch(qs, ts, mask_len=ml, **prm)
        finally:
            del os.environ["DB200_SW_NO_PIPE"]
        assert (piped.status == 0).all() and (seq.status == 0).all()
        assert np.array_equal(piped.fixed, seq.fixed), np.nonzero((piped.fixed != seq.fixed).any(axis=1))[0][:10]
        assert np.array_equal(piped.cigar, seq.cigar) and np.array_equal(piped.cigar_off, seq.cigar_off)
    sub = list(range(0, 160, 8))
    res = batch.sw_align_batch([qs[i] for i in sub], [ts[i] for i in sub], **prm)
    ref = O.oracle_sw([qs[i] for i in sub], [ts[i] for i in sub], prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"])
    compare([qs[i] for i in sub], [ts[i] for i in sub], res, ref, "pipelined tiles")


def test_device_entry_point_sub_batches_share_one_cigar_csr():
    """DB200_DEVICE_SUBBATCHES (off by default) cuts a device-resident batch into sub-batches on the engine's slot streams;
    the outputs - fixed records, status, and the cigar CSR the sub-batches append to in order - must not change."""
    import ctypes as C
    import os
    import torch
    from dysgu_b200 import _lib
    L = _lib.lib()
    q, t, prm = workloads.remap_windows(30000, seed=3)
    n, qb, tb = len(q), int(q.off[-1]), int(t.off[-1])
    p = _lib.SwParams()
    p.match, p.mismatch, p.gap_open, p.gap_extend, p.score_size, p.flag = prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"], 2, 1
    dev = torch.device("cuda", 0)
    dq, dt = torch.from_numpy(q.buf).to(dev), torch.from_numpy(t.buf).to(dev)
    dqo, dto = torch.from_numpy(q.off).to(dev), torch.from_numpy(t.off).to(dev)
    cap = 4 * n + 1024
    outs = {}
    for nsub in (1, 3, 7):
        os.environ["DB200_DEVICE_SUBBATCHES"] = str(nsub)
        try:
            res = torch.zeros((n, 8), dtype=torch.int32, device=dev)
            cig = torch.zeros(cap, dtype=torch.int32, device=dev)
            coff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            st = torch.full((n,), -1, dtype=torch.int32, device=dev)
            total = C.c_int64(-1)
            rc = L.db200_sw_batch_device(0, C.byref(p), dq.data_ptr(), dqo.data_ptr(), qb, dt.data_ptr(), dto.data_ptr(), tb, n,
                                         int(q.lengths.max()), int(t.lengths.max()), None, res.data_ptr(), cig.data_ptr(), cap,
                                         coff.data_ptr(), st.data_ptr(), C.addressof(total), None)
            _lib.check(rc, "db200_sw_batch_device")
            torch.cuda.synchronize()
            outs[nsub] = (res.cpu(), cig.cpu(), coff.cpu(), st.cpu(), total.value)
        finally:
            del os.environ["DB200_DEVICE_SUBBATCHES"]
    ref = outs[1]
    assert ref[4] == int(ref[2][-1]) and (ref[3] == 0).all()
    for nsub in (3, 7):
        for a, b in zip(ref[:4], outs[nsub][:4]):
            assert torch.equal(a, b), nsub
        assert outs[nsub][4] == ref[4]
    oracle = O.oracle_sw((q.buf, q.off), (t.buf, t.off), prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"], csr=True)
    assert np.array_equal(ref[0].numpy(), oracle.fixed)


# ---- fine-grained kernel configurations (K = 8..18 columns per virtual lane, sw_launch.cuh) --------------------------
# Calls of >= 196,608 pairs choose them per pair by cost; the tests switch them on for small batches (DB200_SW_FINE_MIN_PAIRS=0)
# and compare with the oracle and with the classic configurations.

@pytest.mark.parametrize("scoring", SCORINGS)
def test_fine_configurations_random_classes(scoring, monkeypatch):
    monkeypatch.setenv("DB200_SW_FINE_MIN_PAIRS", "0")
    qs, ts = pairgen.sw_pairs(9000 + sum(scoring), 3000, 700)
    m, x, go, ge = scoring
    res = batch.sw_align_batch(qs, ts, m, x, go, ge)
    ref = O.oracle_sw(qs, ts, m, x, go, ge)
    compare(qs, ts, res, ref, "fine " + str(scoring))


def test_fine_configurations_every_target_length(monkeypatch):
    # one related and one unrelated pair for every target length 1..1024: every (K, G) and every padding remainder;
    # a fixed short mask asks for the column maxima variants (second best) as well, flag 0 for the score-only route
    import random
    rng = random.Random(31)
    qs, ts = [], []
    for n in range(1, 1025):
        t = pairgen.rand_dna(rng, n)
        q = pairgen.rand_dna(rng, rng.randint(0, 30)) + pairgen.mutate(rng, t, 0.03, 0.01, 0.01) + pairgen.rand_dna(rng, rng.randint(0, 30))
        qs.append(q.encode()); ts.append(t.encode())
        qs.append(pairgen.rand_dna(rng, rng.randint(20, 200)).encode()); ts.append(t.lower().encode())
    monkeypatch.setenv("DB200_SW_FINE_MIN_PAIRS", "0")
    for kw in (dict(), dict(mask_len=15), dict(flag=0)):
        res = batch.sw_align_batch(qs, ts, 2, -8, 6, 1, **kw)
        okw = dict(kw)
        if "mask_len" in okw:
            okw["mask_len"] = np.full(len(qs), okw["mask_len"], np.int32)
        ref = O.oracle_sw(qs, ts, 2, -8, 6, 1, **okw)
        compare(qs, ts, res, ref, "fine lengths %r" % (kw,))
    monkeypatch.setenv("DB200_SW_NO_FINE", "1")
    classic = batch.sw_align_batch(qs, ts, 2, -8, 6, 1)
    monkeypatch.delenv("DB200_SW_NO_FINE")
    fine = batch.sw_align_batch(qs, ts, 2, -8, 6, 1)
    assert np.array_equal(classic.fixed, fine.fixed) and np.array_equal(classic.cigar, fine.cigar)


def test_fine_configurations_headline_shape_and_large_match_scores(monkeypatch):
    # 40,000 window / clip pairs through the fine configurations (by default only calls of >= 196,608 pairs take them);
    # match = 5 leaves the tag bits only to the narrower configurations, the others fall back to the classic ones
    # inside the same call
    monkeypatch.setenv("DB200_SW_FINE_MIN_PAIRS", "0")
    q, t, prm = workloads.remap_windows(40000, seed=23)
    res = batch.sw_align_batch(q, t, **prm)
    ref = O.oracle_sw((q.buf, q.off), (t.buf, t.off), prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"], csr=True)
    assert np.array_equal(res.fixed, ref.fixed) and np.array_equal(res.cigar, ref.cigar) and np.array_equal(res.cigar_off, ref.cigar_off)
    qs, ts = pairgen.sw_pairs(515, 3000, 600)
    res = batch.sw_align_batch(qs, ts, 5, -4, 8, 2)
    ref = O.oracle_sw(qs, ts, 5, -4, 8, 2)
    compare(qs, ts, res, ref, "fine match=5")


# ---- second half of round 2: perfect-diagonal shortcut, word-level reversal, anti-diagonal traceback storage ----------

def test_exact_and_nearly_exact_clips_at_every_offset():
    # Exact clips skip the reverse pass (sw_mid_kernel's perfect-diagonal shortcut); clips with one substitution, one N, a
    # lower-case base or an indel must not.  Offsets 0..40 and lengths 1..70 put the alignment's ends on every position of
    # the packed words (the reversed prefixes are written eight codes at a time, the first seven positions one by one).
    import random
    rng = random.Random(77)
    qs, ts = [], []
    for length in list(range(1, 41)) + [47, 48, 49, 56, 63, 64, 65, 70]:
        for off in (0, 1, 2, 3, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33, 40):
            w = pairgen.rand_dna(rng, off + length + rng.randint(0, 50))
            clip = w[off:off + length]
            variants = [clip, clip.lower()]
            if length >= 3:
                mid = length // 2
                variants.append(clip[:mid] + ("A" if clip[mid] != "A" else "C") + clip[mid + 1:])      # one substitution
                variants.append(clip[:mid] + "N" + clip[mid + 1:])                                      # an N pair scores 0
                variants.append(clip[:mid] + clip[mid + 1:])                                            # deletion
                variants.append(clip[:mid] + "G" + clip[mid:])                                          # insertion
            for v in variants:
                qs.append(w.encode()); ts.append(v.encode())
                qs.append(v.encode()); ts.append(w.encode())          # roles swapped: long target, short query
    # tandem repeats: several perfect diagonals end in the same column / row
    for unit_len in (3, 5, 8, 13):
        u = pairgen.rand_dna(rng, unit_len)
        for qrep, trep in ((6, 2), (2, 6), (4, 4), (7, 3)):
            qs.append((pairgen.rand_dna(rng, 5) + u * qrep).encode()); ts.append((u * trep + pairgen.rand_dna(rng, 4)).encode())
    for scoring in ((2, -8, 6, 1), (2, -3, 10, 1), (5, -4, 8, 2), (1, -1, 1, 1)):
        m, x, go, ge = scoring
        res = batch.sw_align_batch(qs, ts, m, x, go, ge)
        ref = O.oracle_sw(qs, ts, m, x, go, ge)
        compare(qs, ts, res, ref, "exact clips " + str(scoring))


def test_diagonal_shortcut_equals_the_reverse_pass(monkeypatch):
    q, t, prm = workloads.remap_windows(6000, seed=41)
    a = batch.sw_align_batch(q, t, **prm)
    monkeypatch.setenv("DB200_SW_NO_DIAG_SHORTCUT", "1")
    b = batch.sw_align_batch(q, t, **prm)
    assert np.array_equal(a.fixed, b.fixed) and np.array_equal(a.cigar, b.cigar) and np.array_equal(a.cigar_off, b.cigar_off)
