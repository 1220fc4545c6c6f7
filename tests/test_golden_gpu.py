"""GPU: the CUDA path against the committed golden vectors (generated from the reference's compiled code), incl. the
wrapper-level view (attributes, aligned sequences, index base) and multi-GPU sharding when several devices exist."""
import json
import os

import numpy as np
import pytest

from dysgu_b200 import _lib, batch, shard, workloads
from dysgu_b200.edlib import align
from dysgu_b200.scikitbio import StripedSmithWaterman

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_sw_core_golden():
    for block in load("sw_core.json"):
        qs = [p[0].encode() for p in block["pairs"]]
        ts = [p[1].encode() for p in block["pairs"]]
        m, x, go, ge = block["scoring"]
        r = batch.sw_align_batch(qs, ts, m, x, go, ge, score_size=block["score_size"], flag=block["flag"])
        for i, want in enumerate(block["results"]):
            assert r.status[i] == 0 and list(r.as_tuple(i)) == want, (block["scoring"], block["score_size"], block["flag"], qs[i], ts[i])


def test_ed_core_golden():
    for block in load("ed_core.json"):
        qs = [p[0].encode() for p in block["pairs"]]
        ts = [p[1].encode() for p in block["pairs"]]
        r = batch.ed_align_batch(qs, ts, mode=block["mode"], task=block["task"], k=-1 if block["k"] is None else block["k"])
        for i, want in enumerate(block["results"]):
            want = dict(want)
            want["locations"] = [tuple(x) for x in want["locations"]]
            assert r.as_dict(i) == want, (block["mode"], block["task"], block["k"], qs[i], ts[i])


def test_wrapper_level_golden():
    g = load("wrapper.json")
    for case in g["sw"]:
        a = StripedSmithWaterman(case["query"], **case["kwargs"])(case["target"])
        for k, want in case["result"].items():
            assert a[k] == want, (case["query"][:40], case["kwargs"], k)
        assert str(a) == case["str"]
    for case in g["ed"]:
        got = align(case["query"], case["target"], case["mode"], case["task"], case["k"])
        want = dict(case["result"])
        want["locations"] = [tuple(x) for x in want["locations"]]
        assert got == want, case


def test_one_profile_many_targets():
    q = "ACGTACGTTAGCTAGCTAGGATCGATCGTAGCTAGCTAGCATCGAT"
    s = StripedSmithWaterman(q, gap_extend_penalty=1, suppress_sequences=True)
    clip = "CGTTAGCTAGCAGGATCG"
    rc = clip[::-1].translate(str.maketrans("ACGT", "TGCA"))
    many = s.align_many([clip, rc])
    assert [a.optimal_alignment_score for a in many] == [s(clip).optimal_alignment_score, s(rc).optimal_alignment_score]


def test_legacy_c_entry_points():
    import ctypes as C
    L = _lib.lib()

    class SAlign(C.Structure):
        _fields_ = [("score1", C.c_uint16), ("score2", C.c_uint16), ("ref_begin1", C.c_int32), ("ref_end1", C.c_int32),
                    ("read_begin1", C.c_int32), ("read_end1", C.c_int32), ("ref_end2", C.c_int32), ("cigar", C.POINTER(C.c_uint32)),
                    ("cigarLen", C.c_int32)]
    L.db200_ssw_init.restype = C.c_void_p
    L.db200_ssw_align.restype = C.POINTER(SAlign)
    L.db200_ssw_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint16, C.c_int32, C.c_int32]
    enc = {"A": 0, "C": 1, "G": 2, "T": 3}
    q = np.array([enc[c] for c in "ACGTTGCAACGTACGTTTGACCA"], dtype=np.int8)
    t = np.array([enc[c] for c in "TTGCAACGTAGGTTTGA"], dtype=np.int8)
    mat = np.array([[0 if 4 in (a, b) else (2 if a == b else -8) for b in range(5)] for a in range(5)], dtype=np.int8)
    prof = L.db200_ssw_init(q.ctypes.data, len(q), mat.ctypes.data, 5, 2)
    a = L.db200_ssw_align(prof, t.ctypes.data, len(t), 6, 1, 1, 0, 0, 15).contents
    assert (a.score1, a.ref_begin1, a.ref_end1, a.read_begin1, a.read_end1, a.cigarLen, a.cigar[0]) == (24, 0, 16, 3, 19, 1, 17 << 4)


def test_full_size_properties():
    # size-independent properties at a larger batch: cigars consume exactly the reported spans, scores are symmetric under
    # swapping query and target for a symmetric matrix, identical pairs score match * length
    q, t, prm = workloads.remap_windows(60000, seed=123)
    r = batch.sw_align_batch(q, t, **prm)
    assert (r.status == 0).all()
    lens = r.cigar_len
    ops = r.cigar & 0xF
    ln = (r.cigar >> 4).astype(np.int64)
    idx = np.repeat(np.arange(len(q)), lens)
    qspan = np.bincount(idx, weights=ln * (ops != 2), minlength=len(q))
    tspan = np.bincount(idx, weights=ln * (ops != 1), minlength=len(q))
    pos = r.score1 > 0
    assert np.array_equal(qspan[pos], (r.read_end1 - r.read_begin1 + 1)[pos])
    assert np.array_equal(tspan[pos], (r.ref_end1 - r.ref_begin1 + 1)[pos])
    sw = batch.sw_align_batch(t, q, flag=0, **prm)
    assert np.array_equal(sw.score1, r.score1)
    same = batch.sw_align_batch(t, t, flag=0, **prm)
    assert np.array_equal(same.score1, 2 * t.lengths)


@pytest.mark.skipif(_lib.device_count() < 2, reason="needs 2 GPUs")
def test_multi_gpu_threads_match_single_gpu():
    q, t, prm = workloads.remap_windows(20000, seed=5)
    one = batch.sw_align_batch(q, t, **prm)
    parts = shard.align_multi_gpu(batch.sw_align_batch, q, t, devices=list(range(min(_lib.device_count(), 8))), **prm)
    fixed = np.concatenate([p.fixed for p in parts], axis=0)
    assert np.array_equal(fixed, one.fixed)
    off, cig = shard.merge_csr([p.cigar_off for p in parts], [p.cigar for p in parts])
    assert np.array_equal(off, one.cigar_off) and np.array_equal(cig, one.cigar)


def test_multi_device_entry_points_balance_by_work_and_keep_pair_order():
    """db200_sw_batch_host_multi / db200_ed_batch_host_multi: ranges cut at equal work, one host thread per device, outputs in
    pair order.  With one GPU leased the device list names it several times - the ranges then run one after the other
    (per-device lock) through exactly the same split / splice code; with several GPUs they run side by side."""
    ndev = _lib.device_count()
    devices = list(range(min(ndev, 8))) if ndev >= 2 else [0, 0, 0]
    # lengths sorted ascending: equal pair counts would give the last device most of the work
    q, t, prm = workloads.pe_contig_pairs(6000, seed=3)
    order = np.argsort(q.lengths * t.lengths, kind="stable")
    q, t = q.take(order), t.take(order)
    one = batch.sw_align_batch(q, t, **prm)
    many = batch.sw_align_batch(q, t, devices=devices, **prm)
    assert np.array_equal(many.fixed, one.fixed) and np.array_equal(many.status, one.status)
    assert np.array_equal(many.cigar_off, one.cigar_off) and np.array_equal(many.cigar, one.cigar)
    per = many.pairs_per_device
    assert per.sum() == len(q) and per[0] > per[-1] * 1.5          # short pairs first: the first range holds many more of them
    cells = (q.lengths * t.lengths).astype(np.float64)
    b = np.concatenate(([0], np.cumsum(per)))
    work = np.array([cells[b[i]:b[i + 1]].sum() for i in range(len(per))])
    assert work.max() / work.mean() < 1.15, work
    c, w, eprm = workloads.remap_chain(9000, seed=4)
    e1 = batch.ed_align_batch(c, w, **eprm)
    e2 = batch.ed_align_batch(c, w, devices=devices, **eprm)
    assert np.array_equal(e1.fixed, e2.fixed) and np.array_equal(e1.loc, e2.loc) and np.array_equal(e1.loc_off, e2.loc_off)
    small = batch.sw_align_batch(q.take(np.arange(3)), t.take(np.arange(3)), devices=devices, **prm)   # fewer pairs than 2 x devices
    assert np.array_equal(small.fixed, one.fixed[:3])


def test_additional_equalities_golden():
    """edlib's additionalEqualities (symmetric, not transitive; edlib.cpp:60-91) against vectors produced by the
    reference's own wrapper (tests/golden/make_golden_equalities.py): wildcards, IUPAC pairs, case folding, a symbol that
    only occurs in queries, a non-transitive chain - through the batch call and through edlib.align."""
    for block in load("ed_equalities.json"):
        eqs = [tuple(e) for e in block["equalities"]]
        qs = [c["query"].encode() for c in block["cases"]]
        ts = [c["target"].encode() for c in block["cases"]]
        r = batch.ed_align_batch(qs, ts, mode=block["mode"], task=block["task"], k=block["k"], additional_equalities=eqs)
        for i, c in enumerate(block["cases"]):
            want = dict(c["result"], status=0)
            want["locations"] = [tuple(x) for x in want["locations"]]
            assert r.as_dict(i) == want, (block["name"], block["mode"], block["k"], c["query"][:40], c["target"][:40])
        c = block["cases"][0]
        got = align(c["query"], c["target"], block["mode"], block["task"], block["k"], additionalEqualities=eqs)
        assert got["editDistance"] == c["result"]["editDistance"] and got["locations"] == [tuple(x) for x in c["result"]["locations"]]
