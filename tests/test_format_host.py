"""CPU: the batch text formatters of the C ABI (db200_sw_cigar_text, db200_sw_aligned_text, db200_ed_cigar_text; host-only
code) against the reference wrapper's own strings (tests/golden/wrapper.json, ed_path.json) and against the
wrapper's algorithm restated in Python (_ssw_wrapper.pyx:240-275, 345-364) on oracle results of random pairs."""
import json
import os

import numpy as np

from dysgu_b200 import batch
from dysgu_b200.workloads import SeqBatch
from oracle import oracle as O

import pairgen

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sw_results(qs, ts, *scoring):
    r = O.oracle_sw(qs, ts, *scoring)
    return batch.SwResults(r.fixed, r.cigar, r.cigar_off, r.status)


def aligned_py(sequence, cigar_ops, begin, end, gap_op):
    """_get_aligned_sequence (_ssw_wrapper.pyx:345-364) on BAM-encoded ops"""
    seq = sequence[begin:end + 1]
    pieces, index = [], 0
    for c in cigar_ops:
        length, op = int(c) >> 4, int(c) & 0xF
        if op == gap_op:
            pieces.append("-" * length)
        else:
            pieces.append(seq[index:index + length])
            index += length
    pieces.append(seq[index:end - begin + 1])
    return "".join(pieces)


def test_wrapper_strings_of_the_reference():
    g = json.load(open(os.path.join(G, "wrapper.json")))
    plain = {"gap_open_penalty", "gap_extend_penalty", "match_score", "mismatch_score"}   # default flag, sequences kept, zero-based
    cases = [c for c in g["sw"] if set(c["kwargs"]) <= plain and "aligned_query_sequence" in c["result"]]
    assert len(cases) >= 15
    by_scoring = {}
    for c in cases:
        k = c["kwargs"]
        key = (k.get("match_score", 2), k.get("mismatch_score", -3), k.get("gap_open_penalty", 5), k.get("gap_extend_penalty", 2))
        by_scoring.setdefault(key, []).append(c)
    for key, cs in by_scoring.items():
        qs = [c["query"].encode() for c in cs]
        ts = [c["target"].encode() for c in cs]
        r = sw_results(qs, ts, *key)
        cig = r.cigar_strings()
        aq, at = r.aligned_strings(qs, ts)
        for i, c in enumerate(cs):
            assert cig[i] == c["result"]["cigar"], (key, c["query"][:30])
            assert aq[i] == c["result"]["aligned_query_sequence"] and at[i] == c["result"]["aligned_target_sequence"], (key, c["query"][:30])


def test_batch_strings_equal_the_wrapper_algorithm():
    qs, ts = pairgen.sw_pairs(4242, 1500, 500)
    r = sw_results(qs, ts, 2, -3, 5, 2)
    cig = r.cigar_strings()
    aq, at = r.aligned_strings(qs, ts)
    assert len(cig) == len(aq) == len(at) == len(qs)
    gaps = 0
    for i in range(len(qs)):
        ops = r.cigar_of(i)
        assert cig[i] == r.cigar_str(i)
        q, t = qs[i].decode("latin-1"), ts[i].decode("latin-1")
        f = r.fixed[i]
        tb = int(f[2]) if f[2] >= 0 else -1
        assert aq[i] == aligned_py(q, ops, int(f[4]), int(f[5]), 2), i      # query: read_begin1 / read_end1, gaps at D
        assert at[i] == aligned_py(t, ops, tb, int(f[3]), 1), i             # target: ref_begin1 / ref_end1, gaps at I
        gaps += "-" in aq[i] or "-" in at[i]
    assert gaps > 100
    # empty batch and empty cigars
    e = batch.SwResults(np.zeros((0, 8), np.int32), np.zeros(0, np.uint32), np.zeros(1, np.int64), np.zeros(0, np.int32))
    assert e.cigar_strings() == [] and e.aligned_strings([], []) == ([], [])
    one = batch.SwResults(np.array([[0, 0, 0, 0, 0, 0, 0, 0]], np.int32), np.zeros(0, np.uint32), np.zeros(2, np.int64), np.zeros(1, np.int32))
    assert one.cigar_strings() == [""] and one.aligned_strings([b"ACGT"], [b"ACGT"]) == (["A"], ["A"])


def test_path_cigars_of_a_batch():
    for b in json.load(open(os.path.join(G, "ed_path.json")))[:3]:
        cs = [c for c in b["cases"] if len(c["query"]) <= 4000]
        r = O.oracle_ed_path([c["query"].encode() for c in cs], [c["target"].encode() for c in cs], b["mode"], None if b["k"] == -1 else b["k"])
        off = np.zeros(len(cs) + 1, np.int64)
        np.cumsum(r.path_len, out=off[1:])
        packed = np.concatenate([r.path[r.path_off[i]:r.path_off[i] + r.path_len[i]] for i in range(len(cs))] + [np.zeros(0, np.uint8)])
        res = batch.EdResults(r.fixed, r.loc, r.loc_off, packed, off)
        got = res.cigar_strings()
        assert got == [c["cigar"] for c in cs]
        std = res.cigar_strings(extended=False)
        assert all((s is None) == (g is None) and (s is None or set(s) <= set("0123456789MID")) for s, g in zip(std, got))
        assert [res.cigar(i) for i in range(len(cs))] == got


def test_formatters_report_bad_input_and_small_buffers():
    import ctypes as C
    from dysgu_b200 import _lib
    L = _lib.lib()
    cig = np.array([(17 << 4) | 0, (2 << 4) | 1, (5 << 4) | 2], np.uint32)
    coff = np.array([0, 3], np.int64)
    off = np.zeros(2, np.int64)
    assert L.db200_sw_cigar_text(cig.ctypes.data, coff.ctypes.data, 1, None, 0, off.ctypes.data) == 0 and off[1] == 7   # "17M2I5D"
    buf = np.zeros(16, np.uint8)
    assert L.db200_sw_cigar_text(cig.ctypes.data, coff.ctypes.data, 1, buf.ctypes.data, 6, off.ctypes.data) == -4          # DB200_ERR_CAPACITY
    assert b"too small" in L.db200_last_error()
    assert L.db200_sw_cigar_text(cig.ctypes.data, coff.ctypes.data, 1, buf.ctypes.data, 7, off.ctypes.data) == 0
    assert buf[:7].tobytes() == b"17M2I5D"
    bad = np.array([(3 << 4) | 7], np.uint32)                                                                              # op 7 is not M/I/D
    assert L.db200_sw_cigar_text(bad.ctypes.data, np.array([0, 1], np.int64).ctypes.data, 1, None, 0, off.ctypes.data) == -2
    ops = np.array([0, 0, 9], np.uint8)
    assert L.db200_ed_cigar_text(ops.ctypes.data, np.array([0, 3], np.int64).ctypes.data, 1, 1, None, 0, off.ctypes.data) == -2
    assert L.db200_ed_cigar_text(ops.ctypes.data, np.array([0, 2], np.int64).ctypes.data, 1, 2, None, 0, off.ctypes.data) == -2   # unknown format
