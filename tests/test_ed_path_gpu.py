"""GPU: edlib task "path" through the C ABI (db200_ed_path_batch_host, db200_edlibAlign, the ticket queue) against
vectors from the reference's compiled wrapper (tests/golden/ed_path.json: direct traceback and the halving regime,
NW / SHW / HW, k cut-offs, additionalEqualities), plus size-independent properties on a large random batch."""
import ctypes as C
import json
import os
import random

import numpy as np
import pytest

import dysgu_b200
from dysgu_b200 import _lib, batch
from dysgu_b200.edlib import align, getNiceAlignment

import pairgen

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def blocks():
    """The vectors the device suite runs: all of them, including the 17 kb pair that takes the global-scratch instance of
    the thread kernel (beyond the warp kernel's 256 query words)."""
    with open(os.path.join(G, "ed_path.json")) as f:
        return json.load(f)


def test_path_batches_match_the_reference():
    checked = none = big = 0
    for b in blocks():
        eqs = [tuple(e) for e in b.get("equalities", [])] or None
        qs = [c["query"].encode() for c in b["cases"]]
        ts = [c["target"].encode() for c in b["cases"]]
        r = batch.ed_align_batch(qs, ts, mode=b["mode"], task="path", k=b["k"], additional_equalities=eqs)
        for i, c in enumerate(b["cases"]):
            d = r.as_dict(i)
            key = (b["mode"], b["k"], i, len(qs[i]), len(ts[i]))
            assert d["editDistance"] == c["editDistance"], key
            assert d["locations"] == [tuple(x) for x in c["locations"]], key
            assert d["cigar"] == c["cigar"], key
            checked += 1
            none += c["cigar"] is None
            big += len(qs[i]) >= 1800
    assert checked > 1500 and none > 100 and big >= 10


def test_wrapper_level_path():
    # SURVEY 8c vectors
    got = align("ACGTTGCA", "TTACGTAGCATT", "HW", "path")
    assert got == {"editDistance": 1, "alphabetLength": 4, "locations": [(2, 9)], "cigar": "4=1X3="}
    assert align("ACGTACGTAA", "ACGTACGT", "NW", "path", k=1) == {"editDistance": -1, "alphabetLength": 4, "locations": [], "cigar": None}
    # an empty sequence returns before the path step (edlib.cpp:161-179)
    assert align("", "ACGT", "HW", "path")["cigar"] is None
    assert align("ACGT", "", "NW", "path")["cigar"] is None
    # the -1 end location: the query against an empty slice = insertions only
    got = align("ACGT", "acgt", "HW", "path")
    assert got["locations"][0] == (0, -1) and got["cigar"] == "4I"
    nice = getNiceAlignment(align("ACGTTGCA", "TTACGTAGCATT", "HW", "path"), "ACGTTGCA", "TTACGTAGCATT")
    assert nice == {"query_aligned": "ACGTTGCA", "matched_aligned": "||||.|||", "target_aligned": "ACGTAGCA"}
    # deferred mode: submit, resolve on first use
    c = blocks()[1]["cases"][:40]
    with dysgu_b200.deferred():
        lazy = [align(x["query"], x["target"], "HW", "path") for x in c]
    assert [l["cigar"] for l in lazy] == [x["cigar"] for x in c]


class EdlibAlignConfig(C.Structure):
    _fields_ = [("k", C.c_int), ("mode", C.c_int), ("task", C.c_int), ("additionalEqualities", C.c_void_p),
                ("additionalEqualitiesLength", C.c_int)]


class EdlibAlignResult(C.Structure):
    _fields_ = [("status", C.c_int), ("editDistance", C.c_int), ("endLocations", C.POINTER(C.c_int)),
                ("startLocations", C.POINTER(C.c_int)), ("numLocations", C.c_int), ("alignment", C.POINTER(C.c_ubyte)),
                ("alignmentLength", C.c_int), ("alphabetLength", C.c_int)]


def test_legacy_edlib_align_fills_the_alignment():
    L = _lib.lib()
    L.db200_edlibAlign.restype = EdlibAlignResult
    L.db200_edlibAlign.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, EdlibAlignConfig]
    L.db200_edlibFreeAlignResult.argtypes = [EdlibAlignResult]
    for b in blocks()[:2]:
        for c in b["cases"][:25] + b["cases"][-2:]:
            cfg = EdlibAlignConfig(b["k"], batch.ED_MODES[b["mode"]], batch.ED_TASKS["path"], None, 0)
            q, t = c["query"].encode(), c["target"].encode()
            r = L.db200_edlibAlign(q, len(q), t, len(t), cfg)
            assert r.status == 0 and r.editDistance == c["editDistance"]
            ops = bytes(r.alignment[:r.alignmentLength])
            assert batch.ops_to_cigar(np.frombuffer(ops, np.uint8)) == c["cigar"]
            assert (r.startLocations[0], r.endLocations[0]) == tuple(c["locations"][0])
            L.db200_edlibFreeAlignResult(r)


def test_path_properties_on_a_large_batch():
    """Whatever the size: the operations consume the whole query and exactly target[start..end], '=' columns are equal
    symbols, 'X' columns differ, and the non-match operations number editDistance; the distance and locations equal the
    task 'locations' run of the same pairs."""
    rng = random.Random(99)
    qs, ts = [], []
    for i in range(12000):
        L = rng.choice([20, 60, 100, 150, 151, 250, 300]) if i % 50 else rng.choice([1200, 2500, 6000])
        q = pairgen.rand_dna(rng, L)
        e = rng.choice([0.0, 0.02, 0.08, 0.25])
        t = pairgen.mutate(rng, q, e, e, e) or "A"
        if i % 3:
            t = pairgen.rand_dna(rng, rng.randint(0, 200)) + t + pairgen.rand_dna(rng, rng.randint(0, 200))
        qs.append(q.encode()); ts.append(t.encode())
    r = batch.ed_align_batch(qs, ts, mode="HW", task="path")
    ref = batch.ed_align_batch(qs, ts, mode="HW", task="locations")
    assert np.array_equal(r.fixed, ref.fixed) and np.array_equal(r.loc, ref.loc) and np.array_equal(r.loc_off, ref.loc_off)
    for i in range(len(qs)):
        ops = r.alignment(i)
        start, end = r.loc[r.loc_off[i]]
        q = np.frombuffer(qs[i], np.uint8)
        sl = np.frombuffer(ts[i], np.uint8)[start:end + 1]
        assert ops is not None and int((ops != 0).sum()) == r.edit_distance[i]
        qi = np.cumsum(ops != 2) - 1          # query index consumed by each column (I, =, X)
        ti = np.cumsum(ops != 1) - 1
        assert int((ops != 2).sum()) == len(q) and int((ops != 1).sum()) == len(sl)
        eq, ne = ops == 0, ops == 3
        assert np.array_equal(q[qi[eq]], sl[ti[eq]]) and not np.any(q[qi[ne]] == sl[ti[ne]])


def test_small_path_buffer_is_refused():
    """Operations are packed back to back (CSR); a caller buffer they do not fit is an error, never a truncated path."""
    L = _lib.lib()
    q = np.frombuffer(b"ACGTACGTAC" * 2, np.uint8).copy()
    t = np.frombuffer(b"ACGTTCGTAC" * 2, np.uint8).copy()
    off = np.array([0, 10, 20], np.int64)
    prm = _lib.EdParams()
    prm.mode, prm.task, prm.k = batch.ED_MODES["NW"], batch.ED_TASKS["path"], -1
    fixed = np.zeros((2, 4), np.int32); loc = np.zeros((16, 2), np.int32); loff = np.zeros(3, np.int64)
    poff = np.zeros(3, np.int64)
    for cap, ok in ((20, True), (19, False)):
        path = np.zeros(64, np.uint8)
        rc = L.db200_ed_path_batch_host(0, C.byref(prm), q.ctypes.data, off.ctypes.data, t.ctypes.data, off.ctypes.data, 2, None,
                                        fixed.ctypes.data, loc.ctypes.data, 16, loff.ctypes.data, path.ctypes.data, cap, poff.ctypes.data)
        assert (rc == 0) == ok
        if ok:
            assert list(poff) == [0, 10, 20] and batch.ops_to_cigar(path[:10]) == "4=1X5="
        else:
            assert rc == -4 and b"too small" in L.db200_last_error()
