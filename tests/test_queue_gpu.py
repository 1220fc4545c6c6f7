"""GPU: the ticket queue (db200_queue_*) and the deferred mode of the drop-in classes against the golden vectors -
every pair of every golden block goes through submit / one flush / fetch, mixed parameter sets in one queue."""
import json
import os

import numpy as np
import pytest

import dysgu_b200
from dysgu_b200 import batch
from dysgu_b200.edlib import align
from dysgu_b200.queue import AlignQueue
from dysgu_b200.scikitbio import StripedSmithWaterman

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def cigar_str(ops):
    return "".join("%d%s" % (c >> 4, "MID"[c & 0xF]) for c in ops)


def test_queue_mixed_parameter_sets_match_golden():
    q = AlignQueue(0)
    want_sw, want_ed = [], []
    for block in load("sw_core.json"):
        m, x, go, ge = block["scoring"]
        for (a, b), want in zip(block["pairs"], block["results"]):
            t = q.submit_sw(a.encode(), b.encode(), m, x, go, ge, score_size=block["score_size"], flag=block["flag"])
            want_sw.append((t, want))
    for block in load("ed_core.json"):
        for (a, b), want in zip(block["pairs"], block["results"]):
            t = q.submit_ed(a.encode(), b.encode(), mode=block["mode"], task=block["task"], k=-1 if block["k"] is None else block["k"])
            want_ed.append((t, want))
    assert q.pending == len(want_sw) + len(want_ed)
    q.flush()
    assert q.pending == 0
    for t, want in want_sw:
        fields, ops, status = q.fetch_sw(t)
        assert status == 0 and [int(v) for v in fields[:7]] + [cigar_str(ops)] == want
    for t, want in want_ed:
        got = q.fetch_ed(t)
        want = dict(want)
        want["locations"] = [tuple(v) for v in want["locations"]]
        assert got == want
    # tickets stay readable after later submissions and flushes
    t2 = q.submit_sw(b"ACGTTGCAACGTACGTTTGACCA", b"ttgcaacgtaggtttga", 2, -8, 6, 1)
    f2, ops2, _ = q.fetch_sw(t2)     # implicit flush of one pair = the legacy call
    assert [int(v) for v in f2[:7]] + [cigar_str(ops2)] == [24, 2, 0, 16, 3, 19, 0, "17M"]   # s_align order: ref begin/end, read begin/end
    fields, ops, _ = q.fetch_sw(want_sw[0][0])
    assert [int(v) for v in fields[:7]] + [cigar_str(ops)] == want_sw[0][1]
    q.reset()
    assert q.pending == 0
    with pytest.raises(dysgu_b200._lib.AlignEngineError):
        q.fetch_sw(t2)
    q.close()


def test_deferred_mode_of_the_drop_in_classes():
    g = load("wrapper.json")
    with dysgu_b200.deferred(0) as q:
        alns = [StripedSmithWaterman(c["query"], **c["kwargs"])(c["target"]) for c in g["sw"]]
        eds = [align(c["query"], c["target"], c["mode"], c["task"], c["k"]) for c in g["ed"]]
        assert q.pending == len(alns) + len(eds)     # nothing has run yet
    assert q.pending == 0                            # the block flushed once on exit
    for a, case in zip(alns, g["sw"]):
        for k, want in case["result"].items():
            assert a[k] == want, (case["query"][:40], case["kwargs"], k)
        assert str(a) == case["str"]
    for d, case in zip(eds, g["ed"]):
        want = dict(case["result"])
        want["locations"] = [tuple(v) for v in want["locations"]]
        assert dict(d) == want
    # outside the block the classes are synchronous again
    a = StripedSmithWaterman("ACGTTGCAACGTACGTTTGACCA", 6, 1, match_score=2, mismatch_score=-8)("ttgcaacgtaggtttga")
    assert type(a).__name__ == "AlignmentStructure" and a.cigar == "17M"


def test_deferred_resolves_on_first_use_inside_the_block():
    with dysgu_b200.deferred(0) as q:
        a = StripedSmithWaterman("ACGTACGTACGTACGTAC")("ACGTACGTACGTACGTAC")
        b = StripedSmithWaterman("A" * 200)("A" * 200)
        assert q.pending == 2
        assert a.optimal_alignment_score == 36       # first use flushes everything pending
        assert q.pending == 0
        assert b.optimal_alignment_score == 400 and b.cigar == "200M"


def test_queue_matches_batch_on_random_pairs():
    from pairgen import sw_pairs
    qs, ts = sw_pairs(77, 400, max_len=700)
    r = batch.sw_align_batch(qs, ts, 2, -3, 10, 1)
    q = AlignQueue(0)
    tickets = [q.submit_sw(a, b, 2, -3, 10, 1) for a, b in zip(qs, ts)]
    for i, t in enumerate(tickets):
        fields, ops, status = q.fetch_sw(t)
        assert status == r.status[i] and np.array_equal(fields, r.fixed[i]) and np.array_equal(ops, r.cigar_of(i))
