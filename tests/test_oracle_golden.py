"""CPU: the oracle restatement against the golden vectors generated from the reference's compiled code."""
import json
import os

import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_sw_oracle_matches_golden():
    for block in load("sw_core.json"):
        qs = [p[0].encode() for p in block["pairs"]]
        ts = [p[1].encode() for p in block["pairs"]]
        r = O.oracle_sw(qs, ts, *block["scoring"], score_size=block["score_size"], flag=block["flag"])
        for i, want in enumerate(block["results"]):
            assert list(r.as_tuple(i)) == want, (block["scoring"], block["score_size"], block["flag"], qs[i], ts[i])


def test_ed_oracle_matches_golden():
    for block in load("ed_core.json"):
        qs = [p[0].encode() for p in block["pairs"]]
        ts = [p[1].encode() for p in block["pairs"]]
        r = O.oracle_ed(qs, ts, block["mode"], block["task"], block["k"])
        for i, want in enumerate(block["results"]):
            got = r.as_dict(i)
            want = dict(want)
            want["locations"] = [tuple(x) for x in want["locations"]]
            assert got == want, (block["mode"], block["task"], block["k"], qs[i], ts[i])


# known-answer vectors recorded in SURVEY.md 8c from the compiled reference
SW_KNOWN = [
    (("ACGTTGCAACGTACGTTTGACCA", "ttgcaacgtaggtttga"), (2, -8, 6, 1), (24, 2, 3, 19, 0, 16, 0, "17M")),
    (("ACGTNNNNACGTACGTAC", "acgtnnnnacgtacgtac"), (2, -3, 5, 2), (28, 4, 0, 17, 0, 17, 1, "18M")),
    (("A" * 20, "C" * 20), (2, -3, 5, 2), (0, 0, 0, 0, -1, -1, 0, "1M")),
    (("A" * 200, "A" * 200), (2, -3, 5, 2), (400, 198, 0, 199, 0, 199, 98, "200M")),
]


@pytest.mark.parametrize("pair,scoring,want", SW_KNOWN)
def test_sw_known_answers(pair, scoring, want):
    # SURVEY tuple order: score, score2, q_begin, q_end, t_begin, t_end, t_end_sub, cigar
    r = O.oracle_sw([pair[0].encode()], [pair[1].encode()], *scoring)
    s1, s2, tb, te, qb, qe, te2, cig = r.as_tuple(0)
    assert (s1, s2, qb, qe, tb, te, te2, cig) == want


ED_KNOWN = [
    ("ACGTTGCA", "TTACGTAGCATT", "HW", "locations", None, (1, 4, [(2, 9)])),
    ("ACGTTGCA", "ACGTAGCA", "NW", "distance", None, (1, 4, [(None, 7)])),
    ("ACGT", "acgt", "HW", "locations", None, (4, 8, [(0, -1), (0, 0), (0, 1), (0, 2), (0, 3)])),
    ("AAAA", "CCCCCCCC", "HW", "locations", None, (4, 2, [(0, -1), (0, 0), (0, 1), (0, 2), (0, 3), (1, 4), (2, 5), (3, 6), (4, 7)])),
    ("ACGTACGT", "TTTTACGAACGTTTTTACGTACGTTT", "HW", "locations", None, (0, 4, [(16, 23)])),
    ("", "ACGT", "HW", "locations", None, (0, 4, [(None, -1)])),
    ("ACGTACGTAA", "ACGTACGT", "NW", "distance", 1, (-1, 4, [])),
]


@pytest.mark.parametrize("q,t,mode,task,k,want", ED_KNOWN)
def test_ed_known_answers(q, t, mode, task, k, want):
    d = O.oracle_ed([q.encode()], [t.encode()], mode, task, k).as_dict(0)
    assert (d["editDistance"], d["alphabetLength"], d["locations"]) == want


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_oracle_equals_compiled_reference_random():
    import pairgen
    qs, ts = pairgen.sw_pairs(4711, 1500, 300)
    a = O.oracle_sw(qs, ts, 2, -8, 6, 1)
    b = O.ref_sw(qs, ts, 2, -8, 6, 1)
    assert all(a.as_tuple(i) == b.as_tuple(i) for i in range(len(qs)))
    qs, ts = pairgen.ed_pairs(4712, 1500, 300)
    for mode in ("HW", "NW", "SHW"):
        a = O.oracle_ed(qs, ts, mode, "locations")
        b = O.ref_ed(qs, ts, mode, "locations")
        assert all(a.as_dict(i) == b.as_dict(i) for i in range(len(qs)))


def test_ed_path_oracle_matches_golden():
    """The scalar restatement of task "path" (oracle/ed_path_oracle.c) against the reference wrapper's vectors: direct
    traceback and the halving regime (up to 17 kb x 17 kb), NW / SHW / HW, k cut-offs, additionalEqualities."""
    from dysgu_b200.batch import ops_to_cigar
    checked = none = 0
    for b in load("ed_path.json"):
        qs = [c["query"].encode() for c in b["cases"]]
        ts = [c["target"].encode() for c in b["cases"]]
        eqs = [tuple(e) for e in b.get("equalities", [])]
        r = O.oracle_ed_path(qs, ts, b["mode"], None if b["k"] == -1 else b["k"], eqs)
        for i, c in enumerate(b["cases"]):
            key = (b["mode"], b["k"], i, len(qs[i]), len(ts[i]))
            d = r.as_dict(i)
            assert d["editDistance"] == c["editDistance"] and d["locations"] == [tuple(x) for x in c["locations"]], key
            ops = r.path[r.path_off[i]:r.path_off[i] + r.path_len[i]]
            assert (ops_to_cigar(ops) if len(ops) else None) == c["cigar"], key
            checked += 1
            none += c["cigar"] is None
    assert checked > 1500 and none > 100


def test_minimizer_oracle_against_golden_and_reference_hash():
    """oracle/minimizer_oracle.c (restated XXH64 + the window rule of graph.pyx:59-84) against vectors made with the reference's
    own hash, and - where oracle/_ref/libxxh_ref.so exists - hash by hash on random inputs of every length 0..80."""
    import ctypes as C
    import random
    with open(os.path.join(G, "minimizers.json")) as f:
        g = json.load(f)
    L = O.oracle_lib()
    for hx, want in g["hash"]:
        b = bytes.fromhex(hx)
        v = L.oracle_xxh64(b, len(b), 42)
        assert (v - (1 << 64) if v >= (1 << 63) else v) == want
    for c in g["sketches"]:
        vals, off = O.oracle_minimizers([c["seq"].encode()], k=c["k"], m=c["m"])
        assert vals.tolist() == c["values"], (c["seq"][:30], c["k"], c["m"])
    if os.path.exists(os.path.join(O.REF, "libxxh_ref.so")):
        rng = random.Random(3)
        for n in range(81):
            for _ in range(10):
                b = bytes(rng.randrange(256) for _ in range(n))
                assert L.oracle_xxh64(b, n, 42) == O.ref_xxh_lib().ref_xxh64(b, n, 42)
