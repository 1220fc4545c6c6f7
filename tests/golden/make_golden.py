"""Generate golden vectors from the reference's own compiled code (run where /root/reference exists).

    python tests/golden/make_golden.py

Sources of truth (built by oracle/build_ref.sh, unmodified reference sources):
  * oracle/_ref/libssw_ref.so, libedlib_ref.so  - C cores behind oracle/ref_driver.c
  * oracle/_ref/pyref/dysgu/...                  - the reference's Cython wrappers (wrapper-level vectors)
Outputs (committed): tests/golden/sw_core.json, ed_core.json, wrapper.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pyref"))

import pairgen  # noqa: E402
from oracle import oracle as O  # noqa: E402

SW_SCORINGS = [(2, -8, 6, 1), (2, -3, 10, 1), (2, -3, 5, 1), (2, -3, 4, 1), (2, -3, 5, 2)]


def sw_core():
    out = []
    for si, sc in enumerate(SW_SCORINGS):
        qs, ts = pairgen.sw_pairs(9000 + si, 120, 260)
        for ss, flag in ((2, 1), (1, 1), (2, 0), (2, 8)):
            if (ss, flag) != (2, 1) and si > 1:
                continue
            r = O.ref_sw(qs, ts, *sc, score_size=ss, flag=flag)
            out.append({"scoring": sc, "score_size": ss, "flag": flag,
                        "pairs": [[q.decode(), t.decode()] for q, t in zip(qs, ts)],
                        "results": [list(r.as_tuple(i)) for i in range(len(qs))]})
    return out


def ed_core():
    out = []
    for mi, (mode, task, k) in enumerate([("HW", "locations", None), ("HW", "distance", None), ("NW", "distance", None),
                                          ("NW", "locations", None), ("SHW", "locations", None), ("HW", "locations", 4),
                                          ("NW", "distance", 30)]):
        qs, ts = pairgen.ed_pairs(7000 + mi, 150, 260, alphabets=("ACGT", "AC", "ACGTN", "ACGTacgt"))
        keep = [i for i in range(len(qs)) if max(qs[i] + ts[i] + b"\x00") < 128]
        qs, ts = [qs[i] for i in keep], [ts[i] for i in keep]
        r = O.ref_ed(qs, ts, mode, task, k)
        out.append({"mode": mode, "task": task, "k": k, "pairs": [[q.decode(), t.decode()] for q, t in zip(qs, ts)],
                    "results": [r.as_dict(i) for i in range(len(qs))]})
    return out


def wrapper():
    """Wrapper-level behaviour: attribute views, aligned sequences, index base, edlib dicts incl. SURVEY 8c known answers."""
    from dysgu.scikitbio._ssw_wrapper import StripedSmithWaterman
    from dysgu.edlib.edlib import align
    sw_cases = [
        ("ACGTTGCAACGTACGTTTGACCA", "ttgcaacgtaggtttga", dict(gap_open_penalty=6, gap_extend_penalty=1, match_score=2, mismatch_score=-8)),
        ("ACGTNNNNACGTACGTAC", "acgtnnnnacgtacgtac", {}),
        ("ACGTACGTACGTACGTAC", "ACGTACGTACGTACGTAC", {}),
        ("A" * 20, "C" * 20, {}),
        ("N" * 20, "ACGT" * 5, {}),
        ("ACGTXYZ-ACGTACGTAC", "ACGTACGTACGTACGTAC", {}),
        ("A" * 200, "A" * 200, {}),
        ("ACGTACGTTAGCTAGCTAGGATCGATCGTAGCTAGCTAGCATCGAT", "CGTTAGCTAGCAGGATCGTCGTAGCTAGTAGCAT", dict(gap_open_penalty=5, gap_extend_penalty=1, zero_index=False)),
        ("ACGTACGTTAGCTAGCTAGGATCGATCGTAGCTAGCTAGCATCGAT", "CGTTAGCTAGCAGGATCGTCGTAGCTAGTAGCAT", dict(gap_open_penalty=10, gap_extend_penalty=1, suppress_sequences=True)),
        ("ACGTACGTTAGCTAGCTAGGATCGATCGTAGCTAGCTAGCATCGAT", "CGTTAGCTAGCAGGATCGTCGTAGCTAGTAGCAT", dict(score_only=True)),
        ("ACGTACGTTAGCTAGCTAGGATCGATCGTAGCTAGCTAGCATCGAT", "CGTTAGCTAGCAGGATCGTCGTAGCTAGTAGCAT", dict(mask_auto=False, mask_length=15)),
    ]
    qs, ts = pairgen.sw_pairs(31337, 40, 150)
    for q, t in zip(qs, ts):
        sw_cases.append((q.decode(), t.decode(), dict(gap_open_penalty=6, gap_extend_penalty=1, match_score=2, mismatch_score=-8)))
    sw_out = []
    keys = ["optimal_alignment_score", "suboptimal_alignment_score", "query_begin", "query_end", "target_begin",
            "target_end_optimal", "target_end_suboptimal", "cigar", "query_sequence", "target_sequence",
            "aligned_query_sequence", "aligned_target_sequence"]
    for q, t, kw in sw_cases:
        a = StripedSmithWaterman(q, **kw)(t)
        sw_out.append({"query": q, "target": t, "kwargs": kw, "result": {k: a[k] for k in keys}, "str": str(a)})
    ed_cases = [
        ("ACGTTGCA", "TTACGTAGCATT", "HW", "locations", -1), ("ACGTTGCA", "ACGTAGCA", "NW", "distance", -1),
        ("ACGT", "acgt", "HW", "locations", -1), ("AAAA", "CCCCCCCC", "HW", "locations", -1),
        ("ACGTACGT", "TTTTACGAACGTTTTTACGTACGTTT", "HW", "locations", -1), ("", "ACGT", "HW", "locations", -1),
        ("ACGT", "", "NW", "locations", -1), ("ACGTACGTAA", "ACGTACGT", "NW", "distance", 1), ("telephone", "elephant", "NW", "locations", -1),
    ]
    ed_out = [{"query": q, "target": t, "mode": m, "task": ta, "k": k, "result": align(q, t, m, ta, k)} for q, t, m, ta, k in ed_cases]
    return {"sw": sw_out, "ed": ed_out}


if __name__ == "__main__":
    for name, fn in (("sw_core", sw_core), ("ed_core", ed_core), ("wrapper", wrapper)):
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, separators=(",", ":"))
        print(name, os.path.getsize(os.path.join(HERE, name + ".json")), "bytes")
