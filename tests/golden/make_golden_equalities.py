"""Golden vectors for edlib's additionalEqualities, generated from the reference's own compiled wrapper
(oracle/_ref/pyref, built by oracle/build_ref.sh; run where /root/reference exists).

    python tests/golden/make_golden_equalities.py      ->  tests/golden/ed_equalities.json
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pyref"))

import pairgen  # noqa: E402
from dysgu.edlib.edlib import align  # noqa: E402  (the reference)

IUPAC = [("R", "A"), ("R", "G"), ("Y", "C"), ("Y", "T"), ("N", "A"), ("N", "C"), ("N", "G"), ("N", "T")]
SETS = {
    "N wildcard": [("N", "A"), ("N", "C"), ("N", "G"), ("N", "T")],
    "iupac": IUPAC,
    "case fold": [("A", "a"), ("C", "c"), ("G", "g"), ("T", "t")],
    "query-only symbol": [("X", "A")],              # X never occurs in a target: only the row of A changes
    "not transitive": [("A", "C"), ("C", "G")],     # A=C and C=G do not make A=G
}


def main():
    rng = random.Random(31337)
    blocks = []
    for name, eqs in SETS.items():
        extra = "".join(sorted({a for a, _ in eqs} | {b for _, b in eqs}))
        for mode, task, k in (("HW", "locations", -1), ("NW", "distance", -1), ("SHW", "locations", -1), ("HW", "locations", 6)):
            cases = []
            for i in range(40):
                L = rng.choice([1, 5, 17, 63, 64, 65, 130, 200, 330])
                q = pairgen.rand_dna(rng, L)
                t = pairgen.mutate(rng, q, 0.05, 0.02, 0.02) or "A"
                if mode != "NW":
                    t = pairgen.rand_dna(rng, rng.randint(0, 60)) + t + pairgen.rand_dna(rng, rng.randint(0, 60))
                q = pairgen.sprinkle(rng, q, 0.08, extra)
                t = pairgen.sprinkle(rng, t, 0.08, extra)
                r = align(q, t, mode=mode, task=task, k=k, additionalEqualities=eqs)
                cases.append({"query": q, "target": t, "result": {"editDistance": r["editDistance"], "alphabetLength": r["alphabetLength"],
                                                                   "locations": r["locations"]}})
            blocks.append({"name": name, "equalities": eqs, "mode": mode, "task": task, "k": k, "cases": cases})
    with open(os.path.join(HERE, "ed_equalities.json"), "w") as f:
        json.dump(blocks, f, separators=(",", ":"))
    print("wrote", sum(len(b["cases"]) for b in blocks), "cases")


if __name__ == "__main__":
    main()
