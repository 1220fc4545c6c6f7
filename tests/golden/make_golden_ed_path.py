"""Golden vectors for edlib task "path" (cigar in extended format), generated from the reference's own compiled wrapper
(oracle/_ref/pyref; run where /root/reference exists).

    python tests/golden/make_golden_ed_path.py      ->  tests/golden/ed_path.json

Covers additionalEqualities (N wildcards) and the reference's two regimes: direct traceback (state below 1 MB) and the recursive halving above it
(long x long pairs), modes NW / SHW / HW, indel-rich, unrelated and repeat-rich pairs, k cut-offs.
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


def cases(rng):
    out = []
    for i in range(260):                      # small: direct traceback
        kind = i % 6
        L = rng.choice([1, 2, 7, 30, 63, 64, 65, 100, 150, 300, 500])
        q = pairgen.rand_dna(rng, L, "AC" if kind == 5 else "ACGT")
        if kind == 0:
            t = pairgen.mutate(rng, q, 0.02, 0.01, 0.01)
        elif kind == 1:
            t = pairgen.mutate(rng, q, 0.1, 0.08, 0.08)
        elif kind == 2:
            t = pairgen.rand_dna(rng, rng.randint(1, 400))
        elif kind == 3:
            t = q[: L // 2] + pairgen.rand_dna(rng, rng.randint(1, 40)) + q[L // 2:]
        elif kind == 4:
            t = q[: L // 3] + q[L // 3 + min(20, L // 3):]
        else:
            t = pairgen.mutate(rng, q, 0.05, 0.05, 0.05, "AC")
        t = t or "A"
        out.append((q, t))
    for i in range(8):                        # large: the halving regime (state >= 1 MB); the last one is beyond the warp kernel's 256 words
        L = [1800, 2500, 4000, 7000, 10000, 3000, 3500, 17000][i]
        q = pairgen.rand_dna(rng, L)
        if i % 7 == 5:
            t = pairgen.rand_dna(rng, int(L * rng.uniform(0.8, 1.2)))
        elif i % 7 == 6:
            q = pairgen.rand_dna(rng, L, "AC")
            t = pairgen.mutate(rng, q, 0.03, 0.03, 0.03, "AC")
        else:
            e = rng.choice([0.003, 0.02, 0.06])
            t = pairgen.mutate(rng, q, e, e, e)
        out.append((q, t))
    return out


def main():
    rng = random.Random(777)
    blocks = []
    pairs = cases(rng)
    for mode, k in (("NW", -1), ("HW", -1), ("SHW", -1), ("NW", 25), ("HW", 3)):
        cs = []
        for q, t in pairs:
            if len(q) >= 1800 and (mode, k) not in (("NW", -1), ("HW", -1)):
                continue                      # the long pairs only in two blocks (file size)
            if mode != "NW" and len(q) < 5000:   # embed the query's relative in a longer target for the infix / prefix modes
                t2 = pairgen.rand_dna(rng, rng.randint(0, 80)) + t + pairgen.rand_dna(rng, rng.randint(0, 80))
            else:
                t2 = t
            r = align(q, t2, mode=mode, task="path", k=k)
            cs.append({"query": q, "target": t2, "editDistance": r["editDistance"], "locations": r["locations"], "cigar": r["cigar"]})
        blocks.append({"mode": mode, "k": k, "cases": cs})
    # additionalEqualities change which cells match, hence the path: N as a wildcard on either side
    eqs = [["N", "A"], ["N", "C"], ["N", "G"], ["N", "T"]]
    for mode in ("NW", "HW"):
        cs = []
        for q, t in pairs[:120]:
            q2 = "".join("N" if rng.random() < 0.08 else c for c in q)
            t2 = "".join("N" if rng.random() < 0.08 else c for c in t)
            if mode == "HW":
                t2 = pairgen.rand_dna(rng, rng.randint(0, 60)) + t2 + pairgen.rand_dna(rng, rng.randint(0, 60))
            r = align(q2, t2, mode=mode, task="path", k=-1, additionalEqualities=[tuple(e) for e in eqs])
            cs.append({"query": q2, "target": t2, "editDistance": r["editDistance"], "locations": r["locations"], "cigar": r["cigar"]})
        blocks.append({"mode": mode, "k": -1, "equalities": eqs, "cases": cs})
    with open(os.path.join(HERE, "ed_path.json"), "w") as f:
        json.dump(blocks, f, separators=(",", ":"))
    print("wrote", sum(len(b["cases"]) for b in blocks), "cases")


if __name__ == "__main__":
    main()
