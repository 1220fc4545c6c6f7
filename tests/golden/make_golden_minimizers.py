"""tests/golden/minimizers.json: minimizer sketches (graph.pyx:59-84) of a few sequences computed with the REFERENCE's own hash
(dysgu/include/xxhash64.h compiled into oracle/_ref/libxxh_ref.so) and a direct transcription of the window rule for the
generator only.   python tests/golden/make_golden_minimizers.py"""
import json
import os
import random
import sys
from collections import deque

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def ref_hash(b, seed=42):
    v = O.ref_xxh_lib().ref_xxh64(b, len(b), seed)
    return v - (1 << 64) if v >= (1 << 63) else v


def sketch(s, k, m):
    found, win = set(), deque()
    end = len(s) - m + 1
    for i in range(max(end, 0)):
        hx = ref_hash(s[i:i + m])
        if i == 0 or i == end - 1:
            found.add(hx)
        while win and win[-1][0] >= hx:
            win.pop()
        win.append((hx, i))
        while win and win[0][1] <= i - k:
            win.popleft()
        found.add(win[0][0])
    return sorted(found)


def main():
    rng = random.Random(7)
    cases = []
    for L in [0, 1, 6, 7, 8, 15, 22, 23, 24, 60, 150, 300, 1200]:
        s = "".join(rng.choice("ACGT") for _ in range(L))
        cases.append(s)
    cases += ["A" * 80, "ACACACACACACACACACACACACACACACACAC", "acgtNNNNacgtACGTTTGACA" * 4, "".join(rng.choice("ACGTN") for _ in range(500))]
    out = {"hash": [[b.hex(), ref_hash(b)] for b in (b"", b"A", b"ACGTACG", b"ACGTACGT", b"ACGTACGTACGT", bytes(range(40)), bytes(range(77)))],
           "sketches": [{"seq": s, "k": k, "m": m, "values": sketch(s.encode(), k, m)} for s in cases for k, m in ((16, 7), (4, 3), (1, 7), (30, 12))]}
    with open(os.path.join(HERE, "minimizers.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("minimizers.json", os.path.getsize(os.path.join(HERE, "minimizers.json")), "bytes,", len(out["sketches"]), "sketches")


if __name__ == "__main__":
    main()
