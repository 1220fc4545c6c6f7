"""Randomised + adversarial alignment pairs for the parity tests (SURVEY.md section 8c classes).

Pure python/numpy, seeded; returns lists of bytes so the same pairs can be fed to the oracle, the
compiled reference and the CUDA path.
"""
from __future__ import annotations

import random

EDGE_LENS = [1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 150, 255, 256, 257]


def rand_dna(rng: random.Random, n: int, alphabet: str = "ACGT") -> str:
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng: random.Random, s: str, sub: float, ins: float, dele: float, alphabet: str = "ACGT") -> str:
    out = []
    for ch in s:
        r = rng.random()
        if r < dele:
            pass
        elif r < dele + sub:
            out.append(rng.choice(alphabet))
        else:
            out.append(ch)
        while rng.random() < ins:
            out.append(rng.choice(alphabet))
    return "".join(out)


def sprinkle(rng: random.Random, s: str, p: float, chars: str) -> str:
    return "".join(rng.choice(chars) if rng.random() < p else ch for ch in s)


def mixed_case(rng: random.Random, s: str) -> str:
    return "".join(ch.lower() if rng.random() < 0.5 else ch for ch in s)


def sw_pairs(seed: int, n: int, max_len: int = 400):
    """(query, target) pairs for Smith-Waterman parity: every class of section 8c at small sizes."""
    rng = random.Random(seed)
    qs, ts = [], []

    def add(q, t):
        if len(q) == 0 or len(t) == 0:
            return
        qs.append(q.encode())
        ts.append(t.encode())

    while len(qs) < n:
        kind = rng.randrange(14)
        if kind == 0:    # edge lengths, related
            L = rng.choice([x for x in EDGE_LENS if x <= max_len])
            q = rand_dna(rng, L)
            add(q, mutate(rng, q, 0.05, 0.02, 0.02) or "A")
        elif kind == 1:  # unrelated random
            add(rand_dna(rng, rng.randint(1, max_len)), rand_dna(rng, rng.randint(1, max_len)))
        elif kind == 2:  # window / clip (long query, short target), clip lower-case
            q = rand_dna(rng, rng.randint(max(max_len // 2, 10), max(max_len, 10)))
            L = rng.randint(min(9, len(q)), min(150, len(q)))
            s = rng.randint(0, len(q) - L)
            add(q, mutate(rng, q[s:s + L], 0.02, 0.01, 0.01).lower() or "a")
        elif kind == 3:  # clip / window (short query, long target)
            t = rand_dna(rng, rng.randint(max(max_len // 2, 10), max(max_len, 10)))
            L = rng.randint(min(9, len(t)), min(150, len(t)))
            s = rng.randint(0, len(t) - L)
            add(mutate(rng, t[s:s + L], 0.02, 0.01, 0.01) or "A", t)
        elif kind == 4:  # tandem duplication in the target (ties, second best)
            u = rand_dna(rng, rng.randint(5, 60))
            q = rand_dna(rng, rng.randint(0, 40)) + u + rand_dna(rng, rng.randint(0, 40))
            add(q, rand_dna(rng, rng.randint(0, 30)) + u * rng.randint(2, 4) + rand_dna(rng, rng.randint(0, 30)))
        elif kind == 5:  # tandem duplication in the query
            u = rand_dna(rng, rng.randint(5, 60))
            t = rand_dna(rng, rng.randint(0, 40)) + u + rand_dna(rng, rng.randint(0, 40))
            add(rand_dna(rng, rng.randint(0, 30)) + u * rng.randint(2, 4) + rand_dna(rng, rng.randint(0, 30)), t)
        elif kind == 6:  # N's and non-ACGT bytes, mixed case
            q = rand_dna(rng, rng.randint(10, max(max_len, 10)))
            t = mutate(rng, q, 0.05, 0.02, 0.02) or "A"
            add(mixed_case(rng, sprinkle(rng, q, 0.05, "NnXRY-*")), mixed_case(rng, sprinkle(rng, t, 0.05, "NnKM.")))
        elif kind == 7:  # no match at all / all-N
            L = rng.randint(1, 60)
            if rng.random() < 0.5:
                add("A" * L, "C" * rng.randint(1, 60))
            else:
                add("N" * L, rand_dna(rng, rng.randint(1, 60)))
        elif kind == 8:  # score straddling the 8-bit -> 16-bit switch (bias 8: >=247, bias 3: >=252)
            L = rng.randint(118, 132)
            q = rand_dna(rng, L)
            t = list(q)
            for _ in range(rng.randint(0, 2)):
                t[rng.randrange(L)] = "N"
            add(rand_dna(rng, rng.randint(0, 20)) + q + rand_dna(rng, rng.randint(0, 20)), "".join(t))
        elif kind == 9:  # indel-rich (band doubling in the traceback)
            q = rand_dna(rng, rng.randint(min(50, max_len), max(max_len, 10)))
            add(q, mutate(rng, q, 0.03, 0.08, 0.08) or "A")
        elif kind == 10:  # low complexity / homopolymer / dinucleotide
            u = rng.choice(["A", "AC", "AT", "CAG", "AAAT"])
            q = (u * (max_len // len(u)))[:rng.randint(5, max_len)]
            t = mutate(rng, (u * (max_len // len(u)))[:rng.randint(5, max_len)], 0.03, 0.01, 0.01) or "A"
            add(q, t)
        elif kind == 11:  # large deletion / insertion inside an otherwise identical pair
            q = rand_dna(rng, rng.randint(min(80, max(max_len, 45)), max(max_len, 45)))
            a = rng.randint(20, len(q) - 20)
            gap = rng.randint(1, 60)
            if rng.random() < 0.5:
                add(q, q[:a] + q[min(len(q), a + gap):])
            else:
                add(q, q[:a] + rand_dna(rng, gap) + q[a:])
        elif kind == 12:  # identical
            q = rand_dna(rng, rng.randint(1, max_len))
            add(q, q)
        else:            # generic related pair, 10-20 % mixed error
            q = rand_dna(rng, rng.randint(min(20, max_len), max(max_len, 10)))
            e = rng.uniform(0.0, 0.2)
            add(q, mutate(rng, q, e * 0.6, e * 0.2, e * 0.2) or "A")
    return qs[:n], ts[:n]


def ed_pairs(seed: int, n: int, max_len: int = 400, alphabets=("ACGT", "AC", "ACGTN", "ACGTacgt")):
    """(query, target) pairs for edit-distance parity, including empty sequences."""
    rng = random.Random(seed)
    qs, ts = [], []

    def add(q, t):
        qs.append(q.encode("latin-1"))
        ts.append(t.encode("latin-1"))

    while len(qs) < n:
        kind = rng.randrange(12)
        ab = rng.choice(alphabets)
        if kind == 0:    # edge lengths around the 64-bit word size (W = 0 and W = 63)
            L = rng.choice([1, 2, 63, 64, 65, 127, 128, 129, 191, 192, 193, 256, 320])
            q = rand_dna(rng, L, ab)
            t = rand_dna(rng, rng.randint(0, 30), ab) + mutate(rng, q, 0.03, 0.02, 0.02, ab) + rand_dna(rng, rng.randint(0, 30), ab)
            add(q, t)
        elif kind == 1:  # unrelated
            add(rand_dna(rng, rng.randint(0, max_len), ab), rand_dna(rng, rng.randint(0, max_len), ab))
        elif kind == 2:  # clip in window
            t = rand_dna(rng, rng.randint(max_len // 2, max_len), ab)
            L = rng.randint(9, min(300, len(t)))
            s = rng.randint(0, len(t) - L)
            add(mutate(rng, t[s:s + L], 0.02, 0.01, 0.01, ab), t)
        elif kind == 3:  # distance == len(query): disjoint alphabets
            add("A" * rng.randint(1, 130), "C" * rng.randint(0, 200))
        elif kind == 4:  # many equal end locations (tandem repeats)
            u = rand_dna(rng, rng.randint(1, 12), ab)
            q = (u * 40)[:rng.randint(3, 100)]
            t = rand_dna(rng, rng.randint(0, 20), ab) + (u * 80)[:rng.randint(10, max_len)] + rand_dna(rng, rng.randint(0, 20), ab)
            add(q, t)
        elif kind == 5:  # query longer than target
            t = rand_dna(rng, rng.randint(1, max_len // 3), ab)
            add(rand_dna(rng, rng.randint(0, 50), ab) + t + rand_dna(rng, rng.randint(0, 50), ab), t)
        elif kind == 6:  # high error, long: forces k-doubling past 64 / 128 in the reference
            q = rand_dna(rng, rng.randint(max_len // 2, max_len), ab)
            add(q, mutate(rng, q, 0.25, 0.1, 0.1, ab))
        elif kind == 7:  # case-sensitive mismatch
            q = rand_dna(rng, rng.randint(1, 80), "ACGT")
            add(q, q.lower() if rng.random() < 0.5 else mixed_case(rng, q))
        elif kind == 8:  # identical
            q = rand_dna(rng, rng.randint(0, max_len), ab)
            add(q, q)
        elif kind == 9:  # arbitrary bytes (wide alphabet)
            q = "".join(chr(rng.randrange(1, 256)) for _ in range(rng.randint(1, 150)))
            t = list(q)
            for _ in range(rng.randint(0, 10)):
                t[rng.randrange(len(t))] = chr(rng.randrange(1, 256))
            add(q, "".join(t))
        else:            # generic related
            q = rand_dna(rng, rng.randint(1, max_len), ab)
            e = rng.uniform(0.0, 0.15)
            add(q, mutate(rng, q, e * 0.5, e * 0.25, e * 0.25, ab))
    return qs[:n], ts[:n]
