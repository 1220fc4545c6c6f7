"""GPU: minimizer sketches (csrc/minimizer.cu) against the golden vectors made with the reference's own hash and
against the scalar oracle on random clip batches (lengths 0 .. 20,000: shared-memory and global-memory sort paths)."""
import json
import os
import random

import numpy as np
import pytest

from dysgu_b200 import kmers
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_sketches_and_host_hash():
    with open(os.path.join(G, "minimizers.json")) as f:
        g = json.load(f)
    for hx, want in g["hash"]:
        v = kmers.xxh64(bytes.fromhex(hx))
        assert (v - (1 << 64) if v >= (1 << 63) else v) == want
    by_param = {}
    for c in g["sketches"]:
        by_param.setdefault((c["k"], c["m"]), []).append(c)
    for (k, m), cases in by_param.items():
        vals, off = kmers.minimizers_batch([c["seq"].encode() for c in cases], k=k, m=m)
        for i, c in enumerate(cases):
            assert vals[off[i]:off[i + 1]].tolist() == c["values"], (c["seq"][:30], k, m)


@pytest.mark.parametrize("seed,maxlen,n", [(1, 300, 20000), (2, 3000, 2000), (3, 20000, 60)])
def test_random_batches_match_the_oracle(seed, maxlen, n):
    rng = random.Random(seed)
    seqs = []
    for _ in range(n):
        L = rng.choice([0, 3, 6, 7, 8, 22, 23]) if rng.random() < 0.05 else rng.randint(9, maxlen)
        kind = rng.random()
        if kind < 0.7:
            s = "".join(rng.choice("ACGT") for _ in range(L))
        elif kind < 0.85:
            unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 6)))
            s = (unit * (L // len(unit) + 1))[:L]           # tandem repeats: many equal hashes, ties in the window
        else:
            s = "".join(rng.choice("ACGTNacgt") for _ in range(L))
        seqs.append(s.encode())
    for k, m in ((16, 7), (5, 11)):
        got_v, got_o = kmers.minimizers_batch(seqs, k=k, m=m)
        ref_v, ref_o = O.oracle_minimizers(seqs, k=k, m=m)
        assert np.array_equal(got_o, ref_o)
        assert np.array_equal(got_v, ref_v)
    sets = kmers.minimizer_sets(seqs[:50])
    assert all(len(s) == got_o_i for s, got_o_i in zip(sets, np.diff(kmers.minimizers_batch(seqs[:50])[1])))
