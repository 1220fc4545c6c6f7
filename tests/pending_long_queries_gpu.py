"""NOT COLLECTED (the file name does not match test_*.py): the GPU test to enable first next round - rename it to
test_long_queries_gpu.py once it has passed on a B200.

Edit-distance queries above 10,240 symbols run Myers configurations (192, 256, 512, 1024 words per pair) that no GPU test
has exercised yet, and for task "path" queries above 16,384 symbols take the global-scratch instance of the thread
kernel.  Checker: the scalar oracle (distance, locations, edit operations).
"""
import random

import numpy as np
import pytest

from dysgu_b200 import batch
from oracle import oracle as O

import pairgen

pytestmark = pytest.mark.gpu


def pairs(seed):
    rng = random.Random(seed)
    qs, ts = [], []
    for L in (10241, 11000, 12288, 12289, 15000, 16384, 16385, 17000, 24000, 33000):
        q = pairgen.rand_dna(rng, L)
        e = rng.choice([0.002, 0.02, 0.06])
        t = pairgen.mutate(rng, q, e, e, e)
        if L % 2:
            t = pairgen.rand_dna(rng, rng.randint(0, 300)) + t + pairgen.rand_dna(rng, rng.randint(0, 300))
        qs.append(q.encode()); ts.append(t.encode())
    qs.append(pairgen.rand_dna(rng, 12000).encode()); ts.append(pairgen.rand_dna(rng, 9000).encode())   # unrelated
    return qs, ts


@pytest.mark.parametrize("mode", ["NW", "HW", "SHW"])
def test_long_queries_all_tasks(mode):
    qs, ts = pairs(len(mode))
    ref = O.oracle_ed_path(qs, ts, mode)
    for task in ("distance", "locations", "path"):
        r = batch.ed_align_batch(qs, ts, mode=mode, task=task)
        assert np.array_equal(r.fixed, ref.fixed), (mode, task)
        if task != "distance" or mode == "NW":
            want = ref.loc if task != "distance" else None
            if want is not None:
                assert np.array_equal(r.loc, want) and np.array_equal(r.loc_off, ref.loc_off), (mode, task)
        if task == "path":
            for i in range(len(qs)):
                ops = ref.path[ref.path_off[i]:ref.path_off[i] + ref.path_len[i]]
                got = r.alignment(i)
                assert np.array_equal(got if got is not None else np.zeros(0, np.uint8), ops), (mode, i, len(qs[i]))
