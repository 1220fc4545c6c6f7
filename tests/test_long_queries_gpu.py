"""GPU: edit-distance queries above 10,240 symbols - the 192 / 256 / 512 / 1024-word Myers configurations and, for task
"path", the global-scratch instance of the thread kernel (queries above 16,384 symbols).  dysgu reaches this range through
filter_normals.clip_align_matches (filter_normals.py:676-684: read clips of 16 kb and more go to edlib HW); edlib itself
handles any length (edlib.cpp:193-212).  Checker: the scalar oracle (distance, locations, edit operations).
First run on a B200 in round 2 (profiles/r02_pytest_gpu.txt)."""
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


def test_query_of_65536_symbols():
    """The longest query the product accepts (1024 words): NW and HW distance + locations against the scalar oracle."""
    rng = random.Random(65536)
    q = pairgen.rand_dna(rng, 65536)
    t = pairgen.mutate(rng, q[300:], 0.01, 0.01, 0.01)
    t_hw = pairgen.rand_dna(rng, 500) + pairgen.mutate(rng, q, 0.004, 0.004, 0.004) + pairgen.rand_dna(rng, 700)
    qs, ts = [q.encode(), q.encode()], [t.encode(), t_hw.encode()]
    for mode in ("NW", "HW"):
        ref = O.oracle_ed(qs, ts, mode, "locations")
        r = batch.ed_align_batch(qs, ts, mode=mode, task="locations")
        assert np.array_equal(r.fixed, ref.fixed), mode
        assert np.array_equal(r.loc, ref.loc) and np.array_equal(r.loc_off, ref.loc_off), mode


def test_query_above_65536_is_refused_not_wrong():
    """Documented deviation: beyond 65,536 query symbols the status is EDLIB_STATUS_ERROR, never a number."""
    q = b"ACGT" * 16400   # 65,600
    r = batch.ed_align_batch([q], [q], mode="NW", task="distance")
    assert int(r.fixed[0, 0]) == 1
