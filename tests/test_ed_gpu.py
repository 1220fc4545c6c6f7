"""GPU parity of the edit-distance path against the CPU oracle (bit-exact: distance, alphabet, all locations)."""
import pytest

import pairgen
from oracle import oracle as O
from dysgu_b200 import batch, workloads

pytestmark = pytest.mark.gpu


def compare(qs, ts, res, ref, tag, limit=5):
    bad = []
    for i in range(len(qs)):
        a, b = res.as_dict(i), ref.as_dict(i)
        if a != b:
            bad.append((i, a, b))
    msg = "\n".join("%s pair %d q=%r t=%r\n   gpu=%r\n   ref=%r" % (tag, i, qs[i][:60], ts[i][:60], a, b) for i, a, b in bad[:limit])
    assert not bad, "%d/%d differ\n%s" % (len(bad), len(qs), msg)


@pytest.mark.parametrize("mode", ["HW", "NW", "SHW"])
@pytest.mark.parametrize("task", ["distance", "locations"])
def test_random_classes(mode, task):
    qs, ts = pairgen.ed_pairs(300 + len(mode) * 7 + len(task), 2500, 420)
    res = batch.ed_align_batch(qs, ts, mode=mode, task=task)
    ref = O.oracle_ed(qs, ts, mode, task)
    compare(qs, ts, res, ref, "%s/%s" % (mode, task))


@pytest.mark.parametrize("k", [0, 3, 25, 100])
def test_k_threshold(k):
    qs, ts = pairgen.ed_pairs(900 + k, 1500, 300)
    for mode in ("HW", "NW", "SHW"):
        res = batch.ed_align_batch(qs, ts, mode=mode, task="locations", k=k)
        ref = O.oracle_ed(qs, ts, mode, "locations", k)
        compare(qs, ts, res, ref, "%s k=%d" % (mode, k))


def test_long_queries_multiword():
    a, b, _ = workloads.insertion_linking(40, seed=9, platform="ont", lo=500, hi=9000)
    qs, ts = a.tolist(), b.tolist()
    for mode, task in (("NW", "distance"), ("HW", "locations")):
        res = batch.ed_align_batch(qs, ts, mode=mode, task=task)
        ref = O.oracle_ed(qs, ts, mode, task)
        compare(qs, ts, res, ref, "long %s" % mode)


def test_remap_chain_shape():
    q, t, prm = workloads.remap_chain(3000, seed=5)
    res = batch.ed_align_batch(q, t, **prm)
    ref = O.oracle_ed((q.buf, q.off), (t.buf, t.off), prm["mode"], prm["task"], csr=True)
    compare(q.tolist(), t.tolist(), res, ref, "remap chain")


def test_additional_equalities_match_reference():
    # the oracle batch helper has no equality pairs; check against the compiled reference semantics by hand:
    # 'N' acts as a wildcard for A/C/G/T
    eq = [("N", "A"), ("N", "C"), ("N", "G"), ("N", "T")]
    res = batch.ed_align_batch([b"ACNT", b"NNNN"], [b"TTACGTTT", b"ACGT"], mode="HW", task="locations", additional_equalities=eq)
    assert res.as_dict(0)["editDistance"] == 0 and res.as_dict(0)["locations"] == [(2, 5)]
    assert res.as_dict(1)["editDistance"] == 0


def _band_pairs(seed, n, lo, hi):
    """NW pairs that exercise every banded round: error rates from 0.2 % to unrelated, length differences that
    start a pair in a later round, queries of exactly 2/3/5/9/17 words and one word more."""
    import random
    rng = random.Random(seed)
    qs, ts = [], []
    for i in range(n):
        kind = i % 8
        if kind == 6:      # word-count edges of the rounds
            L = rng.choice([64 * w + d for w in (2, 3, 5, 9, 17) for d in (-1, 0, 1)])
        else:
            L = rng.randint(lo, hi)
        q = pairgen.rand_dna(rng, L)
        if kind == 0:
            t = pairgen.mutate(rng, q, 0.002, 0.003, 0.003)
        elif kind == 1:
            t = pairgen.mutate(rng, q, 0.015, 0.015, 0.02)
        elif kind == 2:
            t = pairgen.mutate(rng, q, 0.06, 0.04, 0.04)
        elif kind == 3:    # unrelated, similar length: every round fails, the full-column kernel answers
            t = pairgen.rand_dna(rng, max(1, int(L * rng.uniform(0.7, 1.3))))
        elif kind == 4:    # a long insertion: the length difference alone rules out the narrow rounds
            cut = rng.randint(0, L)
            t = q[:cut] + pairgen.rand_dna(rng, rng.choice([70, 130, 260, 520, 1100])) + q[cut:]
        elif kind == 5:    # a long deletion
            d = min(L - 1, rng.choice([70, 130, 260, 520, 1100]))
            cut = rng.randint(0, L - d)
            t = q[:cut] + q[cut + d:]
        elif kind == 6:
            t = pairgen.mutate(rng, q, 0.01, 0.01, 0.01)
        else:              # repeat-rich two-letter alphabet
            q = pairgen.rand_dna(rng, L, "AC")
            t = pairgen.mutate(rng, q, 0.03, 0.02, 0.02, "AC")
        qs.append(q.encode())
        ts.append((t or "A").encode())
    return qs, ts


@pytest.mark.parametrize("k", [None, 0, 40, 100, 300, 2000])
def test_banded_global_rounds_match_the_oracle(k):
    qs, ts = _band_pairs(4242 if k is None else 4242 + k, 480, 30, 3000)
    res = batch.ed_align_batch(qs, ts, mode="NW", task="distance", k=-1 if k is None else k)
    ref = O.oracle_ed(qs, ts, "NW", "distance", k)
    compare(qs, ts, res, ref, "banded NW k=%r" % k)


def test_banded_global_equals_full_columns():
    """DB200_ED_BANDED=0 switches the banded rounds off: both paths must agree (and with the oracle) on long pairs."""
    import os
    a, b, _ = workloads.insertion_linking(300, seed=12, platform="ont", lo=200, hi=10000)
    qs, ts = a.tolist(), b.tolist()
    banded = batch.ed_align_batch(qs, ts, mode="NW", task="locations")
    os.environ["DB200_ED_BANDED"] = "0"
    try:
        full = batch.ed_align_batch(qs, ts, mode="NW", task="locations")
    finally:
        del os.environ["DB200_ED_BANDED"]
    compare(qs, ts, banded, full, "banded vs full")
    sub = list(range(0, 300, 10))
    ref = O.oracle_ed([qs[i] for i in sub], [ts[i] for i in sub], "NW", "locations")
    for j, i in enumerate(sub):
        assert banded.as_dict(i) == ref.as_dict(j)
