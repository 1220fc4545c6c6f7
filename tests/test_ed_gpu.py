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
