"""GPU: the one-launch small-batch path (csrc/sw_small.cuh) - what a synchronous StripedSmithWaterman(q)(t) call of an
unmodified dysgu takes.  Batches of at most 64 pairs against the scalar oracle, field by field and cigar by cigar: every
adversarial class of tests/pairgen.py, the three score sizes, flags 0 / 1 / 8 / 9, fixed and automatic mask lengths
(second-best score with the padded profile rows), pairs that exceed the kernel's limits inside an otherwise small batch
(they fall back to the batch pipeline), and equality with the pipeline itself (DB200_NO_SMALL_PATH=1 in a subprocess)."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from dysgu_b200 import _lib, batch
from oracle import oracle as O

import pairgen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(qs, ts, **kw):
    got = batch.sw_align_batch(qs, ts, **kw)
    okw = {k: v for k, v in kw.items() if k != "mask_len"}
    ml = kw.get("mask_len")
    if ml is not None:
        okw["mask_len"] = np.ascontiguousarray(np.broadcast_to(np.asarray(ml, np.int32), (len(qs),)))
    ref = O.oracle_sw(qs, ts, **okw)
    bad_ref = ref.status != 0
    assert np.array_equal(got.status != 0, bad_ref), (got.status, ref.status)
    ok = ~bad_ref
    assert np.array_equal(got.fixed[ok], ref.fixed[ok]), [(i, got.fixed[i], ref.fixed[i]) for i in np.flatnonzero(ok) if not np.array_equal(got.fixed[i], ref.fixed[i])][:3]
    for i in np.flatnonzero(ok):
        assert np.array_equal(got.cigar_of(i), ref.cigar_of(i)), (i, qs[i][:60], ts[i][:60])
    return got


@pytest.mark.parametrize("scoring", [(2, -8, 6, 1), (2, -3, 10, 1), (2, -3, 5, 1), (2, -3, 4, 1), (2, -3, 5, 2), (3, -4, 7, 2)])
def test_small_batches_match_the_oracle(scoring):
    m, x, go, ge = scoring
    launches0 = int(_lib.lib().db200_kernel_launches(batch.default_device(), 1))
    for rep in range(12):
        qs, ts = pairgen.sw_pairs(1000 * rep + go, random.Random(rep).choice([1, 2, 7, 33, 64]), max_len=400)
        check(qs, ts, match=m, mismatch=x, gap_open=go, gap_extend=ge)
    launches = int(_lib.lib().db200_kernel_launches(batch.default_device(), 1))
    assert launches <= 12 * 3, launches      # one launch per batch unless a pair fell back to the pipeline


def test_flags_score_sizes_and_masks():
    rng = random.Random(5)
    qs, ts = pairgen.sw_pairs(77, 48, max_len=300)
    for flag in (0, 1, 8, 9):
        for score_size in (0, 1, 2):
            check(qs, ts, match=2, mismatch=-3, gap_open=5, gap_extend=2, flag=flag, score_size=score_size)
    for ml in (0, 14, 15, 16, 40, 1000):
        check(qs, ts, match=2, mismatch=-8, gap_open=6, gap_extend=1, mask_len=ml)
    check(qs, ts, mask_len=[rng.choice([3, 15, 20, 60, 200]) for _ in qs])


def test_second_best_with_padded_profile_rows():
    """Short queries against longer targets with repeats: score2 / ref_end2 come from per-column maxima that include the
    padded rows of the striped profile (ssw.c:109, 364)."""
    rng = random.Random(9)
    qs, ts = [], []
    for _ in range(60):
        L = rng.choice([9, 15, 16, 17, 23, 31, 33, 47])
        q = pairgen.rand_dna(rng, L)
        t = "".join(pairgen.mutate(rng, q, 0.05, 0.02, 0.02) + pairgen.rand_dna(rng, rng.randint(0, 25)) for _ in range(rng.randint(2, 9)))[:500]
        qs.append(q.encode()); ts.append(t.encode())
    got = check(qs, ts, match=2, mismatch=-3, gap_open=5, gap_extend=2)
    assert (got.fixed[:, 1] > 0).sum() > 30


def test_limits_and_fallback_inside_a_small_batch():
    rng = random.Random(11)
    q_long = pairgen.rand_dna(rng, 8192)
    q_over = pairgen.rand_dna(rng, 8200)
    t = pairgen.mutate(rng, q_long[4000:4300], 0.03, 0.01, 0.01)
    wide_t = pairgen.rand_dna(rng, 700)
    gappy_q = pairgen.rand_dna(rng, 500)
    gappy_t = gappy_q[:200] + gappy_q[320:]                    # one 120-base gap: the band doubles several times
    many_q = pairgen.rand_dna(rng, 3000)
    many_t = pairgen.mutate(rng, many_q[:480], 0.0, 0.08, 0.08)
    qs = [q_long, q_over, pairgen.mutate(rng, wide_t, 0.02, 0.01, 0.01), gappy_q, many_q, "ACGT"]
    ts = [t, t, wide_t, gappy_t, many_t, "ACGT"]
    qs = [s.encode() for s in qs]; ts = [s.encode() for s in ts]
    got = check(qs, ts, match=2, mismatch=-3, gap_open=10, gap_extend=1)
    assert (got.status == 0).all()
    # an empty sequence is reported per pair (the reference reads out of bounds there, ssw.c:195), the others are unaffected
    got = batch.sw_align_batch(qs + [b""], ts + [b"ACGT"], match=2, mismatch=-3, gap_open=10, gap_extend=1)
    assert got.status[6] == 1 and (got.status[:6] == 0).all() and got.cigar_off[7] == got.cigar_off[6]


def test_small_path_equals_the_batch_pipeline():
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import pairgen\nfrom dysgu_b200 import batch\n"
            "qs, ts = pairgen.sw_pairs(4242, 64, max_len=350)\n"
            "r = batch.sw_align_batch(qs, ts, match=2, mismatch=-8, gap_open=6, gap_extend=1)\n"
            "print(json.dumps([r.fixed.tolist(), r.cigar.tolist(), r.cigar_off.tolist(), r.status.tolist()]))\n" % (ROOT, os.path.join(ROOT, "tests")))
    outs = []
    for env in ({}, {"DB200_NO_SMALL_PATH": "1"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]
