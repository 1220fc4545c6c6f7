"""GPU: batched call-site helpers (dysgu_b200.callers) against golden vectors produced by the reference's wrappers."""
import json
import os

import pytest

from dysgu_b200 import callers

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callers.json")))


def test_check_contig_match_batch():
    g = G["check_contig_match"]
    assert callers.check_contig_match_batch([tuple(p) for p in g["pairs"]]) == g["result"]


def test_contig_alignments_batch_has_cigars():
    g = G["check_contig_match"]
    alns = callers.contig_alignments_batch([tuple(p) for p in g["pairs"][:20]])
    assert all(a == 0 or (a.cigar and a.cigar[-1] in "MID") for a in alns)


def test_badclip_scores_batch():
    g = G["badclip"]
    got = callers.badclip_scores_batch(g["contigs"], g["clips"])
    assert [list(x) for x in got] == g["result"]


def test_clip_align_matches_batch():
    g = G["clip_align_matches"]
    assert callers.clip_align_matches_batch(g["seq1"], g["seq2"], g["sides"], True) == g["paired"]
    assert callers.clip_align_matches_batch(g["seq1"], g["seq2"], g["sides"], False) == g["single"]


def test_remap_clip_batch():
    g = G["remap"]
    got = callers.remap_clip_batch(g["clips"], g["adjacent"], g["windows"])
    for rec, want in zip(got, g["result"]):
        assert rec["adjacent_match"] == want["adjacent_match"]
        assert [list(x) for x in rec["locs"]] == want["locs"]
        if want["aln"] is None:
            assert rec["alignment"] is None
        else:
            a = rec["alignment"]
            assert [a.optimal_alignment_score, a.query_begin, a.query_end, a.target_begin, a.target_end_optimal, a.cigar,
                    a.aligned_query_sequence, a.aligned_target_sequence] == want["aln"]
