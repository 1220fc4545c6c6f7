"""Batched equivalents of dysgu's alignment call sites (the callers either side of the path).

The drop-in classes are synchronous per pair; these helpers take what a dysgu loop would otherwise feed to
`StripedSmithWaterman(...)(...)` / `edlib.align(...)` one at a time, run ONE device batch, and replay the
reference's decision logic on the integer results.  Results are pure functions of the inputs, so collecting
candidates speculatively and replaying control flow afterwards cannot change a VCF.

    check_contig_match_batch   consensus.check_contig_match(a, b, return_int=True)      consensus.pyx:977-1018
    contig_alignments_batch    consensus.check_contig_match(a, b, return_int=False)      consensus.pyx:977-989
    badclip_scores_batch       the two alignments per contig in post_call.get_badclip_metric   post_call.py:50-79
    clip_align_matches_batch   filter_normals.clip_align_matches                         filter_normals.py:657-697
    remap_clip_batch           the edlib -> edlib -> SSW chain of re_map.process_contig   re_map.py:142-153, 208-223
    matches_with_max_gap_batch merge_svs.matches_with_max_gap on the cigar arrays of a batch    merge_svs.pyx:196-222
"""
from __future__ import annotations

import zlib
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import batch as _batch
from .scikitbio._ssw_wrapper import AlignmentStructure

_COMP = bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn")


def _b(s) -> bytes:
    return s if isinstance(s, bytes) else str(s).encode("ascii")


def reverse_complement(seq) -> bytes:
    return _b(seq).translate(_COMP)[::-1]


# ---------------------------------------------------------------------------------------------------------------
def contig_alignments_batch(pairs: Sequence[Tuple[str, str]], device=None, want_cigar=True) -> List[Optional[AlignmentStructure]]:
    """check_contig_match(a, b, return_int=False) for every (a, b): an AlignmentStructure, or 0 when a or b is empty.

    Scoring (2,-3,10,1), inputs cut to 10,000 bases, sequences suppressed (consensus.pyx:978-989)."""
    idx, qs, ts = [], [], []
    for i, (a, b) in enumerate(pairs):
        if not a or not b:
            continue
        idx.append(i)
        qs.append(_b(a)[:10000])
        ts.append(_b(b)[:10000])
    out: List[Optional[AlignmentStructure]] = [0] * len(pairs)
    if not idx:
        return out
    r = _batch.sw_align_batch(qs, ts, match=2, mismatch=-3, gap_open=10, gap_extend=1, flag=1 if want_cigar else 8, device=device,
                              want_cigar=want_cigar)
    for k, i in enumerate(idx):
        if r.status[k] != 0:
            raise RuntimeError("alignment failed: %s" % _batch.PAIR_STATUS.get(int(r.status[k]), r.status[k]))
        out[i] = AlignmentStructure(r.fixed[k], r.cigar_of(k).copy() if want_cigar else np.zeros(0, np.uint32), "", "", 0)
    return out


def check_contig_match_batch(pairs: Sequence[Tuple[str, str]], diffs=8, ol_length=21, device=None) -> List[int]:
    """check_contig_match(a, b, return_int=True): 1 if the two contigs match, else 0 (consensus.pyx:991-1018).

    Only score and begin/end coordinates are consumed, so the batch runs without the traceback stage (flag 8)."""
    alns = contig_alignments_batch(pairs, device=device, want_cigar=False)
    out = []
    for (a, b), al in zip(pairs, alns):
        if al == 0:
            out.append(0)
            continue
        la, lb = min(len(a), 10000), min(len(b), 10000)
        qs, qe = al.query_begin, al.query_end
        als, ale = al.target_begin, al.target_end_optimal
        extent_left = min(qs, als)
        extent_right = min(la - qe, lb - ale)
        total_overhangs = extent_left + extent_right
        aln_s = al.optimal_alignment_score
        expected = (qe - qs) * 2
        if expected < 2 * ol_length:
            out.append(0)
            continue
        rel_diff = (aln_s - total_overhangs) / expected
        diff = expected - aln_s + total_overhangs
        if aln_s > 100 and total_overhangs < qe - qs and rel_diff > 0.7:
            out.append(1)
        elif not diff > diffs:
            out.append(1)
        else:
            out.append(0)
    return out


def matches_with_max_gap_batch(results, max_gap: int, gaps_tol: float) -> np.ndarray:
    """merge_svs.matches_with_max_gap for every pair of a batch (merge_svs.pyx:196-222): the matched bases, or 0 when
    an I/D run is longer than max_gap or gaps / matches exceeds gaps_tol (float32 division as in the reference)."""
    st = results.cigar_stats()
    matches, gaps, longest = st[:, 0], st[:, 1], st[:, 2]
    ratio = np.divide(gaps.astype(np.float32), matches.astype(np.float32), out=np.zeros(len(st), np.float32), where=matches > 0)
    bad = (longest > max_gap) | ((matches > 0) & (ratio > np.float32(gaps_tol)))
    return np.where(bad, 0, matches)


# ---------------------------------------------------------------------------------------------------------------
def badclip_scores_batch(contigs: Sequence[str], clips: Sequence[str], device=None) -> List[Tuple[int, int]]:
    """(forward score, reverse-complement score) per (contig, clip): StripedSmithWaterman(contig, gap_extend_penalty=1,
    suppress_sequences=True) against clip and revcomp(clip) - scoring (2,-3,5,1), score only (post_call.py:53-62)."""
    qs, ts = [], []
    for c, f in zip(contigs, clips):
        qs += [_b(c), _b(c)]
        ts += [_b(f), reverse_complement(f)]
    if not qs:
        return []
    r = _batch.sw_align_batch(qs, ts, match=2, mismatch=-3, gap_open=5, gap_extend=1, flag=0, device=device, want_cigar=False)
    s = r.score1
    return [(int(s[2 * i]), int(s[2 * i + 1])) for i in range(len(contigs))]


# ---------------------------------------------------------------------------------------------------------------
RIGHT_CLIP = 1  # SeqType.RIGHT_CLIP in filter_normals


def clip_align_matches_batch(seq1s, seq2s, clip_sides, paired_end: bool, is_alt_seq=False, device=None) -> List[bool]:
    """filter_normals.clip_align_matches for many (seq1, seq2): SSW (2,-3,4,1) when both are < 16 kb and the shorter
    is < 2 kb, else edlib HW/locations; plus the prefix/suffix shortcut for tiny clips and the zlib heuristic."""
    n = len(seq1s)
    out: List[Optional[bool]] = [None] * n
    sw_i, ed_i, min_seqs = [], [], [0] * n
    for i, (s1, s2, side) in enumerate(zip(seq1s, seq2s, clip_sides)):
        if not s1 or not s2:
            out[i] = False
            continue
        min_seq = len(s2) if is_alt_seq else min(len(s1), len(s2))
        min_seqs[i] = min_seq
        if min_seq < 8:
            if side == RIGHT_CLIP:
                part = s2[int(len(s2) / 2) + (1 if len(s2) % 2 == 0 else 0):]
                out[i] = part.startswith(s1)
            else:
                part = s2[:int(len(s2) / 2)]
                out[i] = part.endswith(s1)
            continue
        if len(s1) < 16000 and len(s2) < 16000 and min_seq < 2000:
            sw_i.append(i)
        else:
            ed_i.append(i)
    if sw_i:
        r = _batch.sw_align_batch([_b(seq1s[i]) for i in sw_i], [_b(seq2s[i]) for i in sw_i], match=2, mismatch=-3, gap_open=4,
                                  gap_extend=1, flag=0, device=device, want_cigar=False)
        for k, i in enumerate(sw_i):
            if r.score1[k] / (min_seqs[i] * 2) > 0.7:
                out[i] = True
    if ed_i:
        r = _batch.ed_align_batch([_b(seq1s[i]) for i in ed_i], [_b(seq2s[i]) for i in ed_i], mode="HW", task="locations", device=device)
        for k, i in enumerate(ed_i):
            d = r.as_dict(k)
            if d["locations"]:
                pct = 1 - (d["editDistance"] / min_seqs[i])
                if pct > 0.7 and (d["locations"][0][1] - d["locations"][0][0]) / min_seqs[i] > 0.7:
                    out[i] = True
    for i in range(n):
        if out[i] is None:
            out[i] = False
            if paired_end:
                s1, s2 = _b(seq1s[i]), _b(seq2s[i])
                if len(zlib.compress(s1 + s2)) / (len(s1) + len(s2)) < 0.5:
                    out[i] = True
    return out  # type: ignore[return-value]


# ---------------------------------------------------------------------------------------------------------------
def merge_align_regions(locations):
    """Same selection as re_map.merge_align_regions (re_map.py:85-111)."""
    if len(locations) <= 1:
        return [list(x) for x in locations]
    merged = []
    for s, e in locations:
        if not merged:
            merged.append([s, e])
        last = merged[-1]
        if abs(s - last[0]) < 10 and abs(e - last[1]) < 10:
            merged[-1][1] = e
        else:
            merged.append([s, e])
    best_i, dist = 0, 100_000
    for i, (s, e) in enumerate(merged):
        if abs((e - s) - 1000) < dist:
            dist = abs((e - s) - 1000)
            best_i = i
    return [merged[best_i]]


def remap_clip_batch(clips: Sequence[str], adjacent_refs: Sequence[str], windows: Sequence[str], device=None):
    """The three alignments re_map.process_contig issues per soft clip, for many clips in three device batches:

      1. edlib HW/locations  clip.upper()  vs  adjacent reference of the clip's length   (matches_adjacent_ref_seq)
      2. edlib HW/locations  clip.upper()  vs  the +-500 bp reference window             (process_contig:209)
      3. SSW (2,-8,6,1)      query = window[l_start:l_end+1], target = clip              (process_contig:220-223)

    Returns one dict per clip: adjacent_match (bool), locs (merged region or []), ref_offset (l_start) and the
    AlignmentStructure of step 3 with sequences attached (None when steps 1/2 stop the reference early)."""
    n = len(clips)
    up = [_b(c).upper() for c in clips]
    r1 = _batch.ed_align_batch(up, [_b(a).upper() for a in adjacent_refs], mode="HW", task="locations", device=device)
    r2 = _batch.ed_align_batch(up, [_b(w) for w in windows], mode="HW", task="locations", device=device)
    out, todo, qs, ts = [], [], [], []
    for i in range(n):
        d1, d2 = r1.as_dict(i), r2.as_dict(i)
        clip_length = len(clips[i])
        adj = bool(d1["locations"]) and d1["editDistance"] / clip_length < 0.19 and \
            d1["locations"][0][1] - d1["locations"][0][0] > clip_length * 0.85
        locs = merge_align_regions(d2["locations"]) if not adj else []
        rec = {"adjacent_match": adj, "locs": locs, "ref_offset": None, "alignment": None}
        if not adj and locs:
            l_start, l_end = locs[0]
            rec["ref_offset"] = l_start
            ref_seq2 = windows[i][l_start:l_end + 1]
            if ref_seq2:
                todo.append(i)
                qs.append(_b(ref_seq2))
                ts.append(_b(clips[i]))
                rec["_ref_seq2"] = ref_seq2
        out.append(rec)
    if todo:
        r3 = _batch.sw_align_batch(qs, ts, match=2, mismatch=-8, gap_open=6, gap_extend=1, device=device)
        for k, i in enumerate(todo):
            out[i]["alignment"] = AlignmentStructure(r3.fixed[k], r3.cigar_of(k).copy(), out[i].pop("_ref_seq2"), clips[i], 0)
    return out
