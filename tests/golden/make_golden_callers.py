"""Golden vectors for dysgu_b200.callers, produced by running the reference's decision logic over the reference's own
compiled wrappers (oracle/_ref/pyref, built by oracle/build_ref.sh).  The decision logic below restates, for the test
only, what consensus.check_contig_match (consensus.pyx:977-1018), post_call.get_badclip_metric (post_call.py:53-62),
filter_normals.clip_align_matches (filter_normals.py:657-697) and re_map.process_contig (re_map.py:142-153, 208-223) do
with the alignment results.   python tests/golden/make_golden_callers.py
"""
import json
import os
import random
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pyref"))
from dysgu.scikitbio._ssw_wrapper import StripedSmithWaterman  # noqa: E402  (reference build)
from dysgu.edlib.edlib import align as edlib_align  # noqa: E402
import pairgen  # noqa: E402


def ref_check_contig_match(a, b, diffs=8, ol_length=21):
    if not a or not b:
        return 0
    a, b = a[:10000], b[:10000]
    al = StripedSmithWaterman(str(a), suppress_sequences=True, match_score=2, mismatch_score=-3, gap_open_penalty=10,
                              gap_extend_penalty=1)(str(b))
    qs, qe, als, ale = al.query_begin, al.query_end, al.target_begin, al.target_end_optimal
    total = min(qs, als) + min(len(a) - qe, len(b) - ale)
    s = al.optimal_alignment_score
    expected = (qe - qs) * 2
    if expected < 2 * ol_length:
        return 0
    if s > 100 and total < qe - qs and (s - total) / expected > 0.7:
        return 1
    return 1 if not (expected - s + total) > diffs else 0


def ref_merge_align_regions(locations):
    """re_map.merge_align_regions restated for the generator only (re_map.py:85-111): neighbouring edlib hits whose start and
    end both lie within 10 bp of the block being grown are folded into it; of the blocks, the one whose length is closest
    to 1000 is kept.  Where the reference checkout exists the real function is used and this restatement is checked
    against it."""
    real = _real_merge_align_regions()
    mine = _restated_merge_align_regions(locations)
    if real is not None:
        assert real([tuple(x) for x in locations]) == mine, (locations, mine)
    return mine


def _restated_merge_align_regions(locations):
    if len(locations) <= 1:
        return locations
    blocks = []
    for s, e in locations:
        if blocks and abs(s - blocks[-1][0]) < 10 and abs(e - blocks[-1][1]) < 10:
            blocks[-1][1] = e
        else:
            blocks.append([s, e])
    best = min(range(len(blocks)), key=lambda i: (abs(blocks[i][1] - blocks[i][0] - 1000), i))
    return [blocks[best]]


_REAL = []


def _real_merge_align_regions():
    """The function object from the reference's own re_map.py (source parsed where it lies; nothing is copied)."""
    if not _REAL:
        path = os.path.join(os.environ.get("DYSGU_REFERENCE", "/root/reference"), "dysgu", "re_map.py")
        fn = None
        if os.path.exists(path):
            import ast
            tree = ast.parse(open(path).read())
            node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "merge_align_regions")
            scope = {}
            exec(compile(ast.Module([node], []), path, "exec"), scope)
            fn = scope["merge_align_regions"]
        _REAL.append(fn)
    return _REAL[0]


def revcomp(s):
    return s.translate(str.maketrans("ACGTacgt", "TGCAtgca"))[::-1]


def ref_clip_align_matches(seq1, seq2, clip_side, paired_end, is_alt_seq=False):
    if not seq1 or not seq2:
        return False
    min_seq = len(seq2) if is_alt_seq else min(len(seq1), len(seq2))
    if min_seq < 8:
        if clip_side == 1:
            return seq2[int(len(seq2) / 2) + (1 if len(seq2) % 2 == 0 else 0):].startswith(seq1)
        return seq2[:int(len(seq2) / 2)].endswith(seq1)
    if len(seq1) < 16000 and len(seq2) < 16000 and min_seq < 2000:
        a = StripedSmithWaterman(seq1, match_score=2, mismatch_score=-3, gap_open_penalty=4, gap_extend_penalty=1)(seq2)
        if a.optimal_alignment_score / (min_seq * 2) > 0.7:
            return True
    else:
        el = edlib_align(seq1, seq2, mode="HW", task="locations")
        if el and el["locations"]:
            if 1 - el["editDistance"] / min_seq > 0.7 and (el["locations"][0][1] - el["locations"][0][0]) / min_seq > 0.7:
                return True
    if paired_end:
        if len(zlib.compress(bytes(seq1.encode("ascii") + seq2.encode("ascii")))) / (len(seq1) + len(seq2)) < 0.5:
            return True
    return False


def main():
    rng = random.Random(99)
    out = {}
    # contig pairs
    qs, ts = pairgen.sw_pairs(555, 160, 500)
    pairs = [(q.decode(), t.decode()) for q, t in zip(qs, ts)] + [("", "ACGT"), ("ACGT", "")]
    out["check_contig_match"] = {"pairs": pairs, "result": [ref_check_contig_match(a, b) for a, b in pairs]}
    # bad-clip
    contigs, clips = [], []
    for _ in range(80):
        c = pairgen.rand_dna(rng, rng.randint(60, 600))
        L = rng.randint(9, 60)
        s = rng.randint(0, len(c) - L)
        f = pairgen.mutate(rng, c[s:s + L], 0.03, 0.01, 0.01) or "A"
        if rng.random() < 0.5:
            f = revcomp(f)
        contigs.append(c); clips.append(f.lower() if rng.random() < 0.5 else f)
    res = []
    for c, f in zip(contigs, clips):
        q = StripedSmithWaterman(c, gap_extend_penalty=1, suppress_sequences=True)
        res.append([q(f).optimal_alignment_score, q(revcomp(f)).optimal_alignment_score])
    out["badclip"] = {"contigs": contigs, "clips": clips, "result": res}
    # filter_normals clip matcher
    s1, s2, sides = [], [], []
    for i in range(120):
        kind = rng.randrange(5)
        if kind == 0:
            a = pairgen.rand_dna(rng, rng.randint(1, 7)); b = pairgen.rand_dna(rng, rng.randint(2, 30))
            if rng.random() < 0.5:
                b = b[:len(b) // 2] + a if rng.random() < 0.5 else a + b
        elif kind == 1:
            a = pairgen.rand_dna(rng, rng.randint(8, 300)); b = pairgen.mutate(rng, a, 0.05, 0.02, 0.02) or "A"
        elif kind == 2:
            a = pairgen.rand_dna(rng, rng.randint(8, 300)); b = pairgen.rand_dna(rng, rng.randint(8, 300))
        elif kind == 3:
            a = "AC" * rng.randint(10, 80); b = "AC" * rng.randint(10, 80)
        else:
            a = pairgen.rand_dna(rng, rng.randint(2000, 2600)); b = pairgen.mutate(rng, a, 0.02, 0.01, 0.01)
        s1.append(a); s2.append(b); sides.append(rng.randrange(2))
    out["clip_align_matches"] = {"seq1": s1, "seq2": s2, "sides": sides,
                                 "paired": [ref_clip_align_matches(a, b, sd, True) for a, b, sd in zip(s1, s2, sides)],
                                 "single": [ref_clip_align_matches(a, b, sd, False) for a, b, sd in zip(s1, s2, sides)]}
    # re_map chain
    clips, adj, wins, res = [], [], [], []
    for i in range(100):
        w = pairgen.rand_dna(rng, 1000)
        L = rng.randint(15, 150)
        if rng.random() < 0.7:
            s = rng.randint(0, 1000 - L)
            clip = pairgen.mutate(rng, w[s:s + L], 0.02, 0.005, 0.005) or "A"
        else:
            clip = pairgen.rand_dna(rng, L)
        clip = clip.lower()
        a = w[500 - len(clip):500] if rng.random() < 0.8 else pairgen.mutate(rng, clip.upper(), 0.05, 0.0, 0.0)[:len(clip)].ljust(len(clip), "A")
        clips.append(clip); adj.append(a); wins.append(w)
        d1 = edlib_align(clip.upper(), a.upper(), mode="HW", task="locations")
        adjm = bool(d1["locations"]) and d1["editDistance"] / len(clip) < 0.19 and d1["locations"][0][1] - d1["locations"][0][0] > len(clip) * 0.85
        rec = {"adjacent_match": adjm, "locs": [], "aln": None}
        if not adjm:
            el = edlib_align(clip.upper(), w, mode="HW", task="locations")
            locs = ref_merge_align_regions(el["locations"])
            rec["locs"] = locs
            if locs:
                l0, l1 = locs[0]
                a2 = StripedSmithWaterman(w[l0:l1 + 1], match_score=2, mismatch_score=-8, gap_open_penalty=6, gap_extend_penalty=1)(clip)
                rec["aln"] = [a2.optimal_alignment_score, a2.query_begin, a2.query_end, a2.target_begin, a2.target_end_optimal, a2.cigar,
                              a2.aligned_query_sequence, a2.aligned_target_sequence]
        res.append(rec)
    out["remap"] = {"clips": clips, "adjacent": adj, "windows": wins, "result": res}
    with open(os.path.join(HERE, "callers.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("callers.json", os.path.getsize(os.path.join(HERE, "callers.json")), "bytes")


if __name__ == "__main__":
    main()
