"""TEST INFRASTRUCTURE: stand-in caller loops with the shapes of dysgu's alignment loops.

The real loops (re_map.py:366-481, post_call.py:27-81, merge_svs.pyx:385-541) cannot travel to the GPU box (no reference
checkout there, and merge_svs.pyx does not compile without pysam).  What dysgu_b200.patch has to cope with is their
SHAPE, and these loops - written for the tests, not taken from dysgu - have it:

  * `chain_loop`: per event a result-dependent chain (edit distance vs an adjacent slice -> edit distance vs a window ->
    Smith-Waterman on the slice the window search found), a second candidate only if the first one decided nothing,
    and mutation of the event once a decision is made;
  * `profile_loop`: one profile aligned against a clip and its reverse complement, scores only;
  * `job_loop`: candidate pairs drawn with random.shuffle truncation, state-dependent filtering (edge counters), jobs of up
    to nine sequence pairs with early exit, collected in lists and run through a Pool when procs > 1.

Each takes the two alignment entry points through a module-like namespace (`ns.StripedSmithWaterman`, `ns.edlib`,
`ns.Pool`) - the bindings dysgu_b200.patch.apply() replaces.  tests/test_patch_host.py additionally runs the reference's
real re_map.py / post_call.py where the checkout exists."""
import random
import types

_COMP = str.maketrans("ACGTacgt", "TGCAtgca")


class Event:
    """Plain attribute holder (copy.copy-able like dysgu's EventResult)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def state(self):
        return {k: v for k, v in sorted(self.__dict__.items())}


def make_namespace(ssw_cls, edlib_mod, pool_cls=None):
    ns = types.SimpleNamespace()
    ns.StripedSmithWaterman = ssw_cls
    ns.edlib = edlib_mod
    ns.Pool = pool_cls
    ns.chain_loop = lambda events, ref: chain_loop(ns, events, ref)
    ns.profile_loop = lambda events: profile_loop(ns, events)
    ns.job_loop = lambda G, events, **kw: job_loop(ns, G, events, **kw)
    return ns


def _chain_candidate(ns, e, clip, side, ref, pos):
    adj = ref[max(pos - len(clip), 0):pos] if side == 0 else ref[pos:pos + len(clip)]
    if not adj:
        return "skip"
    a = ns.edlib.align(clip.upper(), adj, mode="HW", task="locations")
    if a["locations"] and a["editDistance"] / len(clip) < 0.2 and a["locations"][0][1] - a["locations"][0][0] > 0.8 * len(clip):
        return "skip"                                   # the clip is just the neighbouring reference
    w0 = max(pos - 400, 0)
    window = ref[w0:pos + 400]
    w = ns.edlib.align(clip.upper(), window, mode="HW", task="locations")
    if not w["locations"]:
        return "skip"
    s, t = w["locations"][0][0], w["locations"][-1][1]
    aln = ns.StripedSmithWaterman(window[s:t + 1], match_score=2, mismatch_score=-8, gap_open_penalty=6, gap_extend_penalty=1)(clip)
    if aln.optimal_alignment_score < 24 or not aln.aligned_query_sequence:
        return None                                     # undecided: the caller tries the other clip
    e.remapped = 1
    e.remap_score = aln.optimal_alignment_score
    e.remap_cigar = aln.cigar
    e.remap_start = w0 + s + aln.query_begin
    e.remap_end = w0 + s + aln.query_end
    e.clip_span = (aln.target_begin, aln.target_end_optimal)
    if e.remap_start > e.pos:
        e.pos, e.pos2 = e.pos2, e.pos
    return "add"


def chain_loop(ns, events, ref):
    kept = []
    for e in events:
        e.remapped, e.remap_score = 0, 0
        cands = sorted([c for c in (e.clip_a, e.clip_b) if c], key=lambda c: -len(c[0]))
        verdict = None
        for clip, side in cands:
            verdict = _chain_candidate(ns, e, clip, side, ref, e.pos if (clip, side) == e.clip_a else e.pos2)
            if verdict in ("add", "skip"):
                break
        if verdict == "add" or (verdict is None and e.support > 4):
            kept.append(e)
    return kept


def profile_loop(ns, events):
    for e in events:
        e.fwd, e.rev = 0, 0
        for contig, clip in ((e.contig, e.clip_a), (e.contig2, e.clip_b)):
            if not contig or not clip or len(contig) >= 1000:
                continue
            prof = ns.StripedSmithWaterman(contig, gap_extend_penalty=1, suppress_sequences=True)
            f = clip[0]
            rc = f.translate(_COMP)[::-1]
            e.rev = max(e.rev, prof(rc).optimal_alignment_score)
            e.fwd = max(e.fwd, prof(f).optimal_alignment_score)
    return events


def _contigs_match(ns, a, b, want_alignment):
    al = ns.StripedSmithWaterman(a[:10000], suppress_sequences=True, match_score=2, mismatch_score=-3, gap_open_penalty=10,
                                 gap_extend_penalty=1)(b[:10000])
    if want_alignment:
        return al
    span = al.query_end - al.query_begin
    overhang = min(al.query_begin, al.target_begin) + min(len(a) - al.query_end, len(b) - al.target_end_optimal)
    return span >= 21 and 2 * span - al.optimal_alignment_score + overhang <= 8


def _job(ns, seqs_i, seqs_j, i, j, same_sample):
    for a in seqs_i:
        for b in seqs_j:
            if not a or not b:
                continue
            if same_sample:
                al = _contigs_match(ns, a, b, True)
                gaps = sum(int(n) for n, op in _cigar_runs(al.cigar) if op != "M")
                if gaps > 12 or al.optimal_alignment_score < 40:
                    return None
                return i, j
            if _contigs_match(ns, a, b, False):
                return i, j
    return None


def _cigar_runs(cigar):
    n = ""
    for ch in cigar:
        if ch.isdigit():
            n += ch
        else:
            yield n, ch
            n = ""


def job_loop(ns, G, events, same_sample=False, max_comparisons=6, procs=1):
    """G: dict of sets (adjacency), mutated.  Returns the number of edges added."""
    degree = {}
    jobs, added = [], 0
    pool = ns.Pool(procs) if procs > 1 else None

    def commit(edge):
        nonlocal added
        if not edge:
            return
        i, j = edge
        G.setdefault(i, set()).add(j)
        G.setdefault(j, set()).add(i)
        degree[i] = degree.get(i, 0) + 1
        degree[j] = degree.get(j, 0) + 1
        added += 1

    seen = set()
    for i, e in enumerate(events):
        near = [j for j, o in enumerate(events) if j != i and abs(o.pos - e.pos) < 60]
        if len(near) > max_comparisons:
            random.shuffle(near)
            near = near[:max_comparisons]
        for j in near:
            key = (min(i, j), max(i, j))
            if key in seen:
                continue
            seen.add(key)
            if degree.get(i, 0) > 3 or degree.get(j, 0) > 3:      # state-dependent filter
                continue
            o = events[j]
            item = (ns, [e.contig, e.contig2, e.alt], [o.contig, o.contig2, o.alt], key[0], key[1], same_sample)
            if pool is not None:
                jobs.append(item)
                if len(jobs) > 40:
                    for edge in pool.starmap(_job, jobs):
                        commit(edge)
                    jobs = []
            else:
                commit(_job(*item))
    if jobs:
        for edge in pool.starmap(_job, jobs):
            commit(edge)
    if pool is not None:
        pool.close()
        pool.join()
    return added


# ---- synthetic inputs -------------------------------------------------------------------------------------------------
def synth(seed, n_events=60, ref_len=6000):
    import pairgen
    rng = random.Random(seed)
    ref = pairgen.rand_dna(rng, ref_len)
    events = []
    for k in range(n_events):
        pos = rng.randint(500, ref_len - 900)
        kind = rng.randrange(5)
        gap = rng.randint(30, 300)

        def clip_from(p, L, err):
            return pairgen.mutate(rng, ref[p:p + L], err, err / 4, err / 4).lower()
        if kind == 0:      # deletion: the right clip maps further downstream
            ca = (clip_from(pos + gap, rng.randint(20, 140), 0.01), 1)
        elif kind == 1:    # left clip mapping upstream
            L = rng.randint(20, 140)
            ca = (clip_from(max(pos - gap - L, 0), L, 0.01), 0)
        elif kind == 2:    # clip equal to the adjacent reference (poor soft clip)
            L = rng.randint(20, 80)
            ca = (clip_from(pos, L, 0.0), 1)
        elif kind == 3:    # novel sequence
            ca = (pairgen.rand_dna(rng, rng.randint(15, 120)).lower(), rng.randrange(2))
        else:              # noisy copy
            ca = (clip_from(pos + gap, rng.randint(30, 200), 0.08), 1)
        cb = None
        if rng.random() < 0.5:
            cb = (clip_from(pos + rng.randint(40, 250), rng.randint(15, 90), rng.choice([0.0, 0.02, 0.3])), rng.randrange(2))
        flank = ref[max(pos - 150, 0):pos]
        contig = flank + ca[0] if ca[1] == 1 else ca[0] + ref[pos:pos + 150]
        contig2 = (ref[max(pos - 80, 0):pos + 40] + cb[0]) if cb else ""
        alt = pairgen.mutate(rng, contig[:120], 0.03, 0.01, 0.01) if rng.random() < 0.3 else ""
        events.append(Event(idx=k, pos=pos, pos2=pos + gap, clip_a=ca, clip_b=cb, support=rng.randint(1, 9), contig=contig, contig2=contig2,
                            alt=alt))
    # clusters of near-identical events for the job loop
    for k in range(n_events // 3):
        src = rng.choice(events[:n_events])
        events.append(Event(idx=len(events), pos=src.pos + rng.randint(-20, 20), pos2=src.pos2, clip_a=src.clip_a, clip_b=None,
                            support=rng.randint(1, 9), contig=pairgen.mutate(rng, src.contig, 0.01, 0.003, 0.003), contig2="",
                            alt=src.alt))
    return ref, events
