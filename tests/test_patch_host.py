"""CPU: dysgu_b200.patch (caller-side batching by speculative replay) and dysgu_b200.capture, checked without a GPU.

The replay logic only needs *an* engine, so here the batches are computed by the reference's compiled cores
(tests/cpu_engine.py) and the unpatched runs use the reference's compiled Python wrappers (oracle/_ref/pyref):
  * stand-in loops with the shapes of dysgu's loops (tests/standin_loops.py): identical outputs, zero misses, O(1) batches;
  * where the reference checkout exists: the reference's REAL re_map.remap_soft_clips and post_call.get_badclip_metric,
    loaded from /root/reference with stubs for the pysam-dependent helper modules, on synthetic events."""
import copy
import importlib.util
import os
import random
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")
REFERENCE = os.environ.get("DYSGU_REFERENCE", "/root/reference")

from dysgu_b200 import capture, patch  # noqa: E402
import standin_loops as SL  # noqa: E402


def _pyref():
    """The reference's compiled wrappers under private module names (the name `dysgu` stays free for the fake package below)."""
    import sysconfig
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    paths = (os.path.join(PYREF, "dysgu", "scikitbio", "_ssw_wrapper" + ext), os.path.join(PYREF, "dysgu", "edlib", "edlib" + ext))
    if not all(os.path.exists(p) for p in paths):
        pytest.skip("oracle/_ref/pyref not built")
    mods = []
    for name, p in zip(("pyref_ssw._ssw_wrapper", "pyref_edlib.edlib"), paths):
        if name in sys.modules:
            mods.append(sys.modules[name])
            continue
        spec = importlib.util.spec_from_file_location(name, p)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        sys.modules[name] = m
        mods.append(m)
    return mods


@pytest.fixture(scope="module")
def engine():
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    from cpu_engine import RefEngine
    return RefEngine()


def _states(events):
    return [e.state() for e in events]


def test_chain_loop_is_batched_with_identical_outcome(engine):
    ssw, edl = _pyref()
    ref, events = SL.synth(11)
    plain = SL.make_namespace(ssw.StripedSmithWaterman, edl)
    ev_a = copy.deepcopy(events)
    kept_a = [e.idx for e in plain.chain_loop(ev_a, ref)]

    ns = SL.make_namespace(patch.MemoStripedSmithWaterman, patch.memo_edlib)
    loop = patch.batched(ns.chain_loop, "chain", lambda a, k: ((patch._scratch_events(a[0]),) + tuple(a[1:]), k), engine)
    ev_b = copy.deepcopy(events)
    kept_b = [e.idx for e in loop(ev_b, ref)]
    st = patch.last_stats["chain"]
    assert kept_a == kept_b and _states(ev_a) == _states(ev_b)
    assert sum(e.remapped for e in ev_a) >= 10
    assert st["misses"] == 0 and st["hits"] > 100
    assert st["passes"] <= 6 and st["batches"] <= 8, st           # ~250 alignments in a handful of batches
    assert 200 <= st["batched_pairs"] <= st["hits"], st            # every alignment of the real run came out of a batch


def test_profile_loop_and_capture(engine):
    ssw, edl = _pyref()
    _, events = SL.synth(12)
    session = capture.CaptureSession()
    plain = SL.make_namespace(ssw.StripedSmithWaterman, edl)
    session.attach({"post_call": plain})
    ev_a = copy.deepcopy(events)
    plain.profile_loop(ev_a)
    session.close()
    assert plain.StripedSmithWaterman is ssw.StripedSmithWaterman          # bindings restored
    recs = session.records
    assert len(recs) > 50 and all(r["kind"] == "sw" and r["site"] == "post_call" and r["params"]["gap_extend_penalty"] == 1 for r in recs)

    ns = SL.make_namespace(patch.MemoStripedSmithWaterman, patch.memo_edlib)
    loop = patch.batched(ns.profile_loop, "profile", lambda a, k: ((patch._scratch_events(a[0]),), k), engine)
    ev_b = copy.deepcopy(events)
    loop(ev_b)
    st = patch.last_stats["profile"]
    assert _states(ev_a) == _states(ev_b)
    assert st["misses"] == 0 and st["passes"] == 2 and st["batches"] == 1, st


@pytest.mark.parametrize("same_sample,procs", [(False, 1), (False, 4), (True, 1), (True, 4)])
def test_job_loop_with_rng_state_and_pool(engine, same_sample, procs):
    ssw, edl = _pyref()
    _, events = SL.synth(13, n_events=80)
    plain = SL.make_namespace(ssw.StripedSmithWaterman, edl, patch.InlinePool)
    random.seed(99)
    Ga = {}
    na = plain.job_loop(Ga, events, same_sample=same_sample, procs=procs)
    after_a = random.random()

    ns = SL.make_namespace(patch.MemoStripedSmithWaterman, patch.memo_edlib, patch.InlinePool)
    loop = patch.batched(ns.job_loop, "jobs", lambda a, k: ((patch._scratch_graph(a[0]),) + tuple(a[1:]), k), engine)
    random.seed(99)
    Gb = {}
    nb = loop(Gb, events, same_sample=same_sample, procs=procs)
    after_b = random.random()
    st = patch.last_stats["jobs"]
    assert na == nb and Ga == Gb and na > 5
    assert after_a == after_b                       # the shuffles consumed the generator exactly once
    assert st["misses"] == 0 and st["batches"] <= 12 and st["batches"] < st["hits"] / 4, st


def test_placeholder_results_are_inert_and_keys_separate_parameters():
    memo = patch.Memo(engine=object())
    with memo.active(collecting=True):
        a = patch.MemoStripedSmithWaterman("ACGTACGTAC", gap_open_penalty=6)("ACGTACG")
        b = patch.MemoStripedSmithWaterman("ACGTACGTAC", gap_open_penalty=7)("ACGTACG")
        d = patch.memo_edlib.align("ACGT", "ACGT", mode="HW", task="locations")
    assert a.optimal_alignment_score == 0 and a.cigar == "" and not a.aligned_query_sequence
    assert b.optimal_alignment_score == 0 and d == {"editDistance": -1, "alphabetLength": 0, "locations": [], "cigar": None}
    assert len(memo.pending_sw) == 2 and len(memo.pending_ed) == 1


# ---- the reference's real loops ------------------------------------------------------------------------------------------
def _fake_dysgu(ssw_mod, edlib_mod):
    """A `dysgu` package holding the reference's real re_map.py and post_call.py (loaded from the reference checkout) over
    the given alignment modules, with small stand-ins for the helpers that live in pysam-dependent Cython modules."""
    if not os.path.exists(os.path.join(REFERENCE, "dysgu", "re_map.py")):
        pytest.skip("reference checkout not available")
    for k in [k for k in sys.modules if k == "dysgu" or k.startswith("dysgu.")]:
        del sys.modules[k]
    pkg = types.ModuleType("dysgu"); pkg.__path__ = []
    sk = types.ModuleType("dysgu.scikitbio"); sk.__path__ = []
    sk._ssw_wrapper = ssw_mod
    ed = types.ModuleType("dysgu.edlib"); ed.__path__ = []
    ed.align = edlib_mod.align
    ed.edlib = edlib_mod
    msu = types.ModuleType("dysgu.map_set_utils")
    msu.echo = lambda *a: None
    msu.is_overlapping = lambda x1, x2, y1, y2: max(x1, y1) <= min(x2, y2)

    def merge_intervals(intervals, srt=True, pad=0, add_indexes=False):   # helper stand-in (map_set_utils.pyx:35-90 semantics)
        iv = sorted(intervals, key=lambda t: (t[0], t[1])) if srt else list(intervals)
        out = []
        for c, s, e, *rest in iv:
            s, e = max(s - pad, 0), e + pad
            if out and out[-1][0] == c and s <= out[-1][2]:
                out[-1][2] = max(out[-1][2], e)
                if add_indexes:
                    out[-1][3].append(rest[0])
            else:
                out.append([c, s, e, [rest[0]]] if add_indexes else [c, s, e])
        return out
    msu.merge_intervals = merge_intervals
    cov = types.ModuleType("dysgu.coverage"); cov.merge_intervals = merge_intervals
    cons = types.ModuleType("dysgu.consensus"); cons.compute_rep = lambda s: 0.0
    io = types.ModuleType("dysgu.io_funcs")
    io.reverse_complement = lambda s, n: s.translate(str.maketrans("ACGTacgt", "TGCAtgca"))[::-1]
    io.intersecter = lambda regions, chrom, a, b: False
    mods = {"dysgu": pkg, "dysgu.scikitbio": sk, "dysgu.scikitbio._ssw_wrapper": ssw_mod, "dysgu.edlib": ed, "dysgu.edlib.edlib": edlib_mod,
            "dysgu.map_set_utils": msu, "dysgu.coverage": cov, "dysgu.consensus": cons, "dysgu.io_funcs": io}
    sys.modules.update(mods)
    out = {}
    for name in ("re_map", "post_call"):
        spec = importlib.util.spec_from_file_location("dysgu." + name, os.path.join(REFERENCE, "dysgu", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["dysgu." + name] = m
        setattr(pkg, name, m)
        spec.loader.exec_module(m)
        out[name] = m
    return out


class _Genome:
    def __init__(self, seq):
        self.seq, self.references, self.fetches = seq, ["chr1"], 0

    def get_reference_length(self, c):
        return len(self.seq)

    def fetch(self, c, s, e):
        self.fetches += 1
        return self.seq[s:e]


class _Counter:
    def sort_arrays(self):
        pass

    def count_near(self, tid, s, e):
        return (s + e) % 7


class _Bam:
    def gettid(self, c):
        return 0


def _dysgu_events(seed, n=70):
    import pairgen
    rng = random.Random(seed)
    genome = pairgen.rand_dna(rng, 9000)
    evs = []
    for k in range(n):
        pos = rng.randint(1200, 7000)
        gap = rng.randint(40, 400)
        kind = rng.randrange(4)
        L = rng.randint(18, 150)
        if kind == 0:     # right clip that maps downstream (deletion)
            clip = pairgen.mutate(rng, genome[pos + gap:pos + gap + L], 0.01, 0.002, 0.002)
            contig = genome[pos - 120:pos] + clip.lower()
        elif kind == 1:   # left clip that maps upstream
            clip = pairgen.mutate(rng, genome[pos - gap - L:pos - gap], 0.01, 0.002, 0.002)
            contig = clip.lower() + genome[pos:pos + 120]
        elif kind == 2:   # clip equals the neighbouring reference
            contig = genome[pos - 120:pos] + genome[pos:pos + L].lower()
        else:             # novel sequence
            contig = genome[pos - 120:pos] + pairgen.rand_dna(rng, L).lower()
        c2 = ""
        if rng.random() < 0.4:
            L2 = rng.randint(12, 100)
            c2 = genome[pos + gap - 100:pos + gap] + pairgen.mutate(rng, genome[pos - 300:pos - 300 + L2], 0.02, 0.0, 0.0).lower()
        e = SL.Event(chrA="chr1", chrB="chr1", posA=pos, posB=pos + rng.choice([0, gap]), spanning=0, svlen=rng.choice([0, 50, gap]),
                     svlen_precise=rng.choice([0, 0, 1]), contig=contig, contig2=c2, contig_ref_start=pos - 120, contig_ref_end=pos,
                     contig2_ref_start=pos + gap - 100, contig2_ref_end=pos + gap, contig_left_weight=float(rng.randint(5, 500)),
                     contig_right_weight=float(rng.randint(5, 500)), contig2_left_weight=float(rng.randint(5, 500)),
                     contig2_right_weight=float(rng.randint(5, 500)), ref_bases=rng.randint(20, 200), site_info=None, su=rng.randint(1, 12),
                     svtype="DEL", cipos95A=0, cipos95B=0, modified=0, ref_seq="", variant_seq="", left_ins_seq="", right_ins_seq="",
                     ref_rep=0.0)
        evs.append(e)
    return genome, evs


def test_reference_remap_soft_clips_and_badclip_metric_under_the_patch(engine):
    ssw, edl = _pyref()
    genome, events = _dysgu_events(21)
    mods = _fake_dysgu(ssw, edl)
    g1 = _Genome(genome)
    ev_a = copy.deepcopy(events)
    kept_a = mods["re_map"].remap_soft_clips(ev_a, g1)
    mods["post_call"].get_badclip_metric(ev_a, _Counter(), _Bam(), None)

    mods = _fake_dysgu(ssw, edl)
    done = patch.apply(mods, engine)
    assert {"re_map.StripedSmithWaterman", "re_map.edlib", "re_map.remap_soft_clips", "post_call.StripedSmithWaterman",
            "post_call.get_badclip_metric"} <= set(done)
    g2 = _Genome(genome)
    ev_b = copy.deepcopy(events)
    kept_b = mods["re_map"].remap_soft_clips(ev_b, g2)
    st_remap = dict(patch.last_stats["remap_soft_clips"])
    mods["post_call"].get_badclip_metric(ev_b, _Counter(), _Bam(), None)
    st_bad = dict(patch.last_stats["get_badclip_metric"])

    assert [id(e) for e in kept_b] == [id(ev_b[ev_a.index(e)]) for e in kept_a]     # same events kept, same order
    assert _states(ev_a) == _states(ev_b)
    assert sum(e.remapped for e in ev_a) >= 8 and any(e.ras or e.fas for e in ev_a)
    assert st_remap["misses"] == 0 and st_remap["hits"] >= 100 and st_remap["batches"] <= 8 and st_remap["passes"] <= 6, st_remap
    assert st_bad["misses"] == 0 and st_bad["batches"] == 1 and st_bad["passes"] == 2, st_bad
    assert g2.fetches == g1.fetches                 # the speculative passes re-used the fetched windows
    for k in [k for k in sys.modules if k == "dysgu" or k.startswith("dysgu.")]:
        del sys.modules[k]
