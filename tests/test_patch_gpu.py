"""GPU: caller-side batching (dysgu_b200.patch) and capture replay on the real engine.

  * the stand-in loops (tests/standin_loops.py: result-dependent chain, profile re-use, job lists with early exit, shuffles
    and a pool) run under the patch with the GPU engine; outcomes must equal tests/golden/patch_loops.json, which was
    produced by the same loops over the reference's compiled wrappers (tests/golden/make_golden_patch.py);
  * tests/golden/capture_remap.jsonl.gz - every alignment the reference's real re_map.remap_soft_clips and
    post_call.get_badclip_metric issued on synthetic events, recorded with results by dysgu_b200.capture - is replayed by
    scripts/replay_capture.py: every field of every pair equal, through batches and through the drop-in classes."""
import copy
import json
import os
import random
import subprocess
import sys

import pytest

from dysgu_b200 import patch
import standin_loops as SL

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
KEEP = ("remapped", "remap_score", "remap_cigar", "remap_start", "remap_end", "clip_span", "pos", "pos2", "fwd", "rev")


def golden():
    with open(os.path.join(G, "patch_loops.json")) as f:
        return json.load(f)


def _ns():
    return SL.make_namespace(patch.MemoStripedSmithWaterman, patch.memo_edlib, patch.InlinePool)


def test_chain_and_profile_loops_match_the_reference_run():
    g = golden()["chain_profile"]
    ref, events = SL.synth(g["seed"])
    ns = _ns()
    chain = patch.batched(ns.chain_loop, "chain", lambda a, k: ((patch._scratch_events(a[0]),) + tuple(a[1:]), k))
    profile = patch.batched(ns.profile_loop, "profile", lambda a, k: ((patch._scratch_events(a[0]),), k))
    kept = [e.idx for e in chain(events, ref)]
    profile(events)
    assert kept == g["kept"]
    got = json.loads(json.dumps([{k: getattr(e, k, None) for k in KEEP} for e in events]))
    assert got == g["events"]
    for name in ("chain", "profile"):
        st = patch.last_stats[name]
        assert st["misses"] == 0 and st["batches"] <= 8, (name, st)     # O(1) device batches for the whole loop


def test_job_loops_match_the_reference_run():
    for g in golden()["jobs"]:
        _, events = SL.synth(g["seed"], n_events=80)
        loop = patch.batched(_ns().job_loop, "jobs", lambda a, k: ((patch._scratch_graph(a[0]),) + tuple(a[1:]), k))
        random.seed(g["rng"])
        graph = {}
        n = loop(graph, events, same_sample=g["same_sample"], procs=g["procs"])
        assert n == g["added"] and sorted([a, b] for a, s in graph.items() for b in s if a < b) == g["edges"], g["procs"]
        assert random.random() == g["rng_after"]
        st = patch.last_stats["jobs"]
        assert st["misses"] == 0 and st["batches"] <= 12, st


def test_unpatched_dropins_give_the_same_outcome_one_pair_at_a_time():
    """The plain drop-in classes (a batch of one per call) through the same chain loop: same outcome, just slower."""
    from dysgu_b200.edlib import edlib as gpu_edlib
    from dysgu_b200.scikitbio import StripedSmithWaterman
    g = golden()["chain_profile"]
    ref, events = SL.synth(g["seed"])
    events = events[:25]
    ns = SL.make_namespace(StripedSmithWaterman, gpu_edlib)
    ns.chain_loop(events, ref)
    got = json.loads(json.dumps([{k: getattr(e, k, None) for k in KEEP if k not in ("fwd", "rev")} for e in events]))
    want = [{k: v for k, v in e.items() if k not in ("fwd", "rev")} for e in g["events"][:25]]
    assert got == want


@pytest.mark.parametrize("through", ["batch", "dropin"])
def test_captured_reference_run_replays_bit_exact(through):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "replay_capture.py"), os.path.join(G, "capture_remap.jsonl.gz"),
                        "--through", through], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr[-3000:])
    s = json.loads(r.stdout.strip().splitlines()[-1])
    assert s["mismatches"] == 0 and s["sw_pairs"] > 100 and s["ed_pairs"] > 30 and set(s["sites"]) == {"re_map", "post_call"}
    if through == "batch":
        assert s["device_calls"] <= 4
