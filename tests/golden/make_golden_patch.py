"""Golden files for the caller-batching tests on the GPU box:

  patch_loops.json       outcome of the stand-in loops (tests/standin_loops.py) run over the REFERENCE's compiled wrappers
                         (oracle/_ref/pyref): what dysgu_b200.patch + the GPU engine must reproduce
  capture_remap.jsonl.gz every alignment the reference's REAL re_map.remap_soft_clips and post_call.get_badclip_metric
                         (loaded from the reference checkout) issued on synthetic events, with its result, recorded by
                         dysgu_b200.capture - replayed on the GPU by scripts/replay_capture.py

    python tests/golden/make_golden_patch.py      (needs /root/reference and oracle/_ref)"""
import copy
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import standin_loops as SL  # noqa: E402
import test_patch_host as T  # noqa: E402
from dysgu_b200 import capture, patch  # noqa: E402

KEEP = ("remapped", "remap_score", "remap_cigar", "remap_start", "remap_end", "clip_span", "pos", "pos2", "fwd", "rev")


def main():
    ssw, edl = T._pyref()
    out = {}
    ref, events = SL.synth(31)
    ns = SL.make_namespace(ssw.StripedSmithWaterman, edl, patch.InlinePool)
    ev = copy.deepcopy(events)
    kept = [e.idx for e in ns.chain_loop(ev, ref)]
    ns.profile_loop(ev)
    out["chain_profile"] = {"seed": 31, "kept": kept, "events": [{k: getattr(e, k, None) for k in KEEP} for e in ev]}
    out["jobs"] = []
    for same_sample, procs in ((False, 1), (False, 4), (True, 1), (True, 4)):
        _, events = SL.synth(32, n_events=80)
        random.seed(5)
        G = {}
        n = ns.job_loop(G, events, same_sample=same_sample, procs=procs)
        out["jobs"].append({"seed": 32, "rng": 5, "same_sample": same_sample, "procs": procs, "added": n,
                            "edges": sorted([a, b] for a, s in G.items() for b in s if a < b), "rng_after": random.random()})
    with open(os.path.join(HERE, "patch_loops.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    # the reference's real loops under the capture shim
    genome, events = T._dysgu_events(41, n=60)
    mods = T._fake_dysgu(ssw, edl)
    path = os.path.join(HERE, "capture_remap.jsonl.gz")
    session = capture.CaptureSession(path).attach(mods)
    ev = copy.deepcopy(events)
    mods["re_map"].remap_soft_clips(ev, T._Genome(genome))
    mods["post_call"].get_badclip_metric(ev, T._Counter(), T._Bam(), None)
    session.close()
    recs = capture.load(path)
    print("patch_loops.json", os.path.getsize(os.path.join(HERE, "patch_loops.json")), "bytes;", "capture_remap.jsonl.gz",
          os.path.getsize(path), "bytes,", len(recs), "alignments:", {s: sum(r["site"] == s for r in recs) for s in ("re_map", "post_call")})


if __name__ == "__main__":
    main()
