#!/usr/bin/env python
"""Replay a capture file (dysgu_b200.capture) through the GPU engine and compare every field of every alignment.

    python scripts/replay_capture.py pairs.jsonl[.gz] [--device 0] [--through dropin|batch]

BASELINE configs[0]: capture `dysgu call` on the bundled BAM with the reference's CPU wrappers, then replay here.  Pairs are
grouped by parameter set and aligned in one device batch per group (`--through batch`, default), or one by one through
the drop-in classes (`--through dropin`).  Exit status 0 only if every recorded field is reproduced: Smith-Waterman
optimal / suboptimal score, query and target begin / end, suboptimal end, cigar string; edit distance, alphabet length, the
full location list, cigar.  Prints one JSON summary line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dysgu_b200 import batch, capture  # noqa: E402
from dysgu_b200.edlib import align as gpu_align  # noqa: E402
from dysgu_b200.scikitbio import StripedSmithWaterman  # noqa: E402


def replay(records, through="batch", device=None):
    mismatches = []
    n_sw = n_ed = batches = 0
    if through == "dropin":
        for i, r in enumerate(records):
            if r["kind"] == "sw":
                n_sw += 1
                a = StripedSmithWaterman(r["query"], **r["params"])(r["target"])
                got = {k: a[k] for k in capture.SW_RESULT_KEYS}
            else:
                n_ed += 1
                p = r["params"]
                eq = p.get("additionalEqualities")
                d = gpu_align(r["query"], r["target"], p["mode"], p["task"], p["k"], [tuple(x) for x in eq] if eq else None)
                got = {"editDistance": d["editDistance"], "alphabetLength": d["alphabetLength"], "locations": [list(x) for x in d["locations"]],
                       "cigar": d.get("cigar")}
            if got != r["result"]:
                mismatches.append((i, r["site"], got, r["result"]))
        return n_sw, n_ed, n_sw + n_ed, mismatches
    groups = {}
    for i, r in enumerate(records):
        groups.setdefault((r["kind"], json.dumps(r["params"], sort_keys=True)), []).append(i)
    for (kind, pj), idx in groups.items():
        p = json.loads(pj)
        batches += 1
        if kind == "sw":
            n_sw += len(idx)
            # the constructor turns its kwargs into engine parameters (flag, mask length, index base): let it
            probes = [StripedSmithWaterman(records[i]["query"], **p) for i in idx]
            s0 = probes[0]
            res = batch.sw_align_batch([s._query_bytes for s in probes], [records[i]["target"].encode("ascii") for i in idx],
                                       match=s0.match_score, mismatch=s0.mismatch_score, gap_open=s0.gap_open_penalty,
                                       gap_extend=s0.gap_extend_penalty, score_size=s0.score_size, flag=s0.bit_flag, filters=s0.score_filter,
                                       filterd=s0.distance_filter, mask_len=[s.mask_length for s in probes], matrix=s0.matrix, device=device)
            base = s0.index_starts_at
            cig = res.cigar_strings()
            for k, i in enumerate(idx):
                f = res.fixed[k]
                got = {"optimal_alignment_score": int(f[0]), "suboptimal_alignment_score": int(f[1]),
                       "query_begin": int(f[4]) + base if f[4] >= 0 else -1, "query_end": int(f[5]) + base,
                       "target_begin": int(f[2]) + base if f[2] >= 0 else -1, "target_end_optimal": int(f[3]) + base,
                       "target_end_suboptimal": int(f[6]) + base, "cigar": cig[k]}
                if res.status[k] != 0 or got != records[i]["result"]:
                    mismatches.append((i, records[i]["site"], got, records[i]["result"]))
        else:
            n_ed += len(idx)
            eq = p.get("additionalEqualities")
            res = batch.ed_align_batch([records[i]["query"].encode("latin-1") for i in idx], [records[i]["target"].encode("latin-1") for i in idx],
                                       mode=p["mode"], task=p["task"], k=-1 if p["k"] is None else p["k"],
                                       additional_equalities=[tuple(x) for x in eq] if eq else None, device=device)
            for k, i in enumerate(idx):
                d = res.as_dict(k)
                got = {"editDistance": d["editDistance"], "alphabetLength": d["alphabetLength"], "locations": [list(x) for x in d["locations"]],
                       "cigar": d.get("cigar")}
                if d["status"] != 0 or got != records[i]["result"]:
                    mismatches.append((i, records[i]["site"], got, records[i]["result"]))
    return n_sw, n_ed, batches, mismatches


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("capture")
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--through", choices=["batch", "dropin"], default="batch")
    ap.add_argument("--show", type=int, default=5, help="mismatches to print")
    a = ap.parse_args(argv)
    records = capture.load(a.capture)
    n_sw, n_ed, batches, mism = replay(records, a.through, a.device)
    sites = {}
    for r in records:
        sites[r["site"]] = sites.get(r["site"], 0) + 1
    print(json.dumps({"capture": a.capture, "through": a.through, "sw_pairs": n_sw, "ed_pairs": n_ed, "device_calls": batches, "sites": sites,
                      "mismatches": len(mism)}))
    for m in mism[:a.show]:
        print("MISMATCH record %d (%s)\n  got  %s\n  want %s" % m, file=sys.stderr)
    return 1 if mism else 0


if __name__ == "__main__":
    sys.exit(main())
