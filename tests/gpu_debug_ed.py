"""Manual GPU debugging helper for the edit-distance path (not a pytest module)."""
import sys
import time
import traceback

import conftest  # noqa: F401
import pairgen
from oracle import oracle as O
from dysgu_b200 import batch, workloads


def run(tag, qs, ts, mode, task, k=-1):
    t0 = time.time()
    try:
        res = batch.ed_align_batch(qs, ts, mode=mode, task=task, k=k)
    except Exception:
        print(tag, "EXCEPTION")
        traceback.print_exc()
        return
    t1 = time.time()
    ref = O.oracle_ed(qs, ts, mode, task, None if k < 0 else k)
    bad = {"editDistance": 0, "alphabetLength": 0, "ends": 0, "starts": 0, "status": 0}
    shown = 0
    for i in range(len(qs)):
        a, b = res.as_dict(i), ref.as_dict(i)
        if a == b:
            continue
        for f in ("editDistance", "alphabetLength", "status"):
            if a[f] != b[f]:
                bad[f] += 1
        if [e for _, e in a["locations"]] != [e for _, e in b["locations"]]:
            bad["ends"] += 1
        elif a["locations"] != b["locations"]:
            bad["starts"] += 1
        if shown < 4:
            shown += 1
            print("  DIFF", i, "m", len(qs[i]), "n", len(ts[i]), "\n    gpu", str(a)[:300], "\n    ref", str(b)[:300])
            if len(qs[i]) < 80 and len(ts[i]) < 120:
                print("    q", qs[i], "\n    t", ts[i])
    print("%-32s n=%d gpu %.3fs mismatches %s" % (tag, len(qs), t1 - t0, bad))
    sys.stdout.flush()


if __name__ == "__main__":
    for mx in (40, 130, 400, 1200):
        qs, ts = pairgen.ed_pairs(mx, 800, mx)
        for mode in ("HW", "NW", "SHW"):
            run("classes %s max_len=%d" % (mode, mx), qs, ts, mode, "locations")
    qs, ts = pairgen.ed_pairs(77, 1500, 300)
    run("HW distance", qs, ts, "HW", "distance")
    run("HW k=5", qs, ts, "HW", "locations", 5)
    run("NW k=40", qs, ts, "NW", "distance", 40)
    a, b, _ = workloads.insertion_linking(30, seed=9, platform="ont", lo=500, hi=9000)
    run("long NW", a.tolist(), b.tolist(), "NW", "distance")
    run("long HW", a.tolist(), b.tolist(), "HW", "locations")
