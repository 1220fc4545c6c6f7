"""Manual GPU debugging helper: prints per-class parity statistics (not a pytest module)."""
import sys
import time
import traceback

import conftest  # noqa: F401  (sys.path)
import pairgen
from oracle import oracle as O
from dysgu_b200 import batch, workloads


def run(tag, qs, ts, **kw):
    t0 = time.time()
    try:
        res = batch.sw_align_batch(qs, ts, **kw)
    except Exception:
        print(tag, "EXCEPTION")
        traceback.print_exc()
        return
    t1 = time.time()
    ref = O.oracle_sw(qs, ts, kw.get("match", 2), kw.get("mismatch", -3), kw.get("gap_open", 5), kw.get("gap_extend", 2),
                      score_size=kw.get("score_size", 2), flag=kw.get("flag", 1))
    names = batch.SW_FIELDS
    nbad = [0] * 9
    shown = 0
    for i in range(len(qs)):
        a, b = res.as_tuple(i), ref.as_tuple(i)
        if res.status[i] != 0:
            nbad[8] += 1
        for f in range(8):
            if a[f] != b[f]:
                nbad[f] += 1
        if (a != b or res.status[i] != 0) and shown < 4:
            shown += 1
            print("  DIFF", i, "st", res.status[i], "m", len(qs[i]), "n", len(ts[i]), "\n    gpu", a, "\n    ref", b)
            if len(qs[i]) < 100:
                print("    q", qs[i], "\n    t", ts[i])
    print("%-28s n=%d gpu %.3fs  mismatches per field %s status!=0 %d" %
          (tag, len(qs), t1 - t0, dict(zip(names[:7] + ("cigar",), nbad[:8])), nbad[8]))
    sys.stdout.flush()


if __name__ == "__main__":
    for mx in (30, 60, 120, 250, 500, 1000):
        qs, ts = pairgen.sw_pairs(mx, 600, mx)
        run("classes max_len=%d" % mx, qs, ts, match=2, mismatch=-8, gap_open=6, gap_extend=1)
    qs, ts = pairgen.sw_pairs(5, 2000, 300)
    run("scoring 2,-3,10,1", qs, ts, match=2, mismatch=-3, gap_open=10, gap_extend=1)
    run("score_size=1", qs, ts, score_size=1)
    run("flag=0", qs, ts, flag=0)
    q, t, prm = workloads.remap_windows(3000, seed=3)
    run("remap windows", q.tolist(), t.tolist(), **prm)
    a, b, prm = workloads.contig_pairs(12, seed=5, platform="ont", lo=50, hi=5000)
    run("contigs (tiling)", a.tolist(), b.tolist(), **prm)
