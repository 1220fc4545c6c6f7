"""TEST INFRASTRUCTURE ONLY: pin the CPU oracle (restatement) against the reference's own compiled cores.

usage: python oracle/validate_oracle.py [--pairs N] [--max-len L] [--seed S]
Prints the number of compared pairs and differences per stage; exits 1 on any difference.
"""
import argparse
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import oracle as O  # noqa: E402
import pairgen  # noqa: E402

SCORINGS = [(2, -8, 6, 1), (2, -3, 10, 1), (2, -3, 5, 1), (2, -3, 4, 1), (2, -3, 5, 2), (1, -1, 2, 1), (3, -2, 7, 3), (5, -4, 12, 2)]
# gap_open <= gap_extend is excluded: there the reference's lazy-F shortcuts (ssw.c:255-273, 470-480) stop
# propagating vertical gaps and its scores depend on the striped layout ("Known deviations").


def chunk_sw(args):
    seed, n, max_len = args
    qs, ts = pairgen.sw_pairs(seed, n, max_len)
    m, x, go, ge = SCORINGS[seed % len(SCORINGS)]
    ss = [2, 2, 2, 1][seed % 4]
    flag = [1, 1, 1, 0, 8][seed % 5]
    a = O.oracle_sw(qs, ts, m, x, go, ge, score_size=ss, flag=flag)
    b = O.ref_sw(qs, ts, m, x, go, ge, score_size=ss, flag=flag, nthreads=1)
    bad = 0
    for i in range(len(qs)):
        if a.status[i] != 0 or b.status[i] != 0 or a.as_tuple(i) != b.as_tuple(i):
            bad += 1
            if bad <= 3:
                print("SW DIFF", (m, x, go, ge, ss, flag), qs[i], ts[i], a.as_tuple(i), b.as_tuple(i), a.status[i], b.status[i])
    return len(qs), bad


def chunk_ed(args):
    seed, n, max_len = args
    qs, ts = pairgen.ed_pairs(seed, n, max_len)
    mode = ["HW", "NW", "SHW"][seed % 3]
    task = ["locations", "distance"][(seed // 3) % 2]
    k = [None, None, 5, 40][(seed // 6) % 4]
    a = O.oracle_ed(qs, ts, mode, task, k)
    b = O.ref_ed(qs, ts, mode, task, k, nthreads=1)
    bad = 0
    for i in range(len(qs)):
        if a.as_dict(i) != b.as_dict(i):
            bad += 1
            if bad <= 3:
                print("ED DIFF", mode, task, k, qs[i], ts[i], a.as_dict(i), b.as_dict(i))
    return len(qs), bad


def chunk_path(args):
    """task "path": the scalar restatement (ed_path_oracle.c) against the reference's alignment arrays"""
    seed, n, max_len = args
    qs, ts = pairgen.ed_pairs(seed, n, max_len)
    mode = ["HW", "NW", "SHW"][seed % 3]
    k = [None, None, 5, 40][(seed // 3) % 4]
    a = O.oracle_ed_path(qs, ts, mode, k)
    b = O.ref_ed(qs, ts, mode, "path", k, nthreads=1)
    bad = 0
    for i in range(len(qs)):
        pa = a.path[a.path_off[i]:a.path_off[i] + a.path_len[i]]
        pb = b.path[b.path_off[i]:b.path_off[i] + b.path_len[i]]
        if a.as_dict(i) != b.as_dict(i) or not np.array_equal(pa, pb):
            bad += 1
            if bad <= 3:
                print("PATH DIFF", mode, k, qs[i], ts[i], a.as_dict(i), b.as_dict(i))
    return len(qs), bad, int(b.path_len.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=20000)
    ap.add_argument("--max-len", type=int, default=400)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--chunk", type=int, default=500)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    O.oracle_lib(); O.driver_lib()
    nchunks = max(1, a.pairs // a.chunk)
    jobs = [(a.seed * 100003 + i, a.chunk, a.max_len) for i in range(nchunks)]
    tot = bad = 0
    with ThreadPoolExecutor(a.threads) as ex:
        for n, b in ex.map(chunk_sw, jobs):
            tot += n; bad += b
    print("SW: compared %d pairs, %d differences" % (tot, bad))
    tot2 = bad2 = 0
    with ThreadPoolExecutor(a.threads) as ex:
        for n, b in ex.map(chunk_ed, jobs):
            tot2 += n; bad2 += b
    print("ED: compared %d pairs, %d differences" % (tot2, bad2))
    tot3 = bad3 = ops3 = 0
    with ThreadPoolExecutor(a.threads) as ex:
        for n, b, o in ex.map(chunk_path, jobs):
            tot3 += n; bad3 += b; ops3 += o
    print("ED path: compared %d pairs (%d edit operations), %d differences" % (tot3, ops3, bad3))
    sys.exit(1 if (bad or bad2 or bad3) else 0)


if __name__ == "__main__":
    main()
