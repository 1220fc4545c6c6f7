#!/usr/bin/env python
"""Large randomised parity sweep: CUDA path (through the host C ABI) against the reference's own compiled
ssw.c / edlib.cpp (oracle/_ref, all host threads) - or the oracle restatement when oracle/_ref is absent.

Vectorised generators (numpy) so that millions of pairs are cheap; every integer output is compared:
SW score1, score2, ref_begin1, ref_end1, read_begin1, read_end1, ref_end2, the cigar array; edit distance,
alphabetLength, the full (start, end) location lists and, for task "path", every edit operation.

    python scripts/parity_sweep.py [--pairs 200000] [--seed 1]        # pairs per family and setting

Prints one line per (family, setting) and a total; exits non-zero on any difference.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dysgu_b200 import batch, workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def loglen(rng, lo, hi, n):
    return np.maximum(np.exp(rng.uniform(np.log(lo), np.log(hi + 1), size=n)).astype(np.int64), lo)


def periodic(rng, lens):
    """low-complexity sequences: a random unit of 1-6 bases repeated (ties, second-best scores)."""
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    unit_len = rng.integers(1, 7, size=len(lens))
    units = rng.integers(0, 4, size=(len(lens), 6), dtype=np.uint8)
    pair = np.repeat(np.arange(len(lens)), lens)
    pos = np.arange(total, dtype=np.int64) - np.repeat(off[:-1], lens)
    codes = units[pair, pos % unit_len[pair]]
    return W.SeqBatch(BASES[codes], off)


def sprinkle(rng, b, p, chars=b"NnXRY"):
    buf = b.buf.copy()
    m = rng.random(buf.size) < p
    ch = np.frombuffer(chars, dtype=np.uint8)
    buf[m] = ch[rng.integers(0, len(ch), size=int(m.sum()))]
    return W.SeqBatch(buf, b.off.copy())


def nonempty(b):
    """mutation may delete everything: give empty sequences one base"""
    lens = b.lengths
    if (lens > 0).all():
        return b
    nl = np.maximum(lens, 1)
    off = np.zeros(len(nl) + 1, dtype=np.int64)
    np.cumsum(nl, out=off[1:])
    buf = np.full(int(off[-1]), ord("A"), dtype=np.uint8)
    keep = np.repeat(lens > 0, nl)
    buf[keep] = b.buf
    return W.SeqBatch(buf, off)


def families(rng, n, max_len):
    """yield (name, queries, targets)"""
    L = loglen(rng, 1, max_len, n)
    a = W.random_dna(rng, L)
    yield "related 3%", a, nonempty(W.mutate(rng, a, 0.03, 0.01, 0.01))
    yield "related 15%", a, nonempty(W.mutate(rng, a, 0.10, 0.03, 0.03))
    yield "unrelated", a, W.random_dna(rng, loglen(rng, 1, max_len, n))
    # window vs clip and clip vs window
    wl = rng.integers(max(max_len // 2, 20), max_len + 1, size=n)
    win = W.random_dna(rng, wl)
    cl = np.minimum(rng.integers(9, 151, size=n), wl)
    st = (rng.random(n) * (wl - cl + 1)).astype(np.int64)
    clip = nonempty(W.mutate(rng, W.slices(win, st, cl), 0.02, 0.005, 0.005))
    yield "window/clip", win, W.lower(clip)
    yield "clip/window", clip, win
    # low complexity: periodic query against a mutated, shifted copy
    pl = loglen(rng, 4, min(max_len, 400), n)
    p = periodic(rng, pl)
    yield "periodic", p, nonempty(W.mutate(rng, p, 0.02, 0.01, 0.01))
    # tandem duplication in the target
    ul = rng.integers(5, 61, size=n)
    u = W.random_dna(rng, ul)
    reps = rng.integers(2, 5, size=n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(ul * reps, out=off[1:])
    pos = np.arange(int(off[-1]), dtype=np.int64) - np.repeat(off[:-1], ul * reps)
    src = np.repeat(u.off[:-1], ul * reps) + pos % np.repeat(ul, ul * reps)
    yield "tandem target", u, W.SeqBatch(u.buf[src], off)
    # N's, IUPAC, mixed case
    yield "N / IUPAC", sprinkle(rng, a, 0.05), sprinkle(rng, nonempty(W.mutate(rng, a, 0.03, 0.01, 0.01)), 0.05, b"NnKM")


def cmp_sw(r, ref):
    bad = (r.fixed[:, :7] != ref.fixed[:, :7]).any(axis=1) | (r.status != 0)
    cl_r, cl_f = np.diff(r.cigar_off), np.diff(ref.cigar_off)
    bad |= cl_r != cl_f
    if not bad.any() and len(r.cigar) == len(ref.cigar):
        if (r.cigar != ref.cigar).any():
            pair = np.repeat(np.arange(len(cl_r)), cl_r)
            bad[np.unique(pair[r.cigar != ref.cigar])] = True
    elif len(r.cigar) != len(ref.cigar) or (r.cigar != ref.cigar).any():
        for i in np.nonzero(~bad)[0]:
            if not np.array_equal(r.cigar_of(i), ref.cigar_of(i)):
                bad[i] = True
    return bad


def cmp_ed(r, ref):
    bad = (r.fixed != ref.fixed).any(axis=1)
    if len(r.loc) != len(ref.loc) or (np.diff(r.loc_off) != np.diff(ref.loc_off)).any():
        bad |= np.diff(r.loc_off) != np.diff(ref.loc_off)
    elif (r.loc != ref.loc).any():
        pair = np.repeat(np.arange(len(bad)), np.diff(r.loc_off))
        bad[np.unique(pair[(r.loc != ref.loc).any(axis=1)])] = True
    return bad


def cmp_path(r, ref):
    """edit operations of task "path": ours back to back (CSR), the reference driver's one slot per pair"""
    ln = ref.path_len.astype(np.int64)
    bad = np.diff(r.path_off) != ln
    if not bad.any() and ln.sum():
        idx = np.repeat(np.asarray(ref.path_off, dtype=np.int64) - (np.cumsum(ln) - ln), ln) + np.arange(int(ln.sum()), dtype=np.int64)
        diff = ref.path[idx] != r.path[:int(ln.sum())]
        if diff.any():
            bad[np.unique(np.repeat(np.arange(len(ln)), ln)[diff])] = True
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", default="all", choices=["all", "path", "nopath", "ed"], help="path: only the edlib task 'path' settings; ed: edit distance only (long sequences)")
    ap.add_argument("--forms", action="store_true", help="Smith-Waterman: also through the pre-packed input form and, for a sample, the one-launch path in batches of at most 64 pairs")
    ap.add_argument("--pairs", type=int, default=100000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-len", type=int, default=600)
    a = ap.parse_args()
    use_ref = O.have_ref()
    cores = os.cpu_count() or 1
    print("checker:", "reference ssw.c / edlib.cpp compiled in oracle/_ref (%d threads)" % cores if use_ref else "oracle restatement")
    rng = np.random.default_rng(a.seed)
    sw_settings = [  # (match, mismatch, gap_open, gap_extend, score_size, flag)
        (2, -8, 6, 1, 2, 1), (2, -3, 10, 1, 2, 1), (2, -3, 5, 1, 2, 0), (2, -3, 4, 1, 2, 8), (2, -3, 5, 2, 2, 1),
        (1, -4, 6, 1, 1, 1), (5, -4, 10, 2, 2, 1), (2, -8, 6, 1, 0, 1)]
    ed_settings = [("NW", "distance", None), ("HW", "locations", None), ("SHW", "locations", None), ("HW", "locations", 20),
                   ("NW", "locations", 40)]
    if use_ref and a.tasks != "nopath":   # the path checker is the compiled reference itself (no restatement under oracle/)
        ed_settings += [("NW", "path", None), ("HW", "path", None), ("SHW", "path", None), ("HW", "path", 20), ("NW", "path", 40)]
    if a.tasks == "path":
        sw_settings = []
        ed_settings = [e for e in ed_settings if e[1] == "path"]
    if a.tasks == "ed":
        sw_settings = []
    total = bad_total = ops_total = 0
    t0 = time.time()
    fams = list(families(rng, a.pairs, a.max_len))
    for name, q, t in fams:
        qc, tc = (q.buf, q.off), (t.buf, t.off)
        for (ma, mi, go, ge, ss, fl) in sw_settings:
            r = batch.sw_align_batch(q, t, ma, mi, go, ge, score_size=ss, flag=fl)
            if use_ref:
                ref = O.ref_sw(qc, tc, ma, mi, go, ge, score_size=ss, flag=fl, nthreads=cores, csr=True)
            else:
                ref = O.oracle_sw(qc, tc, ma, mi, go, ge, score_size=ss, flag=fl, csr=True)
            ok = ref.status == 0   # score_size=0 overflow etc. are reported by both sides
            bad = cmp_sw(r, ref) & ok | ((r.status != 0) != (ref.status != 0))
            total += len(q); bad_total += int(bad.sum())
            print("SW  %-14s (%d,%d,%d,%d) size=%d flag=%d  pairs=%d  diffs=%d" % (name, ma, mi, go, ge, ss, fl, len(q), int(bad.sum())), flush=True)
            if a.forms:
                # the same pairs through the pre-packed form, and a sample through the one-launch path (batches of <= 64 pairs)
                rp = batch.sw_align_packed(batch.pack_seqs(q), batch.pack_seqs(t), ma, mi, go, ge, score_size=ss, flag=fl)
                badp = cmp_sw(rp, ref) & ok | ((rp.status != 0) != (ref.status != 0))
                nsm, bads, pos = min(len(q), 4096), 0, 0
                while pos < nsm:
                    k = int(rng.integers(1, 65))
                    idx = np.arange(pos, min(pos + k, nsm))
                    rs = batch.sw_align_batch(q.take(idx), t.take(idx), ma, mi, go, ge, score_size=ss, flag=fl)
                    sub_ok = ok[idx]
                    same = (rs.fixed[:, :7] == ref.fixed[idx, :7]).all(axis=1) & (np.diff(rs.cigar_off) == np.diff(ref.cigar_off)[idx])
                    for j, i in enumerate(idx):
                        if same[j] and sub_ok[j] and not np.array_equal(rs.cigar_of(j), ref.cigar_of(i)):
                            same[j] = False
                    bads += int((~same & sub_ok).sum()) + int(((rs.status != 0) != (ref.status[idx] != 0)).sum())
                    pos += k
                total += len(q) + nsm; bad_total += int(badp.sum()) + bads
                print("SW  %-14s   ... pre-packed form: pairs=%d diffs=%d; one-launch path: pairs=%d diffs=%d" % (name, len(q), int(badp.sum()), nsm, bads), flush=True)
        for (mode, task, k) in ed_settings:
            r = batch.ed_align_batch(q, t, mode=mode, task=task, k=-1 if k is None else k)
            if use_ref:
                ref = O.ref_ed(qc, tc, mode, task, k, nthreads=cores, csr=True)
            else:
                ref = O.oracle_ed(qc, tc, mode, task, k, csr=True)
            bad = cmp_ed(r, ref)
            if task == "path":
                bad |= cmp_path(r, ref)
                ops_total += int(ref.path_len.sum())
            total += len(q); bad_total += int(bad.sum())
            print("ED  %-14s %s/%s k=%s  pairs=%d  diffs=%d" % (name, mode, task, k, len(q), int(bad.sum())), flush=True)
    print("TOTAL alignments compared: %d (edit operations of task 'path': %d), differences: %d, %.0f s" % (total, ops_total, bad_total, time.time() - t0))
    sys.exit(1 if bad_total else 0)


if __name__ == "__main__":
    main()
