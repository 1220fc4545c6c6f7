"""Synthetic alignment workloads of the shapes named in BASELINE.json.

Everything is produced as CSR batches: a uint8 buffer of ASCII bases plus int64 offsets, which
is exactly what the batched C-ABI (include/dysgu_b200_align.h) consumes.  Pure numpy; seeded.

Shapes mirror the reference call sites:
  * remap_windows     - StripedSmithWaterman(window 1-2 kb, 2,-8,6,1)(soft clip)   re_map.py:220-223
  * remap_chain       - edlib HW clip vs adjacent ref / clip vs <=1000 bp window     re_map.py:148, 209
  * contig_pairs      - check_contig_match SSW (2,-3,10,1), contigs <= 10 kb          consensus.pyx:977-989
  * insertion_linking - edit distance between insertion sequences 50 bp - 10 kb       (BASELINE configs 3/4)
"""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


class SeqBatch:
    """CSR batch of sequences: buf[off[i]:off[i+1]] is sequence i (ASCII bytes)."""

    __slots__ = ("buf", "off")

    def __init__(self, buf: np.ndarray, off: np.ndarray):
        self.buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self.off = np.ascontiguousarray(off, dtype=np.int64)

    @classmethod
    def __new_pinned__(cls, buf, off):
        """Wrap existing arrays without copying (used for page-locked buffers)."""
        self = cls.__new__(cls)
        self.buf, self.off = buf, off
        return self

    def __len__(self):
        return len(self.off) - 1

    @property
    def lengths(self):
        return np.diff(self.off)

    def __getitem__(self, i) -> bytes:
        return self.buf[self.off[i]:self.off[i + 1]].tobytes()

    def tolist(self):
        return [self[i] for i in range(len(self))]

    def take(self, idx) -> "SeqBatch":
        idx = np.asarray(idx, dtype=np.int64)
        lens = self.lengths[idx]
        off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        src = np.repeat(self.off[idx] - off[:-1], lens) + np.arange(off[-1], dtype=np.int64)
        return SeqBatch(self.buf[src] if off[-1] else np.zeros(0, np.uint8), off)

    @staticmethod
    def from_list(seqs) -> "SeqBatch":
        bs = [s.encode("latin-1") if isinstance(s, str) else (s if isinstance(s, bytes) else bytes(s)) for s in seqs]
        off = np.zeros(len(bs) + 1, dtype=np.int64)
        if bs:
            np.cumsum(np.fromiter(map(len, bs), dtype=np.int64, count=len(bs)), out=off[1:])
        # one join, no second copy: the engine only reads the buffer (the array is a read-only view of the joined bytes)
        return SeqBatch.__new_pinned__(np.frombuffer(b"".join(bs), dtype=np.uint8), off)


def _offsets(lens: np.ndarray) -> np.ndarray:
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return off


def random_dna(rng: np.random.Generator, lens: np.ndarray) -> SeqBatch:
    off = _offsets(np.asarray(lens, dtype=np.int64))
    codes = rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)
    return SeqBatch(BASES[codes], off)


def slices(src: SeqBatch, start: np.ndarray, length: np.ndarray) -> SeqBatch:
    """sequence i -> src[i][start[i]:start[i]+length[i]]"""
    length = np.asarray(length, dtype=np.int64)
    off = _offsets(length)
    base = np.repeat(src.off[:-1] + np.asarray(start, dtype=np.int64) - off[:-1], length)
    idx = base + np.arange(off[-1], dtype=np.int64)
    return SeqBatch(src.buf[idx], off)


def mutate(rng: np.random.Generator, seqs: SeqBatch, sub: float, ins: float, dele: float) -> SeqBatch:
    """Per-base substitution / insertion / deletion at the given rates (vectorised)."""
    n = len(seqs)
    total = int(seqs.off[-1])
    buf = seqs.buf.copy()
    if sub > 0 and total:
        m = rng.random(total) < sub
        k = int(m.sum())
        # substitute by a *different* base where the base is ACGT
        cur = np.searchsorted(BASES, buf[m])
        cur = np.where(cur < 4, cur, 0)
        buf[m] = BASES[(cur + rng.integers(1, 4, size=k)) % 4]
    if (ins <= 0 and dele <= 0) or total == 0:
        return SeqBatch(buf, seqs.off.copy())
    keep = (rng.random(total) >= dele).astype(np.int64)
    add = (rng.random(total) < ins).astype(np.int64)
    emit = keep + add
    pair = np.repeat(np.arange(n, dtype=np.int64), seqs.lengths)
    new_len = np.bincount(pair, weights=emit, minlength=n).astype(np.int64)
    src = np.repeat(np.arange(total, dtype=np.int64), emit)
    first = np.ones(len(src), dtype=bool)
    if len(src) > 1:
        first[1:] = src[1:] != src[:-1]
    is_ins = ~first | (keep[src] == 0)      # second emission, or the only emission of a deleted base
    out = buf[src]
    k = int(is_ins.sum())
    out[is_ins] = BASES[rng.integers(0, 4, size=k)]
    return SeqBatch(out, _offsets(new_len))


def lower(seqs: SeqBatch) -> SeqBatch:
    return SeqBatch(seqs.buf | np.uint8(0x20), seqs.off.copy())


# ---------------------------------------------------------------------------------------------
def remap_windows(n: int, seed: int = 20261021, window_lens=(1000, 1500, 2000), unrelated: float = 0.3,
                  tail: float = 0.05):
    """BASELINE config 2 (literal shape): query = 1-2 kb reference window, target = soft clip.

    clip length ~ U{15..150}, 5 % tail U{151..300}; 70 % are window slices with 1 % substitutions and
    0.2 % indels, lower-cased like dysgu's soft clips; 30 % unrelated.  Scoring (2,-8,6,1).
    Returns (queries, targets, params).
    """
    rng = np.random.default_rng(seed)
    wl = rng.choice(np.asarray(window_lens, dtype=np.int64), size=n)
    windows = random_dna(rng, wl)
    cl = rng.integers(15, 151, size=n)
    t = rng.random(n) < tail
    cl[t] = rng.integers(151, 301, size=int(t.sum()))
    start = (rng.random(n) * (wl - cl)).astype(np.int64)
    clips = mutate(rng, slices(windows, start, cl), 0.01, 0.002, 0.002)
    unrel = rng.random(n) < unrelated
    if unrel.any():
        # replace unrelated clips by random sequence of the same (mutated) length
        rnd = random_dna(rng, clips.lengths)
        mask = np.repeat(unrel, clips.lengths)
        clips.buf[mask] = rnd.buf[mask]
    params = dict(match=2, mismatch=-8, gap_open=6, gap_extend=1)
    return windows, lower(clips), params


def remap_chain(n: int, seed: int = 20261022):
    """BASELINE config 2a: the edlib legs of re_map.process_contig: clip (upper) vs 1000-bp window, HW/locations."""
    rng = np.random.default_rng(seed)
    wl = np.full(n, 1000, dtype=np.int64)
    windows = random_dna(rng, wl)
    cl = rng.integers(15, 151, size=n)
    t = rng.random(n) < 0.05
    cl[t] = rng.integers(151, 301, size=int(t.sum()))
    start = (rng.random(n) * (wl - cl)).astype(np.int64)
    clips = mutate(rng, slices(windows, start, cl), 0.01, 0.002, 0.002)
    unrel = rng.random(n) < 0.3
    if unrel.any():
        rnd = random_dna(rng, clips.lengths)
        mask = np.repeat(unrel, clips.lengths)
        clips.buf[mask] = rnd.buf[mask]
    return clips, windows, dict(mode="HW", task="locations")


def _loguniform(rng, lo, hi, n):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.int64)


def insertion_linking(n: int, seed: int = 20261023, platform: str = "hifi", lo: int = 50, hi: int = 10000):
    """BASELINE configs 3/4: insertion sequences 50 bp - 10 kb; 80 % related pairs, 20 % unrelated."""
    rng = np.random.default_rng(seed)
    L = _loguniform(rng, lo, hi, n)
    a = random_dna(rng, L)
    if platform == "hifi":
        b = mutate(rng, a, 0.002, 0.003, 0.003)
    else:  # ont r10
        b = mutate(rng, a, 0.015, 0.015, 0.02)
    unrel = rng.random(n) < 0.2
    if unrel.any():
        ratio = rng.uniform(0.7, 1.3, size=n)
        bl = np.where(unrel, np.maximum((L * ratio).astype(np.int64), 1), b.lengths)
        rnd = random_dna(rng, bl)
        # rebuild b: unrelated -> rnd, related -> b
        parts_off = _offsets(bl)
        out = np.empty(int(parts_off[-1]), dtype=np.uint8)
        rel_mask_dst = np.repeat(~unrel, bl)
        out[~rel_mask_dst] = rnd.buf[np.repeat(unrel, rnd.lengths)]
        out[rel_mask_dst] = b.buf[np.repeat(~unrel, b.lengths)]
        b = SeqBatch(out, parts_off)
    return a, b, dict(mode="NW", task="distance")


def contig_pairs(n: int, seed: int = 20261024, platform: str = "hifi", lo: int = 50, hi: int = 9000):
    """dysgu's long-read contig shape: flank500 + insertion + flank500, capped at 10 kb; check_contig_match scoring."""
    rng = np.random.default_rng(seed)
    ins_len = np.minimum(_loguniform(rng, lo, hi, n) + 1000, 10000)
    a = random_dna(rng, ins_len)
    if platform == "hifi":
        b = mutate(rng, a, 0.002, 0.003, 0.003)
    else:
        b = mutate(rng, a, 0.015, 0.015, 0.02)
    bl = np.minimum(b.lengths, 10000)
    b = slices(b, np.zeros(n, dtype=np.int64), bl)
    return a, b, dict(match=2, mismatch=-3, gap_open=10, gap_extend=1)


def pe_contig_pairs(n: int, seed: int = 20261025):
    """Paired-end contigs 100-500 bp for check_contig_match / get_badclip_metric shapes."""
    rng = np.random.default_rng(seed)
    L = rng.integers(100, 501, size=n)
    a = random_dna(rng, L)
    off = (rng.random(n) * (L // 3)).astype(np.int64)
    bl = np.maximum(L - off - (rng.random(n) * (L // 4)).astype(np.int64), 20)
    b = mutate(rng, slices(a, off, bl), 0.01, 0.003, 0.003)
    return a, b, dict(match=2, mismatch=-3, gap_open=10, gap_extend=1)
