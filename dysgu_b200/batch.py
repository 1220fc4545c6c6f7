"""Batched alignment API over the C ABI: numpy CSR batches in, numpy results out.

This is the call a throughput-minded caller makes (SURVEY.md 8b "new batched ABI"); the drop-in
classes in dysgu_b200.scikitbio / dysgu_b200.edlib are thin single-pair views over it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib
from . import server as _server
from .workloads import SeqBatch

SW_FIELDS = ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len")
ED_MODES = {"NW": 0, "SHW": 1, "HW": 2}
ED_TASKS = {"distance": 0, "locations": 1, "path": 2}
ED_NO_START = -(2 ** 31)

PAIR_STATUS = {0: "ok", 1: "empty sequence", 2: "8-bit score overflow with score_size=0", 3: "score does not fit int16",
               4: "unsupported", 5: "device scratch exhausted"}


def default_device() -> int:
    return int(os.environ.get("DB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def pinned_copy(b: SeqBatch) -> SeqBatch:
    """Copy a batch into page-locked host memory (db200_host_alloc) so that the host entry points copy asynchronously."""
    L = _lib.lib()

    def pin(a):
        n = max(int(a.nbytes), 1)
        ptr = L.db200_host_alloc(n)
        if not ptr:
            raise _lib.AlignEngineError("pinned allocation failed")
        out = np.frombuffer((C.c_uint8 * n).from_address(ptr), dtype=a.dtype, count=a.size)
        out[:] = a
        return out
    return SeqBatch.__new_pinned__(pin(b.buf), pin(b.off))


def _as_batch(x) -> SeqBatch:
    return x if isinstance(x, SeqBatch) else SeqBatch.from_list(x)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def _csr_text(call):
    """Run one of the db200_*_text formatters (sizes first, then the text) -> (bytes, int64 offsets[n+1])."""
    rc, off = call(None, 0)
    _lib.check(rc, "text formatting")
    buf = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
    rc, off = call(buf, int(off[-1]))
    _lib.check(rc, "text formatting")
    return buf.tobytes()[:int(off[-1])], off


def _split_text(text: bytes, off) -> list:
    s = text.decode("ascii")
    return [s[off[i]:off[i + 1]] for i in range(len(off) - 1)]


class SwResults:
    """Results of a Smith-Waterman batch (fields follow s_align, ssw.h:38-48)."""

    def __init__(self, fixed, cigar, cigar_off, status):
        self.fixed = fixed          # int32 [n, 8]
        self.cigar = cigar          # uint32 [total], BAM encoding
        self.cigar_off = cigar_off  # int64 [n+1]
        self.status = status        # int32 [n]

    def __len__(self):
        return self.fixed.shape[0]

    def __getattr__(self, name):
        if name in SW_FIELDS:
            return self.fixed[:, SW_FIELDS.index(name)]
        raise AttributeError(name)

    def cigar_of(self, i):
        return self.cigar[self.cigar_off[i]:self.cigar_off[i + 1]]

    def cigar_str(self, i):
        return "".join("%d%s" % (c >> 4, "MID"[c & 0xF]) for c in self.cigar_of(i))

    def as_tuple(self, i):
        return tuple(int(x) for x in self.fixed[i, :7]) + (self.cigar_str(i),)

    def cigar_strings(self) -> list:
        """AlignmentStructure.cigar of every pair (_ssw_wrapper.pyx:240-275), formatted by one C call (db200_sw_cigar_text)."""
        n = len(self)
        L = _lib.lib()
        off = np.zeros(n + 1, dtype=np.int64)
        cig = self.cigar if self.cigar.size else np.zeros(1, np.uint32)
        coff = np.ascontiguousarray(self.cigar_off, dtype=np.int64)

        def call(buf, cap):
            return L.db200_sw_cigar_text(_ptr(cig), _ptr(coff), n, _ptr(buf), cap, _ptr(off)), off
        return _split_text(*_csr_text(call))

    def aligned_strings(self, queries, targets):
        """(aligned_query_sequence, aligned_target_sequence) of every pair with zero-based coordinates
        (_ssw_wrapper.pyx:302-364), formatted by two C calls (db200_sw_aligned_text) over the batch the results came from."""
        n = len(self)
        L = _lib.lib()
        cig = self.cigar if self.cigar.size else np.zeros(1, np.uint32)
        coff = np.ascontiguousarray(self.cigar_off, dtype=np.int64)
        fixed = np.ascontiguousarray(self.fixed, dtype=np.int32)
        out = []
        for which, seqs in ((0, _as_batch(queries)), (1, _as_batch(targets))):
            if len(seqs) != n:
                raise ValueError("sequences and results differ in length")
            off = np.zeros(n + 1, dtype=np.int64)
            sbuf = seqs.buf if seqs.buf.size else np.zeros(1, np.uint8)
            soff = np.ascontiguousarray(seqs.off, dtype=np.int64)

            def call(buf, cap, which=which, sbuf=sbuf, soff=soff, off=off):
                return L.db200_sw_aligned_text(which, _ptr(sbuf), _ptr(soff), _ptr(fixed), _ptr(cig), _ptr(coff), n, _ptr(buf), cap,
                                               _ptr(off)), off
            out.append(_split_text(*_csr_text(call)))
        return out[0], out[1]

    def cigar_stats(self):
        """Per pair, straight from the BAM-encoded cigar arrays (no string round trip): int64 [n, 3] =
        (bases in M runs, bases in I/D runs, longest I/D run) - the quantities merge_svs.matches_with_max_gap
        parses out of the cigar string (merge_svs.pyx:196-222).  Vectorised over the whole batch."""
        n = len(self)
        out = np.zeros((n, 3), dtype=np.int64)
        if self.cigar.size == 0:
            return out
        lens = (self.cigar >> 4).astype(np.int64)
        is_m = (self.cigar & 0xF) == 0
        starts = self.cigar_off[:-1]
        nonempty = self.cigar_off[1:] > starts
        idx = starts[nonempty]
        out[nonempty, 0] = np.add.reduceat(np.where(is_m, lens, 0), idx)
        out[nonempty, 1] = np.add.reduceat(np.where(is_m, 0, lens), idx)
        out[nonempty, 2] = np.maximum.reduceat(np.where(is_m, 0, lens), idx)
        return out


def sw_align_batch(queries, targets, match=2, mismatch=-3, gap_open=5, gap_extend=2, score_size=2, flag=1, filters=0,
                   filterd=0, mask_len=None, matrix=None, device=None, want_cigar=True, devices=None) -> SwResults:
    """Align queries[i] (profiled sequence) against targets[i]; semantics of ssw_init + ssw_align per pair.
    `devices` (a list of device ids) spreads the batch over several GPUs through db200_sw_batch_host_multi: contiguous
    ranges of equal work, one host thread per device, results in pair order (`SwResults.pairs_per_device`)."""
    q, t = _as_batch(queries), _as_batch(targets)
    n = len(q)
    if len(t) != n:
        raise ValueError("queries and targets differ in length")
    if _server.should_route():   # forked after the parent initialised CUDA: the parent's engine runs the batch
        return SwResults(*_server.remote("sw", (q, t), dict(match=match, mismatch=mismatch, gap_open=gap_open, gap_extend=gap_extend,
                                                            score_size=score_size, flag=flag, filters=filters, filterd=filterd,
                                                            mask_len=mask_len, matrix=matrix, want_cigar=want_cigar, devices=devices)))
    _server.ensure_started()
    L = _lib.lib()
    prm = _lib.SwParams()
    prm.match, prm.mismatch, prm.gap_open, prm.gap_extend = match, mismatch, gap_open, gap_extend
    prm.score_size, prm.flag, prm.filters, prm.filterd = score_size, flag, filters, filterd
    mat_arr = None
    if matrix is not None:
        mat_arr = np.ascontiguousarray(matrix, dtype=np.int8).reshape(25)
        prm.mat = mat_arr.ctypes.data_as(C.POINTER(C.c_int8))
        prm.n = 5
    ml = None if mask_len is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mask_len, dtype=np.int32), (n,)))
    fixed = np.zeros((n, 8), dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    coff = np.zeros(n + 1, dtype=np.int64)
    cap = int(q.off[-1] + t.off[-1] + 2 * n + 16) if want_cigar else 0
    cigar = np.zeros(max(cap, 1), dtype=np.uint32)
    qbuf = q.buf if q.buf.size else np.zeros(1, np.uint8)
    tbuf = t.buf if t.buf.size else np.zeros(1, np.uint8)
    if devices is not None and len(devices) > 1:
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        per_dev = np.zeros(len(devs), dtype=np.int64)
        rc = L.db200_sw_batch_host_multi(_ptr(devs), len(devs), C.byref(prm), _ptr(qbuf), _ptr(q.off), _ptr(tbuf), _ptr(t.off), n, _ptr(ml),
                                         _ptr(fixed), _ptr(cigar) if want_cigar else None, cap, _ptr(coff), _ptr(status), _ptr(per_dev))
        _lib.check(rc, "db200_sw_batch_host_multi")
        res = SwResults(fixed, cigar[:coff[n]] if want_cigar else cigar[:0], coff, status)
        res.pairs_per_device = per_dev
        return res
    if devices is not None and len(devices) == 1:
        device = int(devices[0])
    rc = L.db200_sw_batch_host(default_device() if device is None else device, C.byref(prm), _ptr(qbuf), _ptr(q.off), _ptr(tbuf),
                               _ptr(t.off), n, _ptr(ml), _ptr(fixed), _ptr(cigar) if want_cigar else None, cap, _ptr(coff),
                               _ptr(status))
    _lib.check(rc, "db200_sw_batch_host")
    return SwResults(fixed, cigar[:coff[n]] if want_cigar else cigar[:0], coff, status)


class PackedSeqs:
    """One side of a batch in the engine's pool format (db200_sw_pack_host): uint32 units (4 words = 32 bases per unit),
    int64 unit offsets [n + 1], int32 lengths [n]."""

    def __init__(self, units, unit_off, lengths):
        self.units, self.unit_off, self.lengths = units, unit_off, lengths

    def __len__(self):
        return len(self.lengths)

    @property
    def nbytes(self):
        return int(self.unit_off[-1]) * 16


def pack_seqs(seqs, threads=0, pinned=False) -> PackedSeqs:
    """ASCII batch -> pool format on the host (encoding of _ssw_wrapper.pyx:60-68, 4 bits per base)."""
    b = _as_batch(seqs)
    L = _lib.lib()
    n = len(b)
    uoff = np.zeros(n + 1, dtype=np.int64)
    lens = np.zeros(max(n, 1), dtype=np.int32)
    _lib.check(L.db200_sw_pack_host(None, _ptr(b.off), None, n, None, 0, _ptr(uoff), _ptr(lens), 0), "db200_sw_pack_host (sizing)")
    nwords = max(int(uoff[-1]) * 4, 1)
    if pinned:
        ptr = L.db200_host_alloc(nwords * 4)
        if not ptr:
            raise _lib.AlignEngineError("pinned allocation failed")
        units = np.frombuffer((C.c_uint8 * (nwords * 4)).from_address(ptr), dtype=np.uint32, count=nwords)
    else:
        units = np.zeros(nwords, dtype=np.uint32)
    buf = b.buf if b.buf.size else np.zeros(1, np.uint8)
    _lib.check(L.db200_sw_pack_host(_ptr(buf), _ptr(b.off), None, n, _ptr(units), int(uoff[-1]), _ptr(uoff), _ptr(lens),
                                    threads or (os.cpu_count() or 1)), "db200_sw_pack_host")
    return PackedSeqs(units, uoff, lens[:n])


def _sw_outputs(n, cap):
    return (np.zeros((n, 8), dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(n + 1, dtype=np.int64), np.zeros(max(cap, 1), dtype=np.uint32))


def _sw_params(match, mismatch, gap_open, gap_extend, score_size, flag, filters, filterd, matrix):
    prm = _lib.SwParams()
    prm.match, prm.mismatch, prm.gap_open, prm.gap_extend = match, mismatch, gap_open, gap_extend
    prm.score_size, prm.flag, prm.filters, prm.filterd = score_size, flag, filters, filterd
    mat_arr = None
    if matrix is not None:
        mat_arr = np.ascontiguousarray(matrix, dtype=np.int8).reshape(25)
        prm.mat = mat_arr.ctypes.data_as(C.POINTER(C.c_int8))
        prm.n = 5
    return prm, mat_arr


def sw_align_packed(queries: PackedSeqs, targets: PackedSeqs, match=2, mismatch=-3, gap_open=5, gap_extend=2, score_size=2, flag=1,
                    filters=0, filterd=0, mask_len=None, matrix=None, device=None) -> SwResults:
    """sw_align_batch for batches that are already in pool format (db200_sw_batch_host_packed): half the PCIe bytes of ASCII."""
    n = len(queries)
    if len(targets) != n:
        raise ValueError("queries and targets differ in length")
    prm, keep = _sw_params(match, mismatch, gap_open, gap_extend, score_size, flag, filters, filterd, matrix)
    ml = None if mask_len is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mask_len, dtype=np.int32), (n,)))
    cap = int(queries.lengths.astype(np.int64).sum() + targets.lengths.astype(np.int64).sum() + 2 * n + 16)
    fixed, status, coff, cigar = _sw_outputs(n, cap)
    rc = _lib.lib().db200_sw_batch_host_packed(default_device() if device is None else device, C.byref(prm), _ptr(queries.units),
                                               _ptr(queries.unit_off), _ptr(queries.lengths), _ptr(targets.units), _ptr(targets.unit_off),
                                               _ptr(targets.lengths), n, _ptr(ml), _ptr(fixed), _ptr(cigar), cap, _ptr(coff), _ptr(status))
    _lib.check(rc, "db200_sw_batch_host_packed")
    return SwResults(fixed, cigar[:coff[n]], coff, status)


def sw_align_slices(qsrc, qstart, qlen, targets, match=2, mismatch=-3, gap_open=5, gap_extend=2, score_size=2, flag=1, filters=0,
                    filterd=0, mask_len=None, matrix=None, device=None, tsrc=None, tstart=None, tlen=None) -> SwResults:
    """Queries are slices qsrc[qstart[i] : qstart[i] + qlen[i]] of one shared buffer that is uploaded once (overlapping windows of
    a reference region, re_map.py:429-434); targets either a CSR batch (`targets`) or slices of `tsrc` as well
    (db200_sw_batch_host_slices)."""
    qsrc = np.ascontiguousarray(np.frombuffer(qsrc, dtype=np.uint8) if isinstance(qsrc, (bytes, bytearray)) else qsrc, dtype=np.uint8)
    qstart = np.ascontiguousarray(qstart, dtype=np.int64)
    qlen = np.ascontiguousarray(qlen, dtype=np.int32)
    n = len(qlen)
    if tsrc is None:
        t = _as_batch(targets)
        tb, to, tl, tbytes = (t.buf if t.buf.size else np.zeros(1, np.uint8)), t.off, None, int(t.off[-1])
        tsum = tbytes
    else:
        tb = np.ascontiguousarray(np.frombuffer(tsrc, dtype=np.uint8) if isinstance(tsrc, (bytes, bytearray)) else tsrc, dtype=np.uint8)
        to, tl, tbytes = np.ascontiguousarray(tstart, dtype=np.int64), np.ascontiguousarray(tlen, dtype=np.int32), int(tb.size)
        tsum = int(tl.astype(np.int64).sum())
    prm, keep = _sw_params(match, mismatch, gap_open, gap_extend, score_size, flag, filters, filterd, matrix)
    ml = None if mask_len is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mask_len, dtype=np.int32), (n,)))
    cap = int(qlen.astype(np.int64).sum() + tsum + 2 * n + 16)
    fixed, status, coff, cigar = _sw_outputs(n, cap)
    rc = _lib.lib().db200_sw_batch_host_slices(default_device() if device is None else device, C.byref(prm), _ptr(qsrc), int(qsrc.size),
                                               _ptr(qstart), _ptr(qlen), _ptr(tb), tbytes, _ptr(to), _ptr(tl), n, _ptr(ml), _ptr(fixed),
                                               _ptr(cigar), cap, _ptr(coff), _ptr(status))
    _lib.check(rc, "db200_sw_batch_host_slices")
    return SwResults(fixed, cigar[:coff[n]], coff, status)


class EdResults:
    """Results of an edit-distance batch (fields follow EdlibAlignResult, edlib.h:162-218)."""

    def __init__(self, fixed, loc, loc_off, path=None, path_off=None):
        self.fixed = fixed      # int32 [n,4]: status, editDistance, alphabetLength, numLocations
        self.loc = loc          # int32 [total,2]: (start, end); start == ED_NO_START -> None
        self.loc_off = loc_off
        self.path = path        # task "path": uint8 edit operations (0 match, 1 insertion, 2 deletion, 3 mismatch)
        self.path_off = path_off  # int64 [n+1]

    @property
    def path_len(self):
        return None if self.path_off is None else np.diff(self.path_off)

    def __len__(self):
        return self.fixed.shape[0]

    @property
    def edit_distance(self):
        return self.fixed[:, 1]

    def alignment(self, i):
        """Edit operations of pair i (EdlibAlignResult.alignment), None where the reference returns none."""
        if self.path is None or self.path_off[i + 1] <= self.path_off[i]:
            return None
        return self.path[self.path_off[i]:self.path_off[i + 1]]

    def cigar(self, i, extended=True):
        """edlibAlignmentToCigar of pair i (edlib.cpp:298-347; the wrapper uses the extended format, edlib.pyx:143-147)."""
        ops = self.alignment(i)
        return None if ops is None else ops_to_cigar(ops, extended)

    def cigar_strings(self, extended=True) -> list:
        """The cigar of every pair of a task "path" batch (None where the reference returns none), one C call
        (db200_ed_cigar_text)."""
        if self.path is None:
            return [None] * len(self)
        n = len(self)
        L = _lib.lib()
        off = np.zeros(n + 1, dtype=np.int64)
        path = self.path if self.path.size else np.zeros(1, np.uint8)
        poff = np.ascontiguousarray(self.path_off, dtype=np.int64)

        def call(buf, cap):
            return L.db200_ed_cigar_text(_ptr(path), _ptr(poff), n, 1 if extended else 0, _ptr(buf), cap, _ptr(off)), off
        return [s if s else None for s in _split_text(*_csr_text(call))]

    def as_dict(self, i):
        st, ed, al, nl = (int(x) for x in self.fixed[i])
        locs = [(None if int(s) == ED_NO_START else int(s), int(e)) for s, e in self.loc[self.loc_off[i]:self.loc_off[i + 1]]]
        d = {"status": st, "editDistance": ed, "alphabetLength": al, "locations": locs}
        if self.path is not None:
            d["cigar"] = self.cigar(i)
        return d


def ops_to_cigar(ops, extended=True) -> str:
    """Run-length text of an edit-operation array; standard format merges match and mismatch into M."""
    ops = np.asarray(ops, dtype=np.uint8)
    if ops.size == 0:
        return ""
    letters = np.frombuffer(b"=IDX" if extended else b"MIDM", dtype=np.uint8)[ops]
    cut = np.flatnonzero(letters[1:] != letters[:-1]) + 1
    starts = np.concatenate(([0], cut))
    ends = np.concatenate((cut, [ops.size]))
    return "".join("%d%s" % (e - b, chr(letters[b])) for b, e in zip(starts, ends))


def ed_align_batch(queries, targets, mode="NW", task="distance", k=-1, additional_equalities=None, device=None, devices=None) -> EdResults:
    """edlib.align semantics per pair (edlib.pyx:57-156): tasks 'distance', 'locations' and 'path'.  `devices`: as in
    sw_align_batch (db200_ed_batch_host_multi; tasks 'distance' and 'locations')."""
    q, t = _as_batch(queries), _as_batch(targets)
    n = len(q)
    if len(t) != n:
        raise ValueError("queries and targets differ in length")
    if _server.should_route():
        return EdResults(*_server.remote("ed", (q, t), dict(mode=mode, task=task, k=k, additional_equalities=additional_equalities, devices=devices)))
    _server.ensure_started()
    L = _lib.lib()
    if not hasattr(L, "db200_ed_batch_host"):
        raise _lib.AlignEngineError("this build of libdysgu_b200_align.so has no edit-distance entry points")
    prm = _lib.EdParams()
    prm.mode = ED_MODES[mode] if isinstance(mode, str) else int(mode)
    prm.task = ED_TASKS[task] if isinstance(task, str) else int(task)
    karr = None
    if np.ndim(k) == 0:
        prm.k = int(k)
    else:
        prm.k = -1
        karr = np.ascontiguousarray(k, dtype=np.int32)
    eq = None
    if additional_equalities:
        eq = b"".join(bytes([a if isinstance(a, int) else ord(a), b if isinstance(b, int) else ord(b)]) for a, b in additional_equalities)
        prm.eq_pairs = eq
        prm.n_eq_pairs = len(eq) // 2
    fixed = np.zeros((n, 4), dtype=np.int32)
    loff = np.zeros(n + 1, dtype=np.int64)
    cap = int(t.off[-1] + 2 * n + 16)
    loc = np.zeros((cap, 2), dtype=np.int32)
    qbuf = q.buf if q.buf.size else np.zeros(1, np.uint8)
    tbuf = t.buf if t.buf.size else np.zeros(1, np.uint8)
    if prm.task == ED_TASKS["path"]:
        pcap = int(q.off[-1] + t.off[-1]) + 1
        path = np.zeros(pcap, dtype=np.uint8)
        poff = np.zeros(n + 1, dtype=np.int64)
        rc = L.db200_ed_path_batch_host(default_device() if device is None else device, C.byref(prm), _ptr(qbuf), _ptr(q.off),
                                        _ptr(tbuf), _ptr(t.off), n, _ptr(karr), _ptr(fixed), _ptr(loc), cap, _ptr(loff), _ptr(path),
                                        pcap, _ptr(poff))
        _lib.check(rc, "db200_ed_path_batch_host")
        return EdResults(fixed, loc[:loff[n]], loff, path[:poff[n]], poff)
    if devices is not None and len(devices) > 1:
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        per_dev = np.zeros(len(devs), dtype=np.int64)
        cap += 16 * len(devs)
        loc = np.zeros((cap, 2), dtype=np.int32)
        rc = L.db200_ed_batch_host_multi(_ptr(devs), len(devs), C.byref(prm), _ptr(qbuf), _ptr(q.off), _ptr(tbuf), _ptr(t.off), n, _ptr(karr),
                                         _ptr(fixed), _ptr(loc), cap, _ptr(loff), _ptr(per_dev))
        _lib.check(rc, "db200_ed_batch_host_multi")
        res = EdResults(fixed, loc[:loff[n]], loff)
        res.pairs_per_device = per_dev
        return res
    if devices is not None and len(devices) == 1:
        device = int(devices[0])
    rc = L.db200_ed_batch_host(default_device() if device is None else device, C.byref(prm), _ptr(qbuf), _ptr(q.off), _ptr(tbuf),
                               _ptr(t.off), n, _ptr(karr), _ptr(fixed), _ptr(loc), cap, _ptr(loff))
    _lib.check(rc, "db200_ed_batch_host")
    return EdResults(fixed, loc[:loff[n]], loff)
