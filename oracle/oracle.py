"""TEST INFRASTRUCTURE ONLY.

ctypes bindings for (a) the CPU oracle restatement (oracle/_build/liboracle.so, built from
sw_oracle.c / ed_oracle.c) and (b) the reference's own compiled cores behind the batch driver
(oracle/_build/librefdriver.so -> oracle/_ref/libssw_ref.so, libedlib_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (dysgu_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF = os.path.join(HERE, "_ref")

INT32_MIN = -(2 ** 31)
MODES = {"NW": 0, "SHW": 1, "HW": 2}
TASKS = {"distance": 0, "locations": 1, "path": 2}


def build(verbose: bool = False) -> None:
    """Compile the oracle restatement + driver (and, where the reference checkout exists, the
    reference cores).  Building the checker is not using it."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "CC=/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "CC=gcc"], stdout=out)
    subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")], stdout=out)


def have_ref() -> bool:
    return os.path.exists(os.path.join(REF, "libssw_ref.so")) and os.path.exists(os.path.join(REF, "libedlib_ref.so"))


_oracle = None
_driver = None


def _p(arr, typ):
    return arr.ctypes.data_as(C.POINTER(typ))


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(BUILD, "liboracle.so")
        if not os.path.exists(path):
            build()
        _oracle = C.CDLL(path)
        _oracle.sw_oracle_batch_nt.restype = C.c_int
        _oracle.ed_oracle_batch.restype = C.c_int
        _oracle.ed_path_oracle_batch.restype = C.c_int
        _oracle.oracle_xxh64.restype = C.c_uint64
        _oracle.oracle_xxh64.argtypes = [C.c_char_p, C.c_int64, C.c_uint64]
        _oracle.minimizer_oracle_batch.restype = C.c_int
    return _oracle


_xxh_ref = None


def ref_xxh_lib():
    """The reference's own XXHash64 (oracle/_ref/libxxh_ref.so, compiled from dysgu/include/xxhash64.h by build_ref.sh)."""
    global _xxh_ref
    if _xxh_ref is None:
        _xxh_ref = C.CDLL(os.path.join(REF, "libxxh_ref.so"))
        _xxh_ref.ref_xxh64.restype = C.c_uint64
        _xxh_ref.ref_xxh64.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64]
    return _xxh_ref


def oracle_minimizers(seqs, k=16, m=7, seed=42, use_reference_hash=False, csr=False):
    """graph.pyx:59-84 for every sequence: (int64 sorted values, int64 offsets[n + 1]).  use_reference_hash: hash with the
    reference's compiled xxhash64.h instead of the restated XXH64 (pinning / golden generation)."""
    buf, off = seqs if csr else to_csr(seqs)
    n = len(off) - 1
    cap = int(off[-1]) + 2 * n + 16
    out = np.zeros(cap, dtype=np.int64)
    ooff = np.zeros(n + 1, dtype=np.int64)
    fn = C.cast(ref_xxh_lib().ref_xxh64, C.c_void_p) if use_reference_hash else None
    rc = oracle_lib().minimizer_oracle_batch(_p(buf, C.c_uint8), _p(off, C.c_int64), C.c_int64(n), C.c_int(k), C.c_int(m), C.c_uint64(seed), fn,
                                             _p(out, C.c_int64), C.c_int64(cap), _p(ooff, C.c_int64))
    if rc != 0:
        raise RuntimeError("minimizer_oracle_batch failed")
    return out[:ooff[n]], ooff


def driver_lib():
    global _driver
    if _driver is None:
        path = os.path.join(BUILD, "librefdriver.so")
        if not os.path.exists(path):
            build()
        if not have_ref():
            raise RuntimeError("oracle/_ref is missing: run oracle/build_ref.sh where the reference checkout exists")
        _driver = C.CDLL(path)
        _driver.ref_sw_batch_nt.restype = C.c_double
        _driver.ref_ed_batch.restype = C.c_double
        _driver.ref_ed_path_batch.restype = C.c_double
        rc = _driver.ref_driver_load(os.path.join(REF, "libssw_ref.so").encode(), os.path.join(REF, "libedlib_ref.so").encode())
        if rc != 0:
            raise RuntimeError("ref_driver_load failed: %d" % rc)
    return _driver


def to_csr(seqs: Sequence[bytes]):
    """list of bytes/str -> (uint8 buffer, int64 offsets[n+1])."""
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if off[-1] > 0 else np.zeros(1, dtype=np.uint8)
    return buf, off


def default_mask_len(qlens: np.ndarray) -> np.ndarray:
    """reference: _ssw_wrapper.pyx:583-588 (mask_auto): max(int(len/2), 15)."""
    return np.maximum(qlens // 2, 15).astype(np.int32)


class SwBatchResult:
    FIELDS = ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len")

    def __init__(self, fixed, cigar, cigar_off, status, seconds=None):
        self.fixed = fixed          # int32 [n, 8]
        self.cigar = cigar          # uint32 [total]
        self.cigar_off = cigar_off  # int64 [n+1]
        self.status = status        # int32 [n]
        self.seconds = seconds

    def cigar_of(self, i):
        return self.cigar[self.cigar_off[i]:self.cigar_off[i + 1]]

    def cigar_str(self, i):
        return "".join("%d%s" % (c >> 4, "MID"[c & 0xF]) for c in self.cigar_of(i))

    def as_tuple(self, i):
        f = self.fixed[i]
        return tuple(int(x) for x in f[:7]) + (self.cigar_str(i),)


def _sw_batch(fn, qcsr, tcsr, match, mismatch, gap_open, gap_extend, score_size, flag, filters, filterd, mask_len,
              want_cigar, nthreads=None):
    qbuf, qoff = qcsr
    tbuf, toff = tcsr
    n = len(qoff) - 1
    if mask_len is None:
        mask_len = default_mask_len(np.diff(qoff))
    mask_len = np.ascontiguousarray(mask_len, dtype=np.int32)
    fixed = np.zeros((n, 8), dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    cap = int((np.diff(qoff) + np.diff(toff)).sum()) + 2 * n + 16 if want_cigar else 0
    cigar = np.zeros(max(cap, 1), dtype=np.uint32)
    coff = np.zeros(n + 1, dtype=np.int64)
    args = [_p(qbuf, C.c_char), _p(qoff, C.c_int64), _p(tbuf, C.c_char), _p(toff, C.c_int64), C.c_int64(n),
            C.c_int(match), C.c_int(mismatch), C.c_int(gap_open), C.c_int(gap_extend), C.c_int(score_size),
            C.c_int(flag), C.c_int(filters), C.c_int(filterd), _p(mask_len, C.c_int32)]
    if nthreads is not None:
        args.append(C.c_int(nthreads))
    args += [_p(fixed, C.c_int32), _p(cigar, C.c_uint32) if want_cigar else None, C.c_int64(cap), _p(coff, C.c_int64),
             _p(status, C.c_int32)]
    rc = fn(*args)
    if rc < 0:
        raise RuntimeError("batch call failed: %r" % rc)
    return SwBatchResult(fixed, cigar[:coff[n]], coff, status, seconds=float(rc) if nthreads is not None else None)


def oracle_sw(queries, targets, match=2, mismatch=-3, gap_open=5, gap_extend=2, score_size=2, flag=1, filters=0,
              filterd=0, mask_len=None, csr=False):
    qcsr = queries if csr else to_csr(queries)
    tcsr = targets if csr else to_csr(targets)
    return _sw_batch(oracle_lib().sw_oracle_batch_nt, qcsr, tcsr, match, mismatch, gap_open, gap_extend, score_size,
                     flag, filters, filterd, mask_len, True)


def ref_sw(queries, targets, match=2, mismatch=-3, gap_open=5, gap_extend=2, score_size=2, flag=1, filters=0,
           filterd=0, mask_len=None, nthreads=1, csr=False, want_cigar=True):
    qcsr = queries if csr else to_csr(queries)
    tcsr = targets if csr else to_csr(targets)
    return _sw_batch(driver_lib().ref_sw_batch_nt, qcsr, tcsr, match, mismatch, gap_open, gap_extend, score_size,
                     flag, filters, filterd, mask_len, want_cigar, nthreads=nthreads)


class EdBatchResult:
    def __init__(self, fixed, loc, loc_off, seconds=None, path=None, path_off=None, path_len=None):
        self.fixed = fixed      # int32 [n,4]: status, editDistance, alphabetLength, numLocations
        self.loc = loc          # int32 [total,2]: start (INT32_MIN = None), end
        self.loc_off = loc_off
        self.seconds = seconds
        self.path = path        # task path: uint8 operations, pair i at path_off[i] .. path_off[i] + path_len[i]
        self.path_off = path_off
        self.path_len = path_len

    def as_dict(self, i):
        st, ed, al, nl = (int(x) for x in self.fixed[i])
        locs = [(None if int(s) == INT32_MIN else int(s), None if int(e) == INT32_MIN else int(e))
                for s, e in self.loc[self.loc_off[i]:self.loc_off[i + 1]]]
        return {"status": st, "editDistance": ed, "alphabetLength": al, "locations": locs}


def _ed_batch(fn, qcsr, tcsr, mode, task, k, nthreads=None, want_loc=True, want_path=False):
    qbuf, qoff = qcsr
    tbuf, toff = tcsr
    n = len(qoff) - 1
    fixed = np.zeros((n, 4), dtype=np.int32)
    cap = int(np.diff(toff).sum()) + 2 * n + 16 if want_loc else 0
    loc = np.zeros((max(cap, 1), 2), dtype=np.int32)
    loff = np.zeros(n + 1, dtype=np.int64)
    karr = None
    if k is not None:
        karr = np.ascontiguousarray(np.broadcast_to(np.asarray(k, dtype=np.int32), (n,)))
    args = [_p(qbuf, C.c_char), _p(qoff, C.c_int64), _p(tbuf, C.c_char), _p(toff, C.c_int64), C.c_int64(n),
            C.c_int32(MODES[mode] if isinstance(mode, str) else mode),
            C.c_int32(TASKS[task] if isinstance(task, str) else task),
            _p(karr, C.c_int32) if karr is not None else None]
    if nthreads is not None:
        args.append(C.c_int(nthreads))
    args += [_p(fixed, C.c_int32), _p(loc, C.c_int32) if want_loc else None, C.c_int64(cap), _p(loff, C.c_int64)]
    path = plen = poff = None
    if want_path:
        poff = (np.asarray(qoff) - qoff[0] + np.asarray(toff) - toff[0])[:n]
        path = np.zeros(int(qoff[n] - qoff[0] + toff[n] - toff[0]) + 1, dtype=np.uint8)
        plen = np.zeros(max(n, 1), dtype=np.int32)
        args += [_p(path, C.c_uint8), _p(plen, C.c_int32)]
    rc = fn(*args)
    if rc < 0:
        raise RuntimeError("batch call failed: %r" % rc)
    return EdBatchResult(fixed, loc[:loff[n]], loff, seconds=float(rc) if nthreads is not None else None, path=path, path_off=poff,
                         path_len=None if plen is None else plen[:n])


def oracle_ed(queries, targets, mode="NW", task="distance", k=None, csr=False):
    qcsr = queries if csr else to_csr(queries)
    tcsr = targets if csr else to_csr(targets)
    return _ed_batch(oracle_lib().ed_oracle_batch, qcsr, tcsr, mode, task, k)


def oracle_ed_path(queries, targets, mode="NW", k=None, additional_equalities=None, csr=False):
    """edlib.align(..., task="path") by the scalar restatement (ed_path_oracle.c): the fields of oracle_ed plus the edit
    operations, pair i at path[path_off[i] : path_off[i] + path_len[i]] (the reference driver's layout)."""
    qbuf, qoff = queries if csr else to_csr(queries)
    tbuf, toff = targets if csr else to_csr(targets)
    n = len(qoff) - 1
    fixed = np.zeros((n, 4), dtype=np.int32)
    cap = int(np.diff(toff).sum()) + 2 * n + 16
    loc = np.zeros((max(cap, 1), 2), dtype=np.int32)
    loff = np.zeros(n + 1, dtype=np.int64)
    karr = None if k is None else np.ascontiguousarray(np.broadcast_to(np.asarray(k, dtype=np.int32), (n,)))
    eq = b"".join(bytes([a if isinstance(a, int) else ord(a), b if isinstance(b, int) else ord(b)]) for a, b in (additional_equalities or []))
    poff = (np.asarray(qoff) - qoff[0] + np.asarray(toff) - toff[0])[:n]
    path = np.zeros(int(qoff[n] - qoff[0] + toff[n] - toff[0]) + 1, dtype=np.uint8)
    plen = np.zeros(max(n, 1), dtype=np.int32)
    rc = oracle_lib().ed_path_oracle_batch(_p(qbuf, C.c_char), _p(qoff, C.c_int64), _p(tbuf, C.c_char), _p(toff, C.c_int64), C.c_int64(n),
                                           C.c_int32(MODES[mode] if isinstance(mode, str) else mode),
                                           _p(karr, C.c_int32) if karr is not None else None, eq if eq else None, C.c_int32(len(eq) // 2),
                                           _p(fixed, C.c_int32), _p(loc, C.c_int32), C.c_int64(cap), _p(loff, C.c_int64),
                                           _p(path, C.c_uint8), _p(plen, C.c_int32))
    if rc != 0:
        raise RuntimeError("ed_path_oracle_batch failed: %r" % rc)
    return EdBatchResult(fixed, loc[:loff[n]], loff, path=path, path_off=poff, path_len=plen[:n])


def ref_ed(queries, targets, mode="NW", task="distance", k=None, nthreads=1, csr=False, want_loc=True):
    qcsr = queries if csr else to_csr(queries)
    tcsr = targets if csr else to_csr(targets)
    if task in ("path", TASKS["path"]):   # the reference's alignment arrays as well (EdlibAlignResult.alignment)
        return _ed_batch(driver_lib().ref_ed_path_batch, qcsr, tcsr, mode, task, k, nthreads=nthreads, want_loc=want_loc, want_path=True)
    return _ed_batch(driver_lib().ref_ed_batch, qcsr, tcsr, mode, task, k, nthreads=nthreads, want_loc=want_loc)
