"""Ticket queue over the C ABI (db200_queue_*: submit / flush / fetch) and the deferred mode of the drop-in classes.

dysgu's loops call `StripedSmithWaterman(q, ...)(t)` and `edlib.align(q, t, ...)` one pair at a time (SURVEY.md 8b,
8f-1).  Inside `with dysgu_b200.deferred():` those calls only *submit*: they return lazy results that resolve on
first use, and everything still pending is aligned in one device batch per parameter set - on the first access to any
lazy result or when the block ends.  Results are pure functions of the inputs, so deferring cannot change them:

    with dysgu_b200.deferred():
        alns = [StripedSmithWaterman(w, match_score=2, mismatch_score=-8, gap_open_penalty=6, gap_extend_penalty=1)(c)
                for w, c in candidates]           # nothing has run yet
    scores = [a.optimal_alignment_score for a in alns]   # one batch
"""
from __future__ import annotations

import contextlib
import ctypes as C
import threading
from collections.abc import Mapping
from typing import Optional

import numpy as np

from . import _lib
from . import batch as _batch


class SwResultC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in _batch.SW_FIELDS]


class EdResultC(C.Structure):
    _fields_ = [("status", C.c_int32), ("edit_distance", C.c_int32), ("alphabet_length", C.c_int32), ("num_locations", C.c_int32)]


def _bind(L):
    if getattr(L, "_queue_bound", False):
        return L
    L.db200_queue_create.restype = C.c_void_p
    L.db200_queue_create.argtypes = [C.c_int]
    L.db200_queue_destroy.argtypes = [C.c_void_p]
    L.db200_queue_submit_sw.restype = C.c_int64
    L.db200_queue_submit_sw.argtypes = [C.c_void_p, C.POINTER(_lib.SwParams), C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32]
    L.db200_queue_submit_ed.restype = C.c_int64
    L.db200_queue_submit_ed.argtypes = [C.c_void_p, C.POINTER(_lib.EdParams), C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
    L.db200_queue_pending.restype = C.c_int64
    L.db200_queue_pending.argtypes = [C.c_void_p]
    L.db200_queue_flush.argtypes = [C.c_void_p]
    L.db200_queue_reset.argtypes = [C.c_void_p]
    L.db200_queue_fetch_sw.argtypes = [C.c_void_p, C.c_int64, C.POINTER(SwResultC), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_int32)]
    L.db200_queue_fetch_ed.argtypes = [C.c_void_p, C.c_int64, C.POINTER(EdResultC), C.POINTER(C.POINTER(C.c_int32))]
    L.db200_queue_fetch_ed_path.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int32)]
    L._queue_bound = True
    return L


class AlignQueue:
    """submit_sw / submit_ed return tickets; flush() runs one device batch per parameter set; fetch_* read by ticket
    (and flush implicitly if the ticket is still pending)."""

    def __init__(self, device: Optional[int] = None):
        self._L = _bind(_lib.lib())
        self.device = _batch.default_device() if device is None else device
        self._q = self._L.db200_queue_create(self.device)
        if not self._q:
            raise _lib.AlignEngineError("db200_queue_create failed: %s" % self._L.db200_last_error().decode("utf-8", "replace"))
        self._keep = []   # parameter blocks referenced by pointer until the submit call returns

    def close(self):
        if self._q:
            self._L.db200_queue_destroy(self._q)
            self._q = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def pending(self) -> int:
        return int(self._L.db200_queue_pending(self._q))

    def submit_sw(self, query: bytes, target: bytes, match=2, mismatch=-3, gap_open=5, gap_extend=2, score_size=2, flag=1,
                  filters=0, filterd=0, mask_len=-1, matrix=None) -> int:
        prm = _lib.SwParams()
        prm.match, prm.mismatch, prm.gap_open, prm.gap_extend = match, mismatch, gap_open, gap_extend
        prm.score_size, prm.flag, prm.filters, prm.filterd = score_size, flag, filters, filterd
        mat = None
        if matrix is not None:
            mat = np.ascontiguousarray(matrix, dtype=np.int8).reshape(25)
            prm.mat = mat.ctypes.data_as(C.POINTER(C.c_int8))
            prm.n = 5
        t = int(self._L.db200_queue_submit_sw(self._q, C.byref(prm), query, len(query), target, len(target), int(mask_len)))
        if t < 0:
            _lib.check(int(t), "db200_queue_submit_sw")
        return t

    def submit_ed(self, query: bytes, target: bytes, mode="NW", task="distance", k=-1, additional_equalities=None) -> int:
        prm = _lib.EdParams()
        prm.mode = _batch.ED_MODES[mode] if isinstance(mode, str) else int(mode)
        prm.task = _batch.ED_TASKS[task] if isinstance(task, str) else int(task)
        prm.k = int(k)
        if additional_equalities:
            eq = b"".join(bytes([a if isinstance(a, int) else ord(a), b if isinstance(b, int) else ord(b)]) for a, b in additional_equalities)
            prm.eq_pairs = eq
            prm.n_eq_pairs = len(eq) // 2
        t = int(self._L.db200_queue_submit_ed(self._q, C.byref(prm), query, len(query), target, len(target)))
        if t < 0:
            _lib.check(int(t), "db200_queue_submit_ed")
        return t

    def flush(self):
        _lib.check(self._L.db200_queue_flush(self._q), "db200_queue_flush")

    def reset(self):
        _lib.check(self._L.db200_queue_reset(self._q), "db200_queue_reset")

    def fetch_sw(self, ticket: int):
        """-> (int32[8] fields in s_align order, uint32[] BAM-encoded cigar, DB200_PAIR_* status)"""
        r = SwResultC()
        cig = C.POINTER(C.c_uint32)()
        st = C.c_int32(0)
        _lib.check(self._L.db200_queue_fetch_sw(self._q, ticket, C.byref(r), C.byref(cig), C.byref(st)), "db200_queue_fetch_sw")
        fields = np.array([getattr(r, n) for n in _batch.SW_FIELDS], dtype=np.int32)
        ops = np.ctypeslib.as_array(cig, shape=(r.cigar_len,)).copy() if r.cigar_len > 0 and cig else np.zeros(0, np.uint32)
        return fields, ops, int(st.value)

    def fetch_ed(self, ticket: int) -> dict:
        r = EdResultC()
        loc = C.POINTER(C.c_int32)()
        _lib.check(self._L.db200_queue_fetch_ed(self._q, ticket, C.byref(r), C.byref(loc)), "db200_queue_fetch_ed")
        locs = []
        if r.num_locations > 0 and loc:
            a = np.ctypeslib.as_array(loc, shape=(2 * r.num_locations,))
            locs = [(None if int(a[2 * i]) == _batch.ED_NO_START else int(a[2 * i]), int(a[2 * i + 1])) for i in range(r.num_locations)]
        ops = C.POINTER(C.c_uint8)()
        nops = C.c_int32(0)
        _lib.check(self._L.db200_queue_fetch_ed_path(self._q, ticket, C.byref(ops), C.byref(nops)), "db200_queue_fetch_ed_path")
        d = {"status": int(r.status), "editDistance": int(r.edit_distance), "alphabetLength": int(r.alphabet_length), "locations": locs}
        if nops.value > 0 and ops:   # task "path" tickets with an alignment
            d["cigar"] = _batch.ops_to_cigar(np.ctypeslib.as_array(ops, shape=(nops.value,)))
        return d


# ---- deferred mode of the drop-in classes ---------------------------------------------------------------------------
_tls = threading.local()


def active_queue() -> Optional[AlignQueue]:
    return getattr(_tls, "queue", None)


@contextlib.contextmanager
def deferred(device: Optional[int] = None):
    """Inside the block StripedSmithWaterman(...)(...) and edlib.align(...) submit to one queue and return lazy results."""
    outer = active_queue()
    q = AlignQueue(device)
    _tls.queue = q
    try:
        yield q
        if q.pending:
            q.flush()
    finally:
        _tls.queue = outer
        # lazy results created in the block keep `q` alive until they have been resolved


class LazyEdlibResult(Mapping):
    """The dict edlib.align returns (edlib.pyx:131-156), resolved from its ticket on first use."""

    def __init__(self, queue: AlignQueue, ticket: int):
        self._queue, self._ticket, self._d = queue, ticket, None

    def _resolve(self):
        if self._d is None:
            d = self._queue.fetch_ed(self._ticket)
            if d["status"] == 1:
                raise Exception("There was an error.")
            self._d = {"editDistance": d["editDistance"], "alphabetLength": d["alphabetLength"], "locations": d["locations"], "cigar": d.get("cigar")}
            self._queue = None
        return self._d

    def __getitem__(self, key):
        return self._resolve()[key]

    def __iter__(self):
        return iter(self._resolve())

    def __len__(self):
        return len(self._resolve())

    def __repr__(self):
        return repr(self._resolve())
