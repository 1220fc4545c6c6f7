"""Minimizer sketches of clip sequences on the GPU.

`minimizers_batch(seqs, k=16, m=7)` returns, for every sequence, the set dysgu's `sliding_window_minimum(k, m, s, found)`
(graph.pyx:59-84; called per soft clip by ClipScoper._add_m_find_candidates, graph.pyx:146, with k = 16 and m = 7 from
construct_graph, graph.pyx:1467) leaves in `found`: the minimum of every window of k consecutive m-mer hashes plus the
first and the last m-mer hash, hashes = XXHash64(m-mer, seed 42) read as signed 64-bit integers (map_set_utils.pxd:28-29).
The values come back sorted ascending, CSR over the batch (the reference holds them in an unordered set).
Computed by csrc/minimizer.cu through db200_minimizers_batch_host; there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import batch as _batch


def _bind(L):
    if not getattr(L, "_kmers_bound", False):
        L.db200_minimizers_batch_host.restype = C.c_int
        L.db200_minimizers_batch_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p,
                                                  C.c_int64, C.c_void_p]
        L.db200_xxh64.restype = C.c_uint64
        L.db200_xxh64.argtypes = [C.c_char_p, C.c_int32, C.c_uint64]
        L._kmers_bound = True
    return L


def xxh64(data: bytes, seed: int = 42) -> int:
    """XXH64 of `data` (host side of the hash the kernel applies to every m-mer)."""
    return int(_bind(_lib.lib()).db200_xxh64(data, len(data), seed))


def minimizers_batch(seqs, k: int = 16, m: int = 7, seed: int = 42, device=None):
    """-> (int64 values sorted per sequence, int64 offsets[n + 1])."""
    b = _batch._as_batch(seqs)
    n = len(b)
    L = _bind(_lib.lib())
    cap = int(b.off[-1]) + 16
    out = np.zeros(cap, dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    buf = b.buf if b.buf.size else np.zeros(1, np.uint8)
    rc = L.db200_minimizers_batch_host(_batch.default_device() if device is None else device, buf.ctypes.data, b.off.ctypes.data, n, k, m, seed,
                                       out.ctypes.data, cap, off.ctypes.data)
    _lib.check(rc, "db200_minimizers_batch_host")
    return out[:off[n]], off


def minimizer_sets(seqs, k: int = 16, m: int = 7, seed: int = 42, device=None):
    """The same as Python sets, one per sequence (what ClipScoper iterates over)."""
    vals, off = minimizers_batch(seqs, k, m, seed, device)
    return [set(int(x) for x in vals[off[i]:off[i + 1]]) for i in range(len(off) - 1)]
