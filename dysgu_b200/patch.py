"""Caller-side batching inside dysgu's own loops without editing or restating them.

dysgu's alignment loops are synchronous per pair - `StripedSmithWaterman(q, ...)(t)` / `edlib.align(q, t, ...)` - and their
control flow depends on the results (re_map.py:142-233: edlib vs adjacent reference -> edlib vs window -> SSW on the slice
the window search found; merge_svs.pyx:355-382: up to nine contig pairs per event pair with early exit;
post_call.py:50-79).  A GPU batch of one is slower than the CPU call it replaces, so the drop-in classes alone cannot speed
`dysgu call` up.  This module turns each of those loops into a handful of device batches while dysgu's OWN code still
makes every decision:

  1. **speculative passes.**  The original function runs on scratch copies of its mutable arguments with the two alignment
     entry points replaced by memo look-ups.  A look-up that misses records the request and returns a harmless
     placeholder (score 0 / no locations); at the end of the pass everything recorded is aligned in one device batch per
     parameter set and stored in the memo.  The pass repeats until one completes without a miss.  By induction the calls
     up to the first miss of a pass are exactly those of the real execution, so every pass extends the exact prefix; a
     pass without misses used real results only and therefore IS the real execution.  In practice re_map needs four passes
     (adjacent edlib + window edlib, SSW, second clip, check), post_call two, a merge job list at most ten.
  2. **the real run.**  The original function then runs once on the real arguments with the memo active: every call hits.
     A call that does not (never observed; counted in `Memo.misses`) is aligned directly - slower, never different.

Alignment results are pure functions of (parameters, query, target) and the engine is bit-exact against the reference's
cores, so the outputs - event attributes, graph edges, and with them the VCF - are those of the unpatched run; RNG state
(merge_svs.pyx:141-143 shuffles) is restored before every pass.  dysgu's multiprocessing.Pool for merge jobs
(merge_svs.pyx:396) is replaced by an in-process stand-in: the pool only existed to parallelise these alignments, and a
process forked after CUDA initialisation could not use the GPU anyway.

    import dysgu_b200, dysgu_b200.patch
    dysgu_b200.patch.install()          # before `dysgu call` runs; python -m dysgu_b200.run_dysgu call ... does exactly this

Nothing here imports dysgu at module level and nothing restates dysgu's logic; `install()` needs an importable dysgu.
"""
from __future__ import annotations

import contextlib
import copy
import functools
import logging
import random
import threading

import numpy as np

from . import batch as _batch
from .scikitbio._ssw_wrapper import AlignmentStructure, StripedSmithWaterman, _require_ascii

_tls = threading.local()


def active_memo():
    return getattr(_tls, "memo", None)


class Memo:
    """(parameters, query, target) -> result, filled by device batches.  `engine` provides sw_align_batch / ed_align_batch
    (default: dysgu_b200.batch, i.e. the GPU; tests inject the reference's wrappers to check the replay logic on CPU)."""

    def __init__(self, engine=None):
        self.engine = engine or _batch
        self.sw, self.ed = {}, {}
        self.pending_sw, self.pending_ed = {}, {}
        self.collecting = False
        self.hits = self.misses = self.batches = self.batched_pairs = self.passes = 0

    # ---- requests from the drop-in entry points ------------------------------------------------------------------
    def lookup_sw(self, key):
        r = self.sw.get(key)
        if r is not None:
            self.hits += 1
            return r
        if self.collecting:
            self.pending_sw.setdefault(key, None)
        return None

    def lookup_ed(self, key):
        r = self.ed.get(key)
        if r is not None:
            self.hits += 1
            return r
        if self.collecting:
            self.pending_ed.setdefault(key, None)
        return None

    # ---- one device batch per parameter set ---------------------------------------------------------------------------
    def flush(self) -> int:
        n = len(self.pending_sw) + len(self.pending_ed)
        groups = {}
        for key in self.pending_sw:
            groups.setdefault(key[0], []).append(key)
        for prm, keys in groups.items():
            match, mismatch, go, ge, score_size, flag, filters, filterd, mat = prm
            qs = [k[1] for k in keys]
            ts = [k[2] for k in keys]
            mask_arr = np.array([k[3] for k in keys], dtype=np.int32)   # the wrapper's per-query mask length (_ssw_wrapper.pyx:583-588)
            res = self.engine.sw_align_batch(qs, ts, match=match, mismatch=mismatch, gap_open=go, gap_extend=ge, score_size=score_size,
                                             flag=flag, filters=filters, filterd=filterd, mask_len=mask_arr,
                                             matrix=None if mat is None else np.array(mat, dtype=np.int8).reshape(5, 5))
            self.batches += 1
            self.batched_pairs += len(keys)
            for i, key in enumerate(keys):
                self.sw[key] = (tuple(int(x) for x in res.fixed[i, :7]), np.array(res.cigar_of(i), dtype=np.uint32), int(res.status[i]))
        groups = {}
        for key in self.pending_ed:
            groups.setdefault(key[0], []).append(key)
        for prm, keys in groups.items():
            mode, task, k, eqs = prm
            res = self.engine.ed_align_batch([kk[1] for kk in keys], [kk[2] for kk in keys], mode=mode, task=task, k=k,
                                             additional_equalities=list(eqs) if eqs else None)
            self.batches += 1
            self.batched_pairs += len(keys)
            for i, key in enumerate(keys):
                d = res.as_dict(i)
                self.ed[key] = d
        self.pending_sw, self.pending_ed = {}, {}
        return n

    @contextlib.contextmanager
    def active(self, collecting):
        outer, outer_c = active_memo(), self.collecting
        _tls.memo, self.collecting = self, collecting
        try:
            yield self
        finally:
            _tls.memo, self.collecting = outer, outer_c

    def stats(self):
        return {"passes": self.passes, "batches": self.batches, "batched_pairs": self.batched_pairs, "hits": self.hits, "misses": self.misses}


# ---- the two entry points dysgu's modules are pointed at ---------------------------------------------------------------
class MemoStripedSmithWaterman(StripedSmithWaterman):
    """StripedSmithWaterman (same constructor, _ssw_wrapper.pyx:539-554) whose calls are served from the active memo.
    Outside a replay it is the plain GPU drop-in."""

    def _key(self, target_bytes):
        mat = None if self.matrix is None else tuple(int(x) for x in np.asarray(self.matrix).reshape(-1))
        prm = (self.match_score, self.mismatch_score, self.gap_open_penalty, self.gap_extend_penalty, self.score_size, self.bit_flag,
               self.score_filter, self.distance_filter, mat)
        return (prm, self._query_bytes, target_bytes, int(self.mask_length))

    def _structure(self, fields, ops, target_sequence):
        if self.suppress_sequences:
            return AlignmentStructure(fields, ops, "", "", self.index_starts_at)
        return AlignmentStructure(fields, ops, self.read_sequence, target_sequence, self.index_starts_at)

    def __call__(self, target_sequence):
        memo = active_memo()
        if memo is None:
            return super().__call__(target_sequence)
        _require_ascii(target_sequence)
        key = self._key(target_sequence.encode("ascii"))
        hit = memo.lookup_sw(key)
        if hit is None:
            if memo.collecting:   # placeholder: no alignment (score 0, no cigar) - the pass is repeated with the real result
                return self._structure((0, 0, -1, -1, -1, 0, 0), np.zeros(0, np.uint32), target_sequence)
            memo.misses += 1
            return super().__call__(target_sequence)
        fields, ops, status = hit
        if status != 0:
            raise RuntimeError("alignment failed: %s" % _batch.PAIR_STATUS.get(status, status))
        return self._structure(fields, ops, target_sequence)


class MemoEdlib:
    """Stands in for the `edlib` module object the callers hold (`import dysgu.edlib as edlib`, re_map.py:6;
    `from dysgu.edlib import edlib`, filter_normals.py:20): `.align` with the reference's signature (edlib.pyx:57)."""

    @staticmethod
    def align(query, target, mode="NW", task="distance", k=-1, additionalEqualities=None):
        from .edlib import edlib as _ed
        memo = active_memo()
        if memo is None:
            return _ed.align(query, target, mode, task, k, additionalEqualities)
        qb, tb, eqs = _ed._to_bytes(query, target, additionalEqualities)
        if mode not in ("NW", "HW", "SHW"):
            mode = "NW"
        if task not in ("distance", "locations", "path"):
            task = "distance"
        if eqs is not None:
            eqs = tuple(((a.encode("utf-8")[0] if isinstance(a, str) else int(a)), (b.encode("utf-8")[0] if isinstance(b, str) else int(b)))
                        for a, b in eqs)
        key = ((mode, task, -1 if k is None else int(k), eqs), qb, tb)
        d = memo.lookup_ed(key)
        if d is None:
            if memo.collecting:
                return {"editDistance": -1, "alphabetLength": 0, "locations": [], "cigar": None}
            memo.misses += 1
            return _ed.align(query, target, mode, task, k, additionalEqualities)
        if d["status"] == 1:
            raise Exception("There was an error.")
        return {"editDistance": d["editDistance"], "alphabetLength": d["alphabetLength"], "locations": list(d["locations"]), "cigar": d.get("cigar")}

    @staticmethod
    def getNiceAlignment(alignResult, query, target, gapSymbol="-"):
        from .edlib import edlib as _ed
        return _ed.getNiceAlignment(alignResult, query, target, gapSymbol)


memo_edlib = MemoEdlib()
memo_edlib.edlib = memo_edlib   # `dysgu.edlib.edlib.align` spelling


# ---- speculative replay ----------------------------------------------------------------------------------------------------
class InlinePool:
    """multiprocessing.Pool stand-in for merge_svs.pyx:396: jobs stay in this process, where the memo and the engine live."""

    def __init__(self, processes=None, *a, **k):
        self.processes = processes

    def starmap(self, func, iterable, chunksize=None):
        return [func(*args) for args in iterable]

    def map(self, func, iterable, chunksize=None):
        return [func(x) for x in iterable]

    def close(self):
        pass

    def join(self):
        pass

    def terminate(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class CachingReference:
    """Proxy of the pysam.FastaFile handed to re_map.remap_soft_clips: the speculative passes fetch the same windows again."""

    def __init__(self, ref):
        self._ref, self._cache = ref, {}

    def fetch(self, *args, **kwargs):
        key = (args, tuple(sorted(kwargs.items())))
        if key not in self._cache:
            self._cache[key] = self._ref.fetch(*args, **kwargs)
        return self._cache[key]

    def __getattr__(self, name):
        return getattr(self._ref, name)


def replay(run_scratch, run_real, engine=None, max_passes=16, memo=None):
    """run_scratch(): the original code on scratch copies; run_real(): on the real arguments.  Returns (result, memo)."""
    memo = memo or Memo(engine)
    rng_state, np_state = random.getstate(), np.random.get_state()
    log_level = logging.root.manager.disable
    for _ in range(max_passes):
        memo.passes += 1
        random.setstate(rng_state)
        np.random.set_state(np_state)
        logging.disable(logging.CRITICAL)   # warnings belong to the real run
        try:
            with memo.active(collecting=True):
                run_scratch()
        except Exception:   # noqa: BLE001 - a placeholder result may send the original code somewhere odd; the pass is only a probe
            pass
        finally:
            logging.disable(log_level)
        if memo.flush() == 0:
            break
    random.setstate(rng_state)
    np.random.set_state(np_state)
    with memo.active(collecting=False):
        return run_real(), memo


last_stats = {}   # function name -> Memo.stats() of its most recent patched call (diagnostics, tests)


def _scratch_events(events):
    return [copy.copy(e) for e in events]


def _scratch_graph(G):
    return G.copy() if hasattr(G, "copy") and not isinstance(G, dict) else copy.deepcopy(G)


def batched(orig, name=None, scratch=None, engine=None):
    """Wrap `orig` so that it runs as speculative passes + one real run (module docstring).  `scratch(args, kwargs)`
    returns the (args, kwargs) of a speculative pass: copies of whatever the function mutates.  The wrapped function
    returns what `orig` returns; the statistics of its last call are in `last_stats[name]`."""
    name = name or getattr(orig, "__name__", "loop")

    @functools.wraps(orig)
    def wrapped(*args, **kwargs):
        def probe():
            a, k = scratch(args, kwargs) if scratch else (args, kwargs)
            return orig(*a, **k)
        result, memo = replay(probe, lambda: orig(*args, **kwargs), engine)
        last_stats[name] = memo.stats()
        return result
    wrapped._db200 = True
    return wrapped


def _wrap_remap(orig, engine=None):
    """re_map.remap_soft_clips(events, ref_genome, keep_unmapped=True, min_support=3): events are mutated, the reference
    is fetched per window group (re_map.py:429-434) - the passes share one caching proxy."""
    @functools.wraps(orig)
    def remap_soft_clips(events, ref_genome, *args, **kwargs):
        ref = CachingReference(ref_genome)
        return inner(events, ref, *args, **kwargs)
    inner = batched(orig, "remap_soft_clips", lambda a, k: ((_scratch_events(a[0]),) + tuple(a[1:]), k), engine)
    remap_soft_clips._db200 = True
    return remap_soft_clips


def _wrap_badclip(orig, engine=None):
    """post_call.get_badclip_metric(events, bad_clip_counter, bam, regions): events are mutated (post_call.py:46-79)."""
    return batched(orig, "get_badclip_metric", lambda a, k: ((_scratch_events(a[0]),) + tuple(a[1:]), k), engine)


def _wrap_enumerate_events(orig, engine=None):
    """merge_svs.enumerate_events(G, potential, ...): the graph is mutated, the events are only read (merge_svs.pyx:385-541)."""
    return batched(orig, "enumerate_events", lambda a, k: ((_scratch_graph(a[0]),) + tuple(a[1:]), k), engine)


def apply(modules: dict, engine=None):
    """Point the given dysgu modules at the memo entry points and wrap their loops.  `modules` maps the names
    'consensus', 're_map', 'post_call', 'merge_svs', 'filter_normals' to module objects (missing ones are skipped).
    Returns what was patched.  Substitution points: consensus.pyx:16, re_map.py:1,6, post_call.py:15,
    filter_normals.py:18,20; loops: re_map.py:366-481, post_call.py:27-81, merge_svs.pyx:385-541."""
    done = []
    for name in ("consensus", "re_map", "post_call", "filter_normals"):
        m = modules.get(name)
        if m is None:
            continue
        if hasattr(m, "StripedSmithWaterman"):
            m.StripedSmithWaterman = MemoStripedSmithWaterman
            done.append(name + ".StripedSmithWaterman")
        if hasattr(m, "edlib"):
            m.edlib = memo_edlib
            done.append(name + ".edlib")
    m = modules.get("re_map")
    if m is not None and hasattr(m, "remap_soft_clips") and not getattr(m.remap_soft_clips, "_db200", False):
        m.remap_soft_clips = _wrap_remap(m.remap_soft_clips, engine)
        done.append("re_map.remap_soft_clips")
    m = modules.get("post_call")
    if m is not None and hasattr(m, "get_badclip_metric") and not getattr(m.get_badclip_metric, "_db200", False):
        m.get_badclip_metric = _wrap_badclip(m.get_badclip_metric, engine)
        done.append("post_call.get_badclip_metric")
    m = modules.get("merge_svs")
    if m is not None and hasattr(m, "enumerate_events") and not getattr(m.enumerate_events, "_db200", False):
        m.enumerate_events = _wrap_enumerate_events(m.enumerate_events, engine)
        if hasattr(m, "Pool"):
            m.Pool = InlinePool
            done.append("merge_svs.Pool")
        done.append("merge_svs.enumerate_events")
    return done


def install(engine=None):
    """dysgu_b200.install() (module routing) + apply() on the installed dysgu package."""
    import importlib
    from . import install as _route
    _route()
    mods = {}
    for name in ("consensus", "re_map", "post_call", "merge_svs", "filter_normals"):
        try:
            mods[name] = importlib.import_module("dysgu." + name)
        except Exception as exc:   # noqa: BLE001
            logging.getLogger("dysgu_b200").warning("dysgu.%s not patched: %s", name, exc)
    return apply(mods, engine)
