"""Capture shim: record every alignment dysgu issues, with its result, at the module substitution points.

BASELINE configs[0] / SURVEY section 4: run `dysgu call` on the bundled test BAM with the reference's own CPU wrappers
under this shim, keep `(call site, parameters, query, target, result)` of every `StripedSmithWaterman(q, ...)(t)` and
`edlib.align(q, t, ...)`, then replay the file through the GPU engine (`scripts/replay_capture.py`): every field of every
pair must be equal.  The shim wraps whatever implementation a dysgu module currently holds - the reference's compiled
wrappers, or this package's drop-ins (then the capture documents what the GPU run computed).

Substitution points (module globals looked up at call time): `StripedSmithWaterman` in dysgu.consensus (consensus.pyx:16),
dysgu.re_map (re_map.py:1), dysgu.post_call (post_call.py:15), dysgu.filter_normals (filter_normals.py:18); `edlib` in
dysgu.re_map (re_map.py:6) and dysgu.filter_normals (filter_normals.py:20).

    from dysgu_b200 import capture
    session = capture.install("pairs.jsonl")        # imports the dysgu modules and wraps their bindings
    ... run dysgu (e.g. dysgu.main.cli(standalone_mode=False)) ...
    session.close()

File format: JSON lines, one alignment per line:
  {"kind": "sw", "site": "consensus", "params": {constructor kwargs}, "query": "...", "target": "...",
   "result": {"optimal_alignment_score": .., "suboptimal_alignment_score": .., "query_begin": .., "query_end": ..,
              "target_begin": .., "target_end_optimal": .., "target_end_suboptimal": .., "cigar": ".."}}
  {"kind": "ed", "site": "re_map", "params": {"mode": .., "task": .., "k": .., "additionalEqualities": ..},
   "query": "...", "target": "...", "result": {"editDistance": .., "alphabetLength": .., "locations": [[s, e], ..], "cigar": ..}}
"""
from __future__ import annotations

import gzip
import json
import threading

SW_RESULT_KEYS = ("optimal_alignment_score", "suboptimal_alignment_score", "query_begin", "query_end", "target_begin",
                  "target_end_optimal", "target_end_suboptimal", "cigar")
SW_CTOR_DEFAULTS = dict(gap_open_penalty=5, gap_extend_penalty=2, score_size=2, mask_length=15, mask_auto=True, score_only=False,
                        score_filter=None, distance_filter=None, override_skip_babp=False, protein=False, match_score=2,
                        mismatch_score=-3, substitution_matrix=None, suppress_sequences=False, zero_index=True)
_SW_ORDER = tuple(SW_CTOR_DEFAULTS)


class CaptureSession:
    def __init__(self, path=None):
        self.path = path
        self.records = []
        self._lock = threading.Lock()
        self._fh = None
        if path:
            self._fh = gzip.open(path, "wt") if str(path).endswith(".gz") else open(path, "w")
        self._undo = []

    def add(self, rec):
        from . import patch
        memo = patch.active_memo()
        if memo is not None and memo.collecting:
            return   # a speculative pass of dysgu_b200.patch: placeholders, not results
        with self._lock:
            if self._fh:
                self._fh.write(json.dumps(rec) + "\n")
            else:
                self.records.append(rec)

    def close(self):
        for obj, name, old in reversed(self._undo):
            setattr(obj, name, old)
        self._undo = []
        if self._fh:
            self._fh.close()
            self._fh = None

    # ---- wrappers ---------------------------------------------------------------------------------------------------
    def wrap_ssw(self, cls, site):
        """A class with the constructor / call surface of `cls` (_ssw_wrapper.pyx:539-554, 613-648) that records each call."""
        session = self

        class CapturedStripedSmithWaterman:
            def __init__(self, query_sequence, *args, **kwargs):
                params = dict(SW_CTOR_DEFAULTS)
                params.update(dict(zip(_SW_ORDER, args)))
                params.update(kwargs)
                self._params = {k: v for k, v in params.items() if k != "substitution_matrix" or v is not None}
                self._query = query_sequence
                self._inner = cls(query_sequence, *args, **kwargs)

            def __call__(self, target_sequence):
                a = self._inner(target_sequence)
                session.add({"kind": "sw", "site": site, "params": self._params, "query": self._query, "target": target_sequence,
                             "result": {k: a[k] for k in SW_RESULT_KEYS}})
                return a

            def __getattr__(self, name):
                return getattr(self._inner, name)

        CapturedStripedSmithWaterman.__name__ = getattr(cls, "__name__", "StripedSmithWaterman")
        return CapturedStripedSmithWaterman

    def wrap_edlib(self, mod, site):
        """An object with `.align` (edlib.pyx:57) and everything else of `mod`."""
        session = self

        class CapturedEdlib:
            def __getattr__(self, name):
                return getattr(mod, name)

            @staticmethod
            def align(query, target, mode="NW", task="distance", k=-1, additionalEqualities=None):
                r = mod.align(query, target, mode, task, k, additionalEqualities)
                def text(s):
                    return s if isinstance(s, str) else (s.decode("latin-1") if isinstance(s, (bytes, bytearray)) else list(s))
                session.add({"kind": "ed", "site": site,
                             "params": {"mode": mode, "task": task, "k": k,
                                        "additionalEqualities": None if additionalEqualities is None else [list(p) for p in additionalEqualities]},
                             "query": text(query), "target": text(target),
                             "result": {"editDistance": r["editDistance"], "alphabetLength": r["alphabetLength"],
                                        "locations": [list(x) for x in r["locations"]], "cigar": r.get("cigar")}})
                return r

        c = CapturedEdlib()
        c.edlib = c
        return c

    def attach(self, modules: dict):
        """Wrap the bindings of the given dysgu modules ({'consensus': module, ...}); undone by close()."""
        for site, m in modules.items():
            if m is None:
                continue
            if hasattr(m, "StripedSmithWaterman"):
                self._undo.append((m, "StripedSmithWaterman", m.StripedSmithWaterman))
                m.StripedSmithWaterman = self.wrap_ssw(m.StripedSmithWaterman, site)
            if hasattr(m, "edlib"):
                self._undo.append((m, "edlib", m.edlib))
                m.edlib = self.wrap_edlib(m.edlib, site)
        return self


def install(path=None) -> CaptureSession:
    """Import dysgu's alignment callers and wrap their bindings (needs an importable dysgu)."""
    import importlib
    mods = {}
    for name in ("consensus", "re_map", "post_call", "filter_normals"):
        try:
            mods[name] = importlib.import_module("dysgu." + name)
        except Exception:   # noqa: BLE001 - a module that cannot be imported cannot align either
            mods[name] = None
    return CaptureSession(path).attach(mods)


def load(path):
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rt") as f:
        return [json.loads(line) for line in f if line.strip()]
