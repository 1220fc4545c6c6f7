"""edlib.align with the reference's Python surface (dysgu/edlib/edlib.pyx:57-156), computed on the GPU."""
from __future__ import annotations

import re

from .. import batch as _batch
from .. import queue as _queue


def _to_bytes(query, target, additional_equalities):
    """str/bytes pass through; other iterables of hashables are mapped to single bytes (edlib.pyx:12-54)."""
    def ascii_bytes(s):
        if isinstance(s, bytes):
            return s
        if isinstance(s, str):
            b = s.encode("utf-8")
            if len(b) == len(s):
                return b
        return None

    qb, tb = ascii_bytes(query), ascii_bytes(target)
    if qb is not None and tb is not None:
        return qb, tb, additional_equalities
    alphabet = set(query).union(set(target))
    if len(alphabet) > 256:
        raise ValueError("query and target combined have more than 256 unique values, this is not supported.")
    mapping = {c: i for i, c in enumerate(alphabet)}
    qb = bytes(mapping[c] for c in query)
    tb = bytes(mapping[c] for c in target)
    if additional_equalities is not None:
        additional_equalities = [(mapping[a], mapping[b]) for a, b in additional_equalities if a in mapping and b in mapping]
    return qb, tb, additional_equalities


def align(query, target, mode="NW", task="distance", k=-1, additionalEqualities=None):
    """Same arguments and result dict as the reference: editDistance, alphabetLength, locations, cigar."""
    qb, tb, eqs = _to_bytes(query, target, additionalEqualities)
    if mode not in ("NW", "HW", "SHW"):
        mode = "NW"           # unknown strings leave the default config untouched (edlib.pyx:103-105)
    if task not in ("distance", "locations", "path"):
        task = "distance"
    if eqs is not None:
        eqs = [((a.encode("utf-8")[0] if isinstance(a, str) else int(a)), (b.encode("utf-8")[0] if isinstance(b, str) else int(b)))
               for a, b in eqs]
    q = _queue.active_queue()
    if q is not None:   # deferred mode (dysgu_b200.deferred()): submit now, resolve on first use
        return _queue.LazyEdlibResult(q, q.submit_ed(qb, tb, mode=mode, task=task, k=-1 if k is None else k, additional_equalities=eqs))
    res = _batch.ed_align_batch([qb], [tb], mode=mode, task=task, k=-1 if k is None else k, additional_equalities=eqs)
    d = res.as_dict(0)
    if d["status"] == 1:
        raise Exception("There was an error.")
    return {"editDistance": d["editDistance"], "alphabetLength": d["alphabetLength"], "locations": d["locations"], "cigar": d.get("cigar")}


def getNiceAlignment(alignResult, query, target, gapSymbol="-"):
    """Same behaviour as the reference helper (edlib.pyx:159-239); needs a cigar, i.e. task='path'."""
    if type(alignResult) is not dict:
        raise Exception("The object alignResult is expected to be a python dictionary. Please check the input alignResult.")
    if "locations" not in alignResult:
        raise Exception("The object alignResult is expected to contain a field 'locations'. Please check the input alignResult.")
    target_pos = alignResult["locations"][0][0] or 0
    query_pos = 0
    if "cigar" not in alignResult:
        raise Exception("The object alignResult is expected to contain a CIGAR string. Please check the input alignResult.")
    cigar = alignResult["cigar"]
    if not cigar:
        raise Exception("The object alignResult contains an empty CIGAR string. Users must run align() with task='path'. "
                        "Please check the input alignResult.")
    t_aln, m_aln, q_aln = [], [], []
    for num, op in re.findall(r"(\d+)(\D)", cigar):
        num = int(num)
        if op in "=X":
            t_aln.append(target[target_pos:target_pos + num]); target_pos += num
            q_aln.append(query[query_pos:query_pos + num]); query_pos += num
            m_aln.append(("|" if op == "=" else ".") * num)
        elif op == "D":
            t_aln.append(target[target_pos:target_pos + num]); target_pos += num
            q_aln.append(gapSymbol * num)
            m_aln.append(gapSymbol * num)
        elif op == "I":
            t_aln.append(gapSymbol * num)
            q_aln.append(query[query_pos:query_pos + num]); query_pos += num
            m_aln.append(gapSymbol * num)
        else:
            raise Exception("The CIGAR string from alignResult contains a symbol not '=', 'X', 'D', 'I'. "
                            "Please check the validity of alignResult and alignResult.cigar")
    return {"query_aligned": "".join(q_aln), "matched_aligned": "".join(m_aln), "target_aligned": "".join(t_aln)}
