"""StripedSmithWaterman / AlignmentStructure with the reference's Python surface, computed on the GPU.

Mirrors dysgu/scikitbio/_ssw_wrapper.pyx (constructor kwargs and defaults :539-554, errors :557-577,
mask length :583-588, flag derivation :654-666, result attributes :130-395).  Each call is a batch of
one through the C ABI; callers that want throughput use `align_many` (one query, many targets) or
dysgu_b200.batch.sw_align_batch.
"""
from __future__ import annotations

import numpy as np

from .. import batch as _batch
from .. import queue as _queue

_NT_ORDER = "ACGTN"


class AlignmentStructure:
    """Result view with the attribute names of the reference wrapper (_ssw_wrapper.pyx:73-395)."""

    _REPR_KEYS = ("optimal_alignment_score", "suboptimal_alignment_score", "query_begin", "query_end", "target_begin",
                  "target_end_optimal", "target_end_suboptimal", "cigar", "query_sequence", "target_sequence")

    def __init__(self, fields, cigar_ops, read_sequence, reference_sequence, index_starts_at):
        (self._score1, self._score2, self._ref_begin1, self._ref_end1, self._read_begin1, self._read_end1,
         self._ref_end2) = (int(x) for x in fields[:7])
        self._cigar_ops = cigar_ops
        self.read_sequence = read_sequence
        self.reference_sequence = reference_sequence
        self.index_starts_at = index_starts_at
        self._cigar_string = None

    def __getitem__(self, key):
        return getattr(self, key)

    def __repr__(self):
        return "{\n%s\n}" % ",\n".join("    {!r}: {!r}".format(k, self[k]) for k in self._REPR_KEYS)

    def __str__(self):
        score = "Score: %d" % self.optimal_alignment_score
        if self.query_sequence and self.cigar:
            target = self.aligned_target_sequence
            query = self.aligned_query_sequence
            n = len(query)
            if n > 13:
                target = target[:10] + "..."
                query = query[:10] + "..."
            return "\n".join([query, target, score, "Length: %d" % n])
        return score

    @property
    def optimal_alignment_score(self):
        return self._score1

    @property
    def suboptimal_alignment_score(self):
        return self._score2

    @property
    def target_begin(self):
        return self._ref_begin1 + self.index_starts_at if self._ref_begin1 >= 0 else -1

    @property
    def target_end_optimal(self):
        return self._ref_end1 + self.index_starts_at

    @property
    def target_end_suboptimal(self):
        return self._ref_end2 + self.index_starts_at

    @property
    def query_begin(self):
        return self._read_begin1 + self.index_starts_at if self._read_begin1 >= 0 else -1

    @property
    def query_end(self):
        return self._read_end1 + self.index_starts_at

    @property
    def cigar(self):
        if self._cigar_string is None:
            self._cigar_string = "".join("%d%s" % (int(c) >> 4, "MID"[int(c) & 0xF]) for c in self._cigar_ops)
        return self._cigar_string

    @property
    def query_sequence(self):
        return self.read_sequence

    @property
    def target_sequence(self):
        return self.reference_sequence

    @property
    def aligned_query_sequence(self):
        if self.query_sequence:
            return self._get_aligned_sequence(self.query_sequence, self._tuples_from_cigar(), self.query_begin,
                                              self.query_end, "D")
        return None

    @property
    def aligned_target_sequence(self):
        if self.target_sequence:
            return self._get_aligned_sequence(self.target_sequence, self._tuples_from_cigar(), self.target_begin,
                                              self.target_end_optimal, "I")
        return None

    def set_zero_based(self, is_zero_based):
        self.index_starts_at = 0 if is_zero_based else 1

    def is_zero_based(self):
        return self.index_starts_at == 0

    def _get_aligned_sequence(self, sequence, tuple_cigar, begin, end, gap_type):
        # begin/end arrive in the caller's index base, exactly as in the reference (:345-364)
        pieces = []
        seq = sequence[begin:end + 1]
        index = 0
        for length, op in tuple_cigar:
            if op == gap_type:
                pieces.append("-" * length)
            else:
                pieces.append(seq[index:index + length])
                index += length
        pieces.append(seq[index:end - begin + 1])
        return "".join(pieces)

    def _tuples_from_cigar(self):
        return [(int(c) >> 4, "MID"[int(c) & 0xF]) for c in self._cigar_ops]


class LazyAlignmentStructure(AlignmentStructure):
    """AlignmentStructure of a pair submitted in deferred mode (dysgu_b200.deferred()): the integer fields and the
    cigar are fetched from the queue on first use; everything else is the parent class."""

    _LAZY = ("_score1", "_score2", "_ref_begin1", "_ref_end1", "_read_begin1", "_read_end1", "_ref_end2", "_cigar_ops")

    def __init__(self, queue, ticket, read_sequence, reference_sequence, index_starts_at):
        self._queue, self._ticket = queue, ticket
        self.read_sequence = read_sequence
        self.reference_sequence = reference_sequence
        self.index_starts_at = index_starts_at
        self._cigar_string = None

    def __getattr__(self, name):   # only reached while the lazy fields are still missing
        if name in LazyAlignmentStructure._LAZY:
            fields, ops, status = self._queue.fetch_sw(self._ticket)
            if status != 0:
                raise RuntimeError("alignment failed: %s" % _batch.PAIR_STATUS.get(status, status))
            (self._score1, self._score2, self._ref_begin1, self._ref_end1, self._read_begin1, self._read_end1,
             self._ref_end2) = (int(x) for x in fields[:7])
            self._cigar_ops = ops
            self._queue = None
            return self.__dict__[name]
        raise AttributeError(name)


def _require_ascii(seq):
    """The reference indexes a 128-entry table per character (_ssw_wrapper.pyx:676-678): the first code point above 127
    raises IndexError.  str.isascii() answers the common case without a Python-level loop over the sequence."""
    if not seq.isascii():
        for ch in seq:
            if ord(ch) > 127:
                raise IndexError("index %d is out of bounds for axis 0 with size 128" % ord(ch))


class StripedSmithWaterman:
    """Same constructor and call surface as the reference class (_ssw_wrapper.pyx:397-709)."""

    def __init__(self, query_sequence, gap_open_penalty=5, gap_extend_penalty=2, score_size=2, mask_length=15,
                 mask_auto=True, score_only=False, score_filter=None, distance_filter=None, override_skip_babp=False,
                 protein=False, match_score=2, mismatch_score=-3, substitution_matrix=None, suppress_sequences=False,
                 zero_index=True, device=None):
        self.read_sequence = query_sequence
        if gap_open_penalty <= 0:
            raise ValueError("`gap_open_penalty` must be > 0")
        self.gap_open_penalty = gap_open_penalty
        if gap_extend_penalty <= 0:
            raise ValueError("`gap_extend_penalty` must be > 0")
        self.gap_extend_penalty = gap_extend_penalty
        self.distance_filter = 0 if distance_filter is None else distance_filter
        self.score_filter = 0 if score_filter is None else score_filter
        self.suppress_sequences = suppress_sequences
        self.is_protein = protein
        self.bit_flag = self._get_bit_flag(override_skip_babp, score_only)
        self.index_starts_at = 0 if zero_index else 1
        self.score_size = score_size
        self.device = device
        if substitution_matrix is None:
            if protein:
                raise Exception("Must provide a substitution matrix for protein sequences")
            self.matrix = None
            self.match_score, self.mismatch_score = match_score, mismatch_score
        else:
            if protein:
                raise NotImplementedError("protein (24x24) matrices are not provided by the device path")
            self.matrix = np.array([[int(substitution_matrix[r][c]) for c in _NT_ORDER] for r in _NT_ORDER], dtype=np.int8)
            self.match_score, self.mismatch_score = 0, 0
        if mask_auto:
            self.mask_length = int(len(query_sequence) / 2)
            if self.mask_length < mask_length:
                self.mask_length = mask_length
        else:
            self.mask_length = mask_length
        _require_ascii(query_sequence)
        self._query_bytes = query_sequence.encode("ascii")

    def _get_bit_flag(self, override_skip_babp, score_only):
        flag = 0
        if score_only:
            return flag
        if override_skip_babp:
            flag |= 0x8
        if self.distance_filter != 0:
            flag |= 0x4
        if self.score_filter != 0:
            flag |= 0x2
        if flag == 0 or flag == 8:
            flag |= 0x1
        return flag

    def _run(self, targets):
        tb = []
        for t in targets:
            _require_ascii(t)
            tb.append(t.encode("ascii"))
        res = _batch.sw_align_batch([self._query_bytes] * len(tb), tb, match=self.match_score, mismatch=self.mismatch_score,
                                    gap_open=self.gap_open_penalty, gap_extend=self.gap_extend_penalty,
                                    score_size=self.score_size, flag=self.bit_flag, filters=self.score_filter,
                                    filterd=self.distance_filter, mask_len=self.mask_length, matrix=self.matrix,
                                    device=self.device)
        out = []
        for i, t in enumerate(targets):
            if res.status[i] != 0:
                raise RuntimeError("alignment failed: %s" % _batch.PAIR_STATUS.get(int(res.status[i]), res.status[i]))
            if self.suppress_sequences:
                out.append(AlignmentStructure(res.fixed[i], res.cigar_of(i).copy(), "", "", self.index_starts_at))
            else:
                out.append(AlignmentStructure(res.fixed[i], res.cigar_of(i).copy(), self.read_sequence, t, self.index_starts_at))
        return out

    def __call__(self, target_sequence):
        q = _queue.active_queue()
        if q is None:
            return self._run([target_sequence])[0]
        # deferred mode: submit now, resolve on first use
        _require_ascii(target_sequence)
        ticket = q.submit_sw(self._query_bytes, target_sequence.encode("ascii"), match=self.match_score, mismatch=self.mismatch_score,
                             gap_open=self.gap_open_penalty, gap_extend=self.gap_extend_penalty, score_size=self.score_size,
                             flag=self.bit_flag, filters=self.score_filter, filterd=self.distance_filter, mask_len=self.mask_length,
                             matrix=self.matrix)
        if self.suppress_sequences:
            return LazyAlignmentStructure(q, ticket, "", "", self.index_starts_at)
        return LazyAlignmentStructure(q, ticket, self.read_sequence, target_sequence, self.index_starts_at)

    def align_many(self, target_sequences):
        """One profile, many targets in a single device batch (e.g. clip and revcomp(clip), post_call.py:53-76)."""
        return self._run(list(target_sequences))
