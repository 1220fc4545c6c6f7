"""TEST INFRASTRUCTURE: a batch 'engine' with the call surface of dysgu_b200.batch, computed by the reference's own
compiled cores (oracle/_ref through oracle/ref_driver.c).  Lets the CPU suite check the replay logic of dysgu_b200.patch
(which only needs *an* engine) without a GPU.  Never used by the product."""
import numpy as np

from oracle import oracle as O


class _Sw:
    def __init__(self, r):
        self.fixed, self.cigar, self.cigar_off, self.status = r.fixed, r.cigar, r.cigar_off, r.status

    def cigar_of(self, i):
        return self.cigar[self.cigar_off[i]:self.cigar_off[i + 1]]


class RefEngine:
    def __init__(self):
        self.sw_batches = self.ed_batches = self.pairs = 0

    def sw_align_batch(self, queries, targets, match=2, mismatch=-3, gap_open=5, gap_extend=2, score_size=2, flag=1, filters=0,
                       filterd=0, mask_len=None, matrix=None, device=None, want_cigar=True):
        assert matrix is None
        self.sw_batches += 1
        self.pairs += len(queries)
        ml = None if mask_len is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mask_len, np.int32), (len(queries),)))
        return _Sw(O.ref_sw(list(queries), list(targets), match, mismatch, gap_open, gap_extend, score_size, flag, filters, filterd,
                            mask_len=ml))

    def ed_align_batch(self, queries, targets, mode="NW", task="distance", k=-1, additional_equalities=None, device=None):
        assert not additional_equalities
        self.ed_batches += 1
        self.pairs += len(queries)
        return O.ref_ed(list(queries), list(targets), mode, task, None if (np.ndim(k) == 0 and k < 0) else k)
