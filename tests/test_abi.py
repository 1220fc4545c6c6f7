"""CPU: the C-ABI library loads and exports every symbol include/dysgu_b200_align.h declares; compute fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dysgu_b200 import _lib, batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dysgu_b200_align.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(db200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_device_count():
    L = _lib.lib()
    assert b"sm_100a" in L.db200_version()
    assert _lib.device_count() >= 0


@pytest.mark.skipif(_lib.device_count() > 0, reason="a GPU is present")
def test_compute_fails_loudly_without_gpu():
    with pytest.raises(_lib.AlignEngineError, match="no CUDA device"):
        batch.sw_align_batch([b"ACGT"], [b"ACGT"])
    with pytest.raises(_lib.AlignEngineError, match="no CUDA device"):
        batch.ed_align_batch([b"ACGT"], [b"ACGT"])
    from dysgu_b200.scikitbio import StripedSmithWaterman
    with pytest.raises(_lib.AlignEngineError):
        StripedSmithWaterman("ACGT")("ACGT")


def test_struct_layouts_match_header():
    # db200_sw_params: 8 int32 + pointer + int32 (+pad); db200_ed_params: 3 int32 (+pad) + pointer + int32 (+pad)
    assert C.sizeof(_lib.SwParams) == 48
    assert C.sizeof(_lib.EdParams) == 32
    assert batch.SW_FIELDS == ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len")


def test_alignment_to_cigar_is_host_side_formatting():
    # edlibAlignmentToCigar (edlib.cpp:298-347): run-length text; standard folds match and mismatch into M
    L = _lib.lib()
    L.db200_edlibAlignmentToCigar.restype = C.c_void_p
    L.db200_edlibAlignmentToCigar.argtypes = [C.c_char_p, C.c_int, C.c_int]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]

    def cigar(ops, fmt):
        p = L.db200_edlibAlignmentToCigar(bytes(ops), len(ops), fmt)
        if not p:
            return None
        s = C.string_at(p).decode()
        libc.free(p)
        return s

    ops = [0, 0, 0, 0, 3, 0, 0, 0]                       # the SURVEY 8c vector: "4=1X3="
    assert cigar(ops, 1) == "4=1X3="
    assert cigar(ops, 0) == "8M"
    assert cigar([1, 1, 0, 3, 3, 2, 2, 2, 0], 1) == "2I1=2X3D1="
    assert cigar([1, 1, 0, 3, 3, 2, 2, 2, 0], 0) == "2I3M3D1M"
    assert cigar([0] * 1234, 1) == "1234="
    assert cigar([], 1) == ""
    assert cigar([0, 4], 1) is None and cigar([0], 2) is None


@pytest.mark.skipif(_lib.device_count() > 0, reason="a GPU is present")
def test_queue_needs_a_device():
    from dysgu_b200.queue import AlignQueue
    with pytest.raises(_lib.AlignEngineError, match="no CUDA device"):
        AlignQueue(0)
    import dysgu_b200
    with pytest.raises(_lib.AlignEngineError):
        with dysgu_b200.deferred():
            pass


def test_python_and_c_cigar_formatting_agree():
    """batch.ops_to_cigar (vectorised, used for batches) and db200_edlibAlignmentToCigar (the C entry point) on random
    operation arrays, both formats."""
    L = _lib.lib()
    L.db200_edlibAlignmentToCigar.restype = C.c_void_p
    L.db200_edlibAlignmentToCigar.argtypes = [C.c_char_p, C.c_int, C.c_int]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(5)
    for _ in range(200):
        n = int(rng.integers(1, 400))
        ops = np.repeat(rng.integers(0, 4, size=n, dtype=np.uint8), rng.integers(1, 6, size=n))
        for fmt in (0, 1):
            p = L.db200_edlibAlignmentToCigar(ops.tobytes(), len(ops), fmt)
            s = C.string_at(p).decode()
            libc.free(p)
            assert s == batch.ops_to_cigar(ops, extended=bool(fmt))
    assert batch.ops_to_cigar(np.zeros(0, np.uint8)) == ""
