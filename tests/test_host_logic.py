"""CPU: host-side logic - workload generators, CSR helpers, wrapper argument handling, sharding."""
import numpy as np
import pytest

from dysgu_b200 import shard, workloads
from dysgu_b200.edlib import edlib as ed_wrap
from dysgu_b200.scikitbio import StripedSmithWaterman


def test_workloads_are_deterministic_and_shaped():
    q1, t1, prm = workloads.remap_windows(500, seed=3)
    q2, t2, _ = workloads.remap_windows(500, seed=3)
    assert np.array_equal(q1.buf, q2.buf) and np.array_equal(t1.off, t2.off)
    assert set(np.unique(q1.lengths)) <= {1000, 1500, 2000}
    assert t1.lengths.min() >= 1 and t1.lengths.max() <= 330
    assert prm == dict(match=2, mismatch=-8, gap_open=6, gap_extend=1)
    assert set(np.unique(q1.buf)) <= set(b"ACGT") and set(np.unique(t1.buf)) <= set(b"acgt")
    a, b, _ = workloads.insertion_linking(200, seed=4, platform="ont")
    assert len(a) == len(b) == 200 and a.lengths.max() <= 10000


def test_seqbatch_take_and_slices():
    b = workloads.SeqBatch.from_list([b"AAAA", b"", b"CCG", b"T"])
    assert b.tolist() == [b"AAAA", b"", b"CCG", b"T"]
    assert b.take([2, 0]).tolist() == [b"CCG", b"AAAA"]
    s = workloads.slices(workloads.SeqBatch.from_list([b"ACGTACGT", b"TTTTGGGG"]), np.array([1, 4]), np.array([3, 4]))
    assert s.tolist() == [b"CGT", b"GGGG"]


def test_mutate_rates():
    rng = np.random.default_rng(1)
    a = workloads.random_dna(rng, np.full(200, 500))
    b = workloads.mutate(rng, a, 0.0, 0.0, 0.0)
    assert np.array_equal(a.buf, b.buf)
    c = workloads.mutate(rng, a, 0.1, 0.0, 0.0)
    frac = float((a.buf != c.buf).mean())
    assert 0.07 < frac < 0.13
    d = workloads.mutate(rng, a, 0.0, 0.05, 0.05)
    assert abs(int(d.off[-1]) - int(a.off[-1])) < 0.02 * int(a.off[-1])


def test_wrapper_argument_errors_match_reference():
    with pytest.raises(ValueError, match="gap_open_penalty"):
        StripedSmithWaterman("ACGT", gap_open_penalty=0)
    with pytest.raises(ValueError, match="gap_extend_penalty"):
        StripedSmithWaterman("ACGT", gap_extend_penalty=-1)
    with pytest.raises(Exception, match="substitution matrix"):
        StripedSmithWaterman("ACGT", protein=True)
    with pytest.raises(IndexError):
        StripedSmithWaterman("ACGä")
    s = StripedSmithWaterman("A" * 100)
    assert s.mask_length == 50 and s.bit_flag == 1
    assert StripedSmithWaterman("A" * 10).mask_length == 15
    assert StripedSmithWaterman("A" * 100, mask_auto=False, mask_length=20).mask_length == 20
    assert StripedSmithWaterman("ACGT", score_only=True).bit_flag == 0
    assert StripedSmithWaterman("ACGT", override_skip_babp=True).bit_flag == 9
    assert StripedSmithWaterman("ACGT", score_filter=10).bit_flag == 2
    assert StripedSmithWaterman("ACGT", distance_filter=10, score_filter=3).bit_flag == 6


def test_edlib_byte_mapping():
    q, t, eq = ed_wrap._to_bytes("ACGT", b"AC", None)
    assert q == b"ACGT" and t == b"AC" and eq is None
    q, t, eq = ed_wrap._to_bytes([12, "x", 12], ["x"], [("x", 12), ("zz", 1)])
    assert len(q) == 3 and q[0] == q[2] and t[0] == q[1] and len(eq) == 1
    with pytest.raises(ValueError):
        ed_wrap._to_bytes(list(range(300)), [1], None)


def test_shard_bounds_and_merge():
    assert shard.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    b = workloads.SeqBatch.from_list([b"AA", b"C", b"GGG", b"T", b""])
    parts = [shard.take_shard(b, lo, hi) for lo, hi in shard.shard_bounds(len(b), 2)]
    assert sum((p.tolist() for p in parts), []) == b.tolist()
    off, val = shard.merge_csr([np.array([0, 2, 3]), np.array([0, 1, 1, 4])], [np.array([7, 8, 9]), np.array([1, 2, 3, 4])])
    assert list(off) == [0, 2, 3, 4, 4, 7] and list(val) == [7, 8, 9, 1, 2, 3, 4]


def test_cigar_stats_match_a_string_parse():
    """SwResults.cigar_stats / callers.matches_with_max_gap_batch against the per-string parse the reference does
    (merge_svs.pyx:196-222), on synthetic cigar arrays incl. empty cigars."""
    import re
    import numpy as np
    from dysgu_b200 import batch, callers
    rng = np.random.default_rng(5)
    cigs, off = [], [0]
    for i in range(300):
        nops = int(rng.integers(0, 9)) if i % 7 else 0
        ops = []
        for j in range(nops):
            op = 0 if j % 2 == 0 else int(rng.integers(1, 3))
            ops.append((int(rng.integers(1, 400 if op == 0 else 40)) << 4) | op)
        cigs += ops
        off.append(len(cigs))
    r = batch.SwResults(np.zeros((300, 8), np.int32), np.array(cigs, dtype=np.uint32), np.array(off, dtype=np.int64), np.zeros(300, np.int32))
    st = r.cigar_stats()
    for max_gap, tol in ((30, 0.1), (15, 0.02), (1000, 1.0)):
        got = callers.matches_with_max_gap_batch(r, max_gap, tol)
        for i in range(300):
            s = r.cigar_str(i)
            matches = gaps = 0
            want = None
            for num, op in re.findall(r"(\d+)([MID])", s):
                num = int(num)
                if op in "DI":
                    if num > max_gap:
                        want = 0
                        break
                    gaps += num
                else:
                    matches += num
            if want is None:
                want = 0 if (matches and np.float32(gaps) / np.float32(matches) > np.float32(tol)) else matches
            assert got[i] == want, (s, max_gap, tol)
    assert st.shape == (300, 3) and (st[np.diff(off) == 0] == 0).all()


def test_install_routes_dysgu_imports(tmp_path, monkeypatch):
    """dysgu_b200.install(): the import statements dysgu's modules use resolve to the engine's drop-ins (a stand-in `dysgu`
    package, since the real one needs pysam to import)."""
    import importlib
    import sys
    import dysgu_b200
    pkg = tmp_path / "dysgu"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    monkeypatch.syspath_prepend(str(tmp_path))
    for name in [n for n in sys.modules if n == "dysgu" or n.startswith("dysgu.")]:
        monkeypatch.delitem(sys.modules, name)
    names = dysgu_b200.install()
    try:
        assert names == ["dysgu.edlib", "dysgu.edlib.edlib", "dysgu.scikitbio", "dysgu.scikitbio._ssw_wrapper"]
        from dysgu.scikitbio._ssw_wrapper import StripedSmithWaterman as S1       # consensus.pyx:16, re_map.py:1
        from dysgu.scikitbio import StripedSmithWaterman as S2, AlignmentStructure  # scikitbio/__init__.py:10-11
        import dysgu.edlib as e1                                                  # re_map.py:6
        from dysgu.edlib import edlib as e2                                       # filter_normals.py:20
        assert S1 is StripedSmithWaterman and S2 is StripedSmithWaterman and AlignmentStructure is not None
        assert e1.align is ed_wrap.align and e2.align is ed_wrap.align and e2.getNiceAlignment is ed_wrap.getNiceAlignment
        assert importlib.import_module("dysgu").__file__.startswith(str(tmp_path))
    finally:
        for name in [n for n in sys.modules if n == "dysgu" or n.startswith("dysgu.")]:
            sys.modules.pop(name, None)


def test_host_packer_writes_the_pool_format():
    """db200_sw_pack_host: codes of _ssw_wrapper.pyx:60-68 as nibbles, 8 per word, sequences padded with code 5 to 32 bases."""
    from dysgu_b200 import batch
    seqs = [b"ACGTNacgtnUu-xyz", b"", b"A", b"ACGT" * 8, b"ACGT" * 8 + b"C", b"T" * 70]
    p = batch.pack_seqs(seqs, threads=3)
    tab = np.full(256, 4, np.uint8)
    for ch, c in zip(b"AaUuCcGgTt", [0, 0, 0, 0, 1, 1, 2, 2, 3, 3]):
        tab[ch] = c
    assert list(p.lengths) == [len(s) for s in seqs]
    for i, s in enumerate(seqs):
        nu = int(p.unit_off[i + 1] - p.unit_off[i])
        assert nu == (len(s) + 31) // 32
        w = p.units[p.unit_off[i] * 4:p.unit_off[i + 1] * 4]
        nib = np.array([(int(w[k >> 3]) >> ((k & 7) * 4)) & 15 for k in range(nu * 32)], dtype=np.uint8)
        exp = np.concatenate([tab[np.frombuffer(s, np.uint8)], np.full(nu * 32 - len(s), 5, np.uint8)]) if nu else np.zeros(0, np.uint8)
        assert np.array_equal(nib, exp), i
