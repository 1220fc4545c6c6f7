"""CPU: the edit-operation path code that the device runs for edlib task "path" (dysgu_b200/csrc/ed_path.cuh), built
for the host by tests/ed_path_host.cpp, against vectors generated from the reference's compiled wrapper
(tests/golden/ed_path.json).  The GPU test (test_ed_path_gpu.py) checks the same vectors through the C ABI."""
import ctypes as C
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
OPS = "=IDX"


def cigar_of(ops):
    out, i = [], 0
    while i < len(ops):
        j = i
        while j < len(ops) and ops[j] == ops[i]:
            j += 1
        out.append("%d%s" % (j - i, OPS[ops[i]]))
        i = j
    return "".join(out)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("edpath") / "ed_path_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(HERE, "ed_path_host.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    L.ed_path_host.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_int)]
    return L


def host_path(L, q, t, eq=b""):
    out = C.create_string_buffer(len(q) + len(t) + 1)
    d = C.c_int(0)
    n = L.ed_path_host(q, len(q), t, len(t), eq, len(eq) // 2, out, C.byref(d))
    assert n >= 0
    return d.value, cigar_of(out.raw[:n])


def test_paths_match_the_reference(lib):
    blocks = json.load(open(os.path.join(HERE, "golden", "ed_path.json")))
    checked = big = 0
    for b in blocks:
        eq = "".join(x + y for x, y in b.get("equalities", [])).encode()
        for c in b["cases"]:
            if c["editDistance"] < 0:           # above k: the reference reports no path
                assert c["cigar"] is None
                continue
            start, end = c["locations"][0]
            start = 0 if start is None else start
            sl = c["target"][start:end + 1].encode()       # the slice the reference aligns globally (edlib.cpp:270-283)
            d, cig = host_path(lib, c["query"].encode(), sl, eq)
            assert d == c["editDistance"] and cig == c["cigar"], (b["mode"], b["k"], len(c["query"]), len(sl))
            checked += 1
            big += len(c["query"]) >= 1800
    assert checked > 1000 and big >= 10


def test_paths_match_the_scalar_oracle_on_random_pairs(lib):
    """The shared path code (bit-vector sweeps, stored-column leaves) against the integer-matrix restatement
    (oracle/ed_path_oracle.c) on random global alignments, with wildcard equalities on a third of them and a handful of
    pairs in the halving regime."""
    import random
    import numpy as np
    import pairgen
    from oracle import oracle as O
    rng = random.Random(31337)
    eqs = [("N", "A"), ("N", "C"), ("N", "G"), ("N", "T")]
    for block in range(3):
        qs, ts = [], []
        for i in range(2500 if block < 2 else 12):
            L = rng.choice([1, 5, 63, 64, 65, 128, 129, 200, 400, 700]) if block < 2 else rng.choice([2600, 3300, 4800])
            alpha = "ACGTN" if block == 1 else ("AC" if i % 7 == 0 else "ACGT")
            q = pairgen.rand_dna(rng, L, alpha)
            e = rng.choice([0.0, 0.01, 0.05, 0.2])
            t = (pairgen.mutate(rng, q, e, e, e, alpha) if i % 5 else pairgen.rand_dna(rng, rng.randint(1, 2 * L), alpha)) or "A"
            qs.append(q.encode()); ts.append(t.encode())
        use = eqs if block == 1 else []
        r = O.oracle_ed_path(qs, ts, "NW", None, use)
        eqb = "".join(a + b for a, b in use).encode()
        for i in range(len(qs)):
            d, cig = host_path(lib, qs[i], ts[i], eqb)
            ops = r.path[r.path_off[i]:r.path_off[i] + r.path_len[i]]
            assert d == r.fixed[i, 1] and cig == cigar_of(bytes(ops)), (block, i, len(qs[i]), len(ts[i]))
