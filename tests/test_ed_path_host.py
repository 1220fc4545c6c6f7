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
