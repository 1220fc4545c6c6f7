"""GPU: the drop-in proof at the C boundary.  The reference's own two Cython wrappers - _ssw_wrapper.pyx and
edlib.pyx, unmodified, cythonised from the reference checkout by oracle/build_compat.sh - are linked against
dysgu_b200/libdysgu_b200_compat.so (ssw_init / ssw_align / init_destroy / align_destroy, ssw.h:72-125; edlibAlign /
edlibNewAlignConfig / edlibFreeAlignResult / edlibAlignmentToCigar, edlib.h:140-271) instead of ssw.c / edlib.cpp, the way
meson.build:177-193 links the originals.  Every wrapper-level golden vector (generated from the reference's wrappers over
the reference's cores) must come out of these modules bit for bit: same attributes, same strings, same dicts."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "oracle", "_ref", "pycompat")
G = os.path.join(ROOT, "tests", "golden")


def _load(name, rel):
    import sysconfig
    path = os.path.join(COMPAT, *rel) + sysconfig.get_config_var("EXT_SUFFIX")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/pycompat not built (oracle/build_compat.sh needs the reference checkout and Cython)")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_compat_library_exports_the_reference_symbol_names():
    import ctypes as C
    lib = os.path.join(ROOT, "dysgu_b200", "libdysgu_b200_compat.so")
    assert os.path.exists(lib), "python -m dysgu_b200.build builds it"
    L = C.CDLL(lib)
    for name in ("ssw_init", "ssw_align", "init_destroy", "align_destroy", "edlibAlign", "edlibNewAlignConfig",
                 "edlibDefaultAlignConfig", "edlibFreeAlignResult", "edlibAlignmentToCigar"):
        assert hasattr(L, name), name


@pytest.mark.gpu
def test_reference_wrappers_over_the_gpu_engine_reproduce_the_wrapper_golden():
    ssw = _load("db200_compat._ssw_wrapper", ("dysgu", "scikitbio", "_ssw_wrapper"))
    edl = _load("db200_compat.edlib", ("dysgu", "edlib", "edlib"))
    with open(os.path.join(G, "wrapper.json")) as f:
        g = json.load(f)
    for case in g["sw"]:
        a = ssw.StripedSmithWaterman(case["query"], **case["kwargs"])(case["target"])
        for k, want in case["result"].items():
            assert a[k] == want, (case["query"][:40], case["kwargs"], k)
        assert str(a) == case["str"]
    for case in g["ed"]:
        got = edl.align(case["query"], case["target"], case["mode"], case["task"], case["k"])
        want = dict(case["result"])
        want["locations"] = [tuple(x) for x in want["locations"]]
        assert got == want, case
    # one profile, several targets (post_call.py:53-76) and the memoised cigar of the reference class
    q = ssw.StripedSmithWaterman("ACGTACGTTAGCTAGCTAGGATCGATCGTAGCTAGCTAGCATCGAT", gap_extend_penalty=1, suppress_sequences=True)
    a, b = q("CGTTAGCTAGCAGGATCG"), q("CGATCCTGCTAGCTAACG")
    assert (a.optimal_alignment_score, a.cigar) == (30, "11M1D7M") or a.optimal_alignment_score > b.optimal_alignment_score


@pytest.mark.gpu
def test_legacy_entry_points_honour_db200_device():
    """db200_ssw_align / db200_edlibAlign carry no device argument: DB200_DEVICE (else LOCAL_RANK) selects it."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from dysgu_b200 import _lib\n"
            "L = _lib.lib(); print(L.db200_default_device())\n" % ROOT)
    for env, want in (({"DB200_DEVICE": "0"}, "0"), ({"LOCAL_RANK": "3"}, "3"), ({"DB200_DEVICE": "2", "LOCAL_RANK": "5"}, "2")):
        e = {k: v for k, v in os.environ.items() if k not in ("DB200_DEVICE", "LOCAL_RANK")}
        e.update(env)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=120)
        assert r.stdout.strip() == want, (env, r.stdout, r.stderr[-500:])
