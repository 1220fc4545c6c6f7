"""Build libdysgu_b200_align.so (+ libdysgu_b200_compat.so, the int_peak micro-benchmark) in-tree with nvcc for sm_100a.

    python -m dysgu_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared library only needs the CUDA runtime (statically
linked); it has no Python or torch dependency - the Python side binds it with ctypes.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libdysgu_b200_align.so")
COMPAT = os.path.join(HERE, "libdysgu_b200_compat.so")
PEAK = os.path.join(HERE, "int_peak")
SOURCES = ["engine.cu", "sw.cu", "ed.cu", "queue.cu", "format.cu", "multi.cu", "minimizer.cu",
           "sw_fine_g1.cu", "sw_fine_g2.cu", "sw_fine_g4.cu", "sw_fine_g8.cu", "sw_fine_g16.cu", "sw_fine_g32.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "dysgu_b200_align.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--diag-suppress", "177"] + os.environ.get("DB200_NVCC_FLAGS", "").split()


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    deps = [os.path.join(CSRC, h) for h in HEADERS if os.path.exists(os.path.join(CSRC, h))]
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + deps):
            jobs.append([NVCC] + ARCH + FLAGS + ["-c", s, "-o", o])
    if jobs:
        with ThreadPoolExecutor(min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(_run, jobs):
                if verbose and out:
                    print(out)
    if force or jobs or _newer(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-o", LIB] + objs)
    # the reference's own symbol names (ssw_init, edlibAlign, ...) as a library of their own over the engine: what dysgu's
    # unmodified _ssw_wrapper.pyx / edlib.pyx link against instead of ssw.c / edlib.cpp (oracle/build_compat.sh)
    compat_src = os.path.join(CSRC, "compat.c")
    if force or jobs or _newer(COMPAT, [compat_src, LIB] + deps):
        _run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-fvisibility=hidden", compat_src, "-o", COMPAT,
              "-L" + HERE, "-ldysgu_b200_align", "-Wl,-rpath,$ORIGIN"])
    peak_src = os.path.join(CSRC, "int_peak.cu")
    if force or _newer(PEAK, [peak_src]):
        _run([NVCC] + ARCH + ["-O3", "-lineinfo", peak_src, "-o", PEAK])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
