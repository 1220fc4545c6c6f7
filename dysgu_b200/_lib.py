"""ctypes binding of libdysgu_b200_align.so (C ABI: include/dysgu_b200_align.h).

The library is the product: there is no CPU fallback.  Importing this module never touches a GPU;
the first compute call does, and fails loudly if the shared library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdysgu_b200_align.so")


class AlignEngineError(RuntimeError):
    pass


class SwParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32),
                ("score_size", C.c_int32), ("flag", C.c_int32), ("filters", C.c_int32), ("filterd", C.c_int32),
                ("mat", C.POINTER(C.c_int8)), ("n", C.c_int32)]


class EdParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("task", C.c_int32), ("k", C.c_int32), ("eq_pairs", C.c_char_p),
                ("n_eq_pairs", C.c_int32)]


_lib = None


def lib():
    """Load the shared library (built in-tree by `python -m dysgu_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AlignEngineError(
            "libdysgu_b200_align.so is missing (%s): build it with `python -m dysgu_b200.build`; "
            "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.db200_last_error.restype = C.c_char_p
    L.db200_version.restype = C.c_char_p
    L.db200_workspace_bytes.restype = C.c_int64
    L.db200_kernel_launches.restype = C.c_int64
    L.db200_ed_word_steps.restype = C.c_int64
    L.db200_forked_after_init.restype = C.c_int
    L.db200_default_device.restype = C.c_int
    L.db200_host_alloc.restype = C.c_void_p
    L.db200_host_alloc.argtypes = [C.c_int64]
    L.db200_host_free.argtypes = [C.c_void_p]
    L.db200_sw_batch_host.restype = C.c_int
    L.db200_sw_batch_host.argtypes = [C.c_int, C.POINTER(SwParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.db200_sw_batch_device.restype = C.c_int
    L.db200_sw_batch_device.argtypes = [C.c_int, C.POINTER(SwParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.db200_sw_batch_host_slices.restype = C.c_int
    L.db200_sw_batch_host_slices.argtypes = [C.c_int, C.POINTER(SwParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                             C.c_void_p]
    L.db200_sw_batch_host_packed.restype = C.c_int
    L.db200_sw_batch_host_packed.argtypes = [C.c_int, C.POINTER(SwParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.db200_sw_batch_device_packed.restype = C.c_int
    L.db200_sw_batch_device_packed.argtypes = [C.c_int, C.POINTER(SwParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                               C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p]
    L.db200_sw_pack_host.restype = C.c_int
    L.db200_sw_pack_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]
    L.db200_sw_batch_host_multi.restype = C.c_int
    L.db200_sw_batch_host_multi.argtypes = [C.c_void_p, C.c_int32, C.POINTER(SwParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.db200_ed_batch_host_multi.restype = C.c_int
    L.db200_ed_batch_host_multi.argtypes = [C.c_void_p, C.c_int32, C.POINTER(EdParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    if hasattr(L, "db200_ed_batch_host"):
        L.db200_ed_batch_host.restype = C.c_int
        L.db200_ed_batch_host.argtypes = [C.c_int, C.POINTER(EdParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.db200_ed_batch_device.restype = C.c_int
        L.db200_ed_batch_device.argtypes = [C.c_int, C.POINTER(EdParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    if hasattr(L, "db200_ed_path_batch_host"):
        L.db200_ed_path_batch_host.restype = C.c_int
        L.db200_ed_path_batch_host.argtypes = [C.c_int, C.POINTER(EdParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                               C.c_void_p]
        L.db200_ed_path_batch_device.restype = C.c_int
        L.db200_ed_path_batch_device.argtypes = [C.c_int, C.POINTER(EdParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                                 C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                                 C.c_void_p, C.c_void_p]
    if hasattr(L, "db200_sw_cigar_text"):
        L.db200_sw_cigar_text.restype = C.c_int
        L.db200_sw_cigar_text.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.db200_sw_aligned_text.restype = C.c_int
        L.db200_sw_aligned_text.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                            C.c_int64, C.c_void_p]
        L.db200_ed_cigar_text.restype = C.c_int
        L.db200_ed_cigar_text.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().db200_last_error().decode("utf-8", "replace")
        raise AlignEngineError("%s failed (code %d): %s" % (what, rc, msg))


def device_count() -> int:
    return int(lib().db200_device_count())
