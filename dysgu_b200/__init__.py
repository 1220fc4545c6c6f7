"""dysgu_b200 - B200-native batched alignment engine for dysgu's SSW / edlib hot path.

  dysgu_b200.scikitbio.StripedSmithWaterman   drop-in for dysgu.scikitbio._ssw_wrapper.StripedSmithWaterman
  dysgu_b200.edlib.align                      drop-in for dysgu.edlib.align
  dysgu_b200.batch.sw_align_batch / ed_align_batch   batched API (numpy CSR in, numpy out)
  dysgu_b200.queue.AlignQueue / dysgu_b200.deferred()   ticket queue (submit / flush / fetch) and the deferred mode of the
                                              drop-in classes: calls inside the block are collected into one device batch
  dysgu_b200.shard                            pair sharding over the GPUs of one box

All compute happens in libdysgu_b200_align.so (CUDA, sm_100a).  There is no CPU fallback.
"""
__version__ = "0.1.0"


def deferred(device=None):
    """Context manager: StripedSmithWaterman(...)(...) / edlib.align(...) inside the block submit to one queue (see queue.py)."""
    from .queue import deferred as _deferred
    return _deferred(device)


def install():
    """Route dysgu's two alignment extension modules to this engine without touching dysgu's files.

    Afterwards `from dysgu.scikitbio._ssw_wrapper import StripedSmithWaterman` (consensus.pyx:16, re_map.py:1,
    post_call.py:15, filter_normals.py:18), `import dysgu.edlib as edlib` (re_map.py:6) and `from dysgu.edlib import edlib`
    (filter_normals.py:20) resolve to dysgu_b200.  Call it before dysgu's own modules are imported; the `dysgu` package
    itself is imported normally.  Returns the names that were installed.
    """
    import sys
    from . import edlib as _edlib_pkg
    from . import scikitbio as _skbio_pkg
    from .edlib import edlib as _edlib_mod
    from .scikitbio import _ssw_wrapper as _ssw_mod
    table = {"dysgu.scikitbio": _skbio_pkg, "dysgu.scikitbio._ssw_wrapper": _ssw_mod,
             "dysgu.edlib": _edlib_pkg, "dysgu.edlib.edlib": _edlib_mod}
    for name, mod in table.items():
        sys.modules[name] = mod
        parent, _, leaf = name.rpartition(".")
        if parent in sys.modules and parent != "dysgu_b200":
            setattr(sys.modules[parent], leaf, mod)   # `from dysgu import edlib` after dysgu has been imported
    return sorted(table)
