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
