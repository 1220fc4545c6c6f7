"""Pair sharding over the GPUs of one box.

Every alignment pair is independent, so the data path needs no collective: a batch is cut into
contiguous, equally sized shards (one per device / rank), each shard is aligned on its own GPU and the
results are scattered back in submission order.  Two drivers:

  * `align_multi_gpu`  - one process, one host thread per device (the ctypes calls release the GIL)
  * `shard_bounds` + `gather_in_order` - one process per GPU under torch.distributed (what bench.py does)
"""
from __future__ import annotations

import threading
from typing import Callable, List, Sequence, Tuple

import numpy as np

from .workloads import SeqBatch


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) per rank; sizes differ by at most one pair."""
    base, extra = divmod(n, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def take_shard(batch: SeqBatch, begin: int, end: int) -> SeqBatch:
    off = batch.off[begin:end + 1] - batch.off[begin]
    return SeqBatch(batch.buf[batch.off[begin]:batch.off[end]], off)


def merge_csr(parts_off: Sequence[np.ndarray], parts_val: Sequence[np.ndarray]):
    """Concatenate per-shard CSR outputs (cigars, locations) back into one CSR in submission order."""
    total_n = sum(len(o) - 1 for o in parts_off)
    off = np.zeros(total_n + 1, dtype=np.int64)
    vals, pos, base = [], 0, 0
    for o, v in zip(parts_off, parts_val):
        k = len(o) - 1
        off[pos + 1:pos + k + 1] = o[1:] + base
        base += int(o[-1])
        pos += k
        vals.append(v[:int(o[-1])])
    return off, (np.concatenate(vals) if vals else np.zeros(0))


def align_multi_gpu(fn: Callable, queries: SeqBatch, targets: SeqBatch, devices: Sequence[int], **kw):
    """Run fn(queries_shard, targets_shard, device=d, **kw) on each device from its own host thread.

    fn is dysgu_b200.batch.sw_align_batch or ed_align_batch; returns the per-shard results in rank order.
    """
    bounds = shard_bounds(len(queries), len(devices))
    results = [None] * len(devices)
    errors = []

    def work(i, dev, b, e):
        try:
            results[i] = fn(take_shard(queries, b, e), take_shard(targets, b, e), device=dev, **kw)
        except Exception as exc:  # propagate to the caller thread
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i, d, b, e)) for i, (d, (b, e)) in enumerate(zip(devices, bounds))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def gather_in_order(local: np.ndarray, dist) -> np.ndarray:
    """all_gather of per-rank fixed-size result rows (rank order == submission order for contiguous shards)."""
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(local))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64))
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype)
    pad[:t.shape[0]] = t
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return np.concatenate([o[:int(s.item())].numpy() for o, s in zip(outs, sizes)], axis=0)
