"""CPU, world_size 2 over gloo: the N>1 host path - contiguous pair shards per rank, results gathered back in
submission order equal the single-process result.  The per-shard aligner here is the CPU oracle (tests only)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from dysgu_b200 import shard, workloads
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    qs, ts, prm = workloads.remap_windows(41, seed=99)  # every rank derives the same batch, then takes its shard
    b, e = shard.shard_bounds(len(qs), world)[rank]
    ql, tl = shard.take_shard(qs, b, e), shard.take_shard(ts, b, e)
    r = O.oracle_sw((ql.buf, ql.off), (tl.buf, tl.off), prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"], csr=True)
    full = shard.gather_in_order(r.fixed, dist)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_preserves_submission_order():
    import torch.multiprocessing as mp
    from dysgu_b200 import workloads
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    qs, ts, prm = workloads.remap_windows(41, seed=99)
    ref = O.oracle_sw((qs.buf, qs.off), (ts.buf, ts.off), prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"], csr=True)
    assert full.shape == ref.fixed.shape and np.array_equal(full, ref.fixed)
