"""world_size-2 gloo test (CPU) of the data-parallel plumbing: sharding + bucketed all-reduce +
the 1/world scale folded into ADAM give the single-process mean gradient."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dpp_b200 import dp


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    full = rng.randn(8, 1000).astype(np.float32)            # per-sample gradients of a global batch of 8
    lo, hi = dp.shard_range(8, rank, world)
    g = torch.from_numpy(full[lo:hi].mean(axis=0).copy())   # each rank: mean over its local batch
    dp.make_allreduce(dist, bucket_elems=300)(g)
    g *= 1.0 / world
    ret[rank] = g.numpy().copy()
    dist.destroy_process_group()


def test_shard_and_allreduce_equals_global_mean():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29571, ret), nprocs=world, join=True)
    full = np.random.RandomState(0).randn(8, 1000).astype(np.float32)
    for r in range(world):
        assert np.allclose(ret[r], full.mean(axis=0), atol=1e-6)


def test_bucket_bounds_cover_arena_tail_first():
    b = dp.bucket_bounds(1000, 300)
    assert b[0] == (700, 1000) and b[-1][0] == 0
    assert sum(hi - lo for lo, hi in b) == 1000
    assert dp.shard_range(10, 3, 4) == (9, 10) and dp.shard_range(10, 0, 4) == (0, 3)
