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


def test_local_rows_partition_every_minibatch():
    """every global minibatch is split into equal contiguous shares; local minibatch m of rank r is rows
    [m*B + r*b, m*B + (r+1)*b) of the aligned set; the ranks together hold every row exactly once"""
    n, B, world = 96, 32, 4
    rows = [dp.local_rows(n, B, r, world) for r in range(world)]
    assert sorted(np.concatenate(rows).tolist()) == list(range(n))
    b = dp.local_batch(B, world)
    for r in range(world):
        for m in range(n // B):
            assert rows[r][m * b:(m + 1) * b].tolist() == list(range(m * B + r * b, m * B + (r + 1) * b))
    assert dp.local_rows(64, 32, 0, 1).tolist() == list(range(64))
    import pytest
    with pytest.raises(ValueError):
        dp.local_batch(30, 4)
    with pytest.raises(ValueError):
        dp.local_rows(50, 32, 0, 2)


def test_plan_buckets_suffix_order():
    """arena of 6 slots (offsets 0,10,20,30,40,50; 60 floats); the backward pass issues them last-to-first, one
    slot out of order: a bucket may only cover a fully issued suffix; the FC tail is cut at once; the bucket that
    reaches offset 0 trails"""
    offs = [0, 10, 20, 30, 40, 50]
    events = [([50], False), ([40], True),        # FC tail = slots 40, 50 -> forced cut after step 1
              ([20], False),                      # out of order: 30 still missing, nothing to cut
              ([30], False),                      # suffix is now [20, 40): 20 floats >= 15 -> cut
              ([10], False), ([0], False)]
    cuts, trailing = dp.plan_buckets(offs, 60, events, 15)
    assert cuts == {1: (40, 60), 3: (20, 40)}
    assert trailing == (0, 20)
    covered = sorted(list(cuts.values()) + [trailing])
    assert covered[0][0] == 0 and covered[-1][1] == 60 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def _worker_syncbn(rank, world, port, ret):
    """SyncBN arithmetic on CPU: per-rank fp64 {sum, sumsq} summed over the ranks give the statistics of the
    global batch; per-rank backward sums likewise; BN parameter gradients scaled by 1/world survive the SUM
    all-reduce + 1/world of the optimiser exactly once (engine.py: _sync_stats, pscale)"""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.RandomState(3)
    x = rng.randn(8, 5)                                  # global batch of 8 "pixels", 5 channels
    dz = rng.randn(8, 5)
    lo, hi = dp.shard_range(8, rank, world)
    st = torch.from_numpy(np.concatenate([x[lo:hi].sum(0), (x[lo:hi] ** 2).sum(0)]))
    dist.all_reduce(st)                                  # what stats_allreduce_fn does to STATS[f0:f0+2C]
    mean = st[:5] / 8
    var = st[5:] / 8 - mean ** 2
    xhat = (torch.from_numpy(x[lo:hi]) - mean) / torch.sqrt(var + 1e-4)
    bst = torch.cat([torch.from_numpy(dz[lo:hi]).sum(0), (torch.from_numpy(dz[lo:hi]) * xhat).sum(0)])
    dist.all_reduce(bst)
    dgamma = bst[5:] * (1.0 / world)                     # pscale
    dist.all_reduce(dgamma)                              # gradient arena SUM
    ret[rank] = (mean.numpy().copy(), var.numpy().copy(), dgamma.numpy().copy())
    dist.destroy_process_group()


def test_syncbn_statistics_equal_global_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_syncbn, args=(world, 29573, ret), nprocs=world, join=True)
    rng = np.random.RandomState(3)
    x = rng.randn(8, 5)
    dz = rng.randn(8, 5)
    xhat = (x - x.mean(0)) / np.sqrt(x.var(0) + 1e-4)
    for r in range(world):
        mean, var, dgamma = ret[r]
        assert np.allclose(mean, x.mean(0), atol=1e-12) and np.allclose(var, x.var(0), atol=1e-12)
        assert np.allclose(dgamma, (dz * xhat).sum(0), atol=1e-10)     # the optimiser's 1/world is applied to all grads alike
