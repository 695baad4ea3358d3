"""Data-parallel plumbing (not in the reference, which is single-device: SURVEY 2.1).
One process per GPU; the minibatch shards over ranks, each rank holds a full weight replica, and
the flat gradient arena is summed with one bucketed all-reduce per step; 1/world is folded into
ADAM (hyper[3]).  torch.distributed is the transport (NCCL on GPUs, gloo in CPU tests)."""
import numpy as np


def shard_range(n_samples, rank, world):
    """contiguous shard [lo, hi) of a batch / data set for this rank"""
    per = (n_samples + world - 1) // world
    lo = min(rank * per, n_samples)
    return lo, min(lo + per, n_samples)


def bucket_bounds(n_elems, bucket_elems):
    """split the flat gradient arena into buckets, LAST elements first: the FC tail sits at the end
    of the arena and its gradients are the first ones complete in the backward pass"""
    out = []
    hi = n_elems
    while hi > 0:
        lo = max(0, hi - bucket_elems)
        out.append((lo, hi))
        hi = lo
    return out


def make_allreduce(dist, group=None, bucket_elems=8 * 1024 * 1024):
    """returns fn(flat_grad_tensor): in-place SUM over ranks, bucket by bucket"""
    def fn(g):
        for lo, hi in bucket_bounds(g.numel(), bucket_elems):
            dist.all_reduce(g[lo:hi], group=group)
    return fn
