"""Data-parallel plumbing (not in the reference, which is single-device: SURVEY 2.1).
One process per GPU; the minibatch shards over ranks, each rank holds a full weight replica, and
the flat gradient arena is summed with one bucketed all-reduce per step; 1/world is folded into
ADAM (hyper[3]).  torch.distributed is the transport (NCCL on GPUs, gloo in CPU tests)."""
import os
import numpy as np


def env_world():
    """(rank, world, local_rank) as torchrun exports them; (0, 1, 0) in a plain process"""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def init_process_group(backend=None):
    """Join the job torchrun started (idempotent).  Returns (dist module or None, rank, world).  One process per
    GPU: the CUDA device is LOCAL_RANK; backend NCCL on GPUs, gloo in the CPU tests."""
    rank, world, local = env_world()
    if world <= 1:
        return None, 0, 1
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return dist, dist.get_rank(), dist.get_world_size()


def local_batch(batch, world):
    """samples of one global minibatch that a rank computes (strong scaling: the reference's batch size stays the
    GLOBAL one, e.g. BASELINE config 3: 512 over 8 GPUs = 64 per GPU)"""
    if batch % world:
        raise ValueError("batch size %d does not divide over %d ranks" % (batch, world))
    return batch // world


def local_rows(n_aligned, batch, rank, world):
    """rows of the (minibatch-aligned) data set that live on this rank: slice [rank*b, (rank+1)*b) of every global
    minibatch, in minibatch order - local minibatch m is rows [m*b, (m+1)*b) of the result"""
    if n_aligned % batch:
        raise ValueError("data set of %d rows is not aligned to the batch size %d" % (n_aligned, batch))
    b = local_batch(batch, world)
    m = np.arange(n_aligned // batch, dtype=np.int64)[:, None] * batch
    return (m + rank * b + np.arange(b, dtype=np.int64)[None, :]).reshape(-1)


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


def plan_buckets(slot_offsets, n_elems, events, min_elems):
    """Exchange plan of the flat gradient arena for the backward pass.

    The arena holds the parameter slots in layer order (``slot_offsets``, ``n_elems`` floats in total); the
    backward pass fills it from the END: ``events[i] = (offsets of the slots whose gradient kernels have been issued
    once step i of the reverse walk is done, force)``.  After every step the longest fully issued SUFFIX [lo, n) of
    the arena is known; a bucket [lo, hi) is cut there as soon as it holds >= ``min_elems`` floats, or at once when
    ``force`` is set (end of the FC tail: 90 % of the bytes, complete when the backward pass has barely begun).
    The bucket that reaches offset 0 is returned separately: it trails the last backward kernel.
    Returns ({step index: (lo, hi)}, (0, hi) or None)."""
    order = sorted(slot_offsets)
    done = set()
    cuts, hi, k = {}, n_elems, len(order)
    for i, (offs, force) in enumerate(events):
        done.update(offs)
        while k > 0 and order[k - 1] in done:
            k -= 1
        lo = order[k] if 0 < k < len(order) else (0 if k == 0 else n_elems)
        if lo > 0 and hi > lo and (hi - lo >= min_elems or force):
            cuts[i] = (lo, hi)
            hi = lo
    return cuts, ((0, hi) if hi > 0 else None)
