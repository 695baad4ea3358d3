#!/usr/bin/env python
"""Data-parallel consistency check, run under torchrun on N >= 2 GPUs:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
Each rank trains the batch-B ResNet on its own shard of a global batch for 3 steps (CUDA graph, NCCL all-reduce of
the gradient arena, FC tail reduced early on the comm stream).  Checks:
  (a) all ranks hold bit-identical weights afterwards (replicas stay in sync);
  (b) the weights equal those of a run that all-reduces the whole arena after the backward pass
      (DPP_EARLY_ALLREDUCE=0): the overlap changes the schedule, not the arithmetic."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-prior-pp_b200'))
from dpp_b200.engine import Engine  # noqa: E402
from net.resnet import ResNet, ResNetParams  # noqa: E402


def run(early, rank, world, B=16, steps=3):
    os.environ['DPP_EARLY_ALLREDUCE'] = '1' if early else '0'
    net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=B, numJoints=1, nDims=30))
    eng = Engine(net, precision=1)
    eng._alloc_training()
    eng.set_world(world, lambda g: dist.all_reduce(g))
    eng.set_lr(1e-3)
    g = torch.Generator(device='cuda').manual_seed(100 + rank)
    for s in range(steps):
        eng.t_in.buf.copy_(torch.randn(eng.t_in.buf.shape, device='cuda', generator=g))
        eng.y_in.copy_(torch.randn(eng.y_in.shape, device='cuda', generator=g))
        eng.train_step(None, use_graph=True)
    torch.cuda.synchronize()
    w = eng.W.clone()
    eng.release()
    return w


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    w_early = run(True, rank, world)
    w_late = run(False, rank, world)
    lo, hi = w_early.clone(), w_early.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same_ranks = bool(torch.equal(lo, hi))
    diff = float((w_early - w_late).abs().max())
    scale = float(w_late.abs().max())
    if rank == 0:
        print("dp_check world=%d replicas_identical=%s early_vs_late_max_abs_diff=%.3e (|w|max %.3e)" % (
            world, same_ranks, diff, scale))
        ok = same_ranks and diff <= 1e-6 * max(scale, 1.0)
        print("DP_CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
