#!/usr/bin/env python
"""Data-parallel consistency check, run under torchrun on N >= 2 GPUs:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
Every rank runs ONE training step (CUDA graph) of the batch-B ResNet on its own shard, three ways:
  local : world = 1, no exchange            -> G_local (this rank's gradient arena)
  early : FC tail all-reduced on the comm stream underneath the conv backward, the rest at the end (default)
  late  : one all-reduce of the whole arena after the backward pass (DPP_EARLY_ALLREDUCE=0)
Checks (gradients, not post-ADAM weights: ADAM's first steps turn roundoff-level gradients - the conv biases have
mathematically zero gradient - into +-lr updates, so weights are not a usable comparison):
  (a) G_early == sum over ranks of G_local            (the exchange computes the global sum)
  (b) G_early == G_late                               (the overlap changes the schedule, not the arithmetic)
  (c) all ranks hold identical G and identical weights after the step (replicas stay in sync)
Tolerance 2e-5 of max|G|: the backward-weights kernels accumulate with floating-point atomics, whose order varies."""
import os
import sys
import threading

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-prior-pp_b200'))
from dpp_b200.engine import Engine  # noqa: E402
from net.resnet import ResNet, ResNetParams  # noqa: E402


def run(mode, rank, world, B=16):
    os.environ['DPP_EARLY_ALLREDUCE'] = '0' if mode == 'late' else '1'
    net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=B, numJoints=1, nDims=30))
    eng = Engine(net, precision=1)
    eng._alloc_training()
    if mode != 'local':
        eng.set_world(world, lambda g: dist.all_reduce(g))
    eng.set_lr(1e-3)
    g = torch.Generator(device='cuda').manual_seed(100 + rank)
    eng.t_in.buf.copy_(torch.randn(eng.t_in.buf.shape, device='cuda', generator=g))
    eng.y_in.copy_(torch.randn(eng.y_in.shape, device='cuda', generator=g))
    eng.train_step(None, use_graph=True)
    torch.cuda.synchronize()
    G, W = eng.G.clone(), eng.W.clone()
    eng._graphs.clear()
    eng.release()
    return G, W


def same_on_all_ranks(t):
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    g_local, _ = run('local', rank, world)
    g_sum = g_local.clone()
    dist.all_reduce(g_sum)
    g_early, w_early = run('early', rank, world)
    g_late, _ = run('late', rank, world)
    scale = float(g_sum.abs().max())
    d_sum = float((g_early - g_sum).abs().max()) / scale
    d_late = float((g_early - g_late).abs().max()) / scale
    sync_g, sync_w = same_on_all_ranks(g_early), same_on_all_ranks(w_early)
    if rank == 0:
        print("dp_check world=%d  |G_early - sum(G_local)|/max|G| = %.2e   |G_early - G_late|/max|G| = %.2e   "
              "replicas identical: G %s, W %s" % (world, d_sum, d_late, sync_g, sync_w))
        ok = d_sum < 2e-5 and d_late < 2e-5 and sync_g and sync_w
        print("DP_CHECK", "PASS" if ok else "FAIL")
        sys.stdout.flush()
    timer = threading.Timer(30.0, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    timer.cancel()


if __name__ == '__main__':
    main()
