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
  (d) the weights after the step are one reference ADAM step from the weights before it, over the WHOLE arena (the
      optimiser pass is split so that its first part hides the exchange of the trailing bucket)
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


def inputs_of(rank, B):
    g = torch.Generator(device='cuda').manual_seed(100 + rank)
    x = torch.randn((B, 128, 128, 1), device='cuda', generator=g)
    y = torch.randn((B, 30), device='cuda', generator=g)
    return x, y


def run(mode, rank, world, B=16):
    """mode: local | early | late | syncbn (DP with BatchNorm statistics summed over the ranks) | full (ONE device
    computing the global batch of world * B samples: what the single-device reference does)"""
    os.environ['DPP_EARLY_ALLREDUCE'] = '0' if mode == 'late' else '1'
    nb = B * world if mode == 'full' else B
    net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=nb, numJoints=1, nDims=30))
    eng = Engine(net, precision=1)
    eng._alloc_training()
    if mode not in ('local', 'full'):
        eng.set_world(world, lambda g: dist.all_reduce(g), rank=rank,
                      syncbn={'syncbn': True, 'syncbn_p2p': 'p2p'}.get(mode, False), dist=dist)
    eng.set_lr(1e-3)
    if mode == 'full':
        xs, ys = zip(*[inputs_of(r, B) for r in range(world)])
        x, y = torch.cat(xs), torch.cat(ys)
    else:
        x, y = inputs_of(rank, B)
    eng.t_in.buf.copy_(x)
    eng.y_in.copy_(y)
    w_before = eng.W.clone()
    cost = eng.train_step(None, use_graph=True)
    torch.cuda.synchronize()
    eng.check_barriers()
    G, W, R, c = eng.G.clone(), eng.W.clone(), eng.R.clone(), float(cost.cpu()[0])
    run.last_w_before, run.last_hyper = w_before, eng.hyper.clone()
    eng._graphs.clear()
    eng.release()
    return G, W, R, c


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
    g_local = run('local', rank, world)[0]
    g_sum = g_local.clone()
    dist.all_reduce(g_sum)
    g_early, w_early = run('early', rank, world)[:2]
    # (d) the optimiser pass is split in data-parallel runs (everything above the trailing bucket first, the trailing
    # bucket after its exchange): every parameter must have received exactly ONE first ADAM step with the exchanged
    # gradient (trainer/optimizer.py:69-88 at t = 1: m = 0.1 g, v = 0.001 g^2, bias corrections 0.1 / 0.001)
    w0, hyper = run.last_w_before, run.last_hyper
    lr, gs = float(hyper[0]), float(hyper[3])
    g1 = g_early * gs
    c1 = 1.0 - torch.tensor(0.9, dtype=torch.float32).pow(1.0)
    c2 = 1.0 - torch.tensor(0.999, dtype=torch.float32).pow(1.0)
    mh, vh = (0.1 * g1) / c1.item(), (0.001 * (g1 * g1)) / c2.item()
    w_ref = w0 - (lr * mh) / (vh.sqrt() + 1e-8)
    d_adam = float((w_early - w_ref).abs().max()) / lr
    g_late = run('late', rank, world)[0]
    scale = float(g_sum.abs().max())
    d_sum = float((g_early - g_sum).abs().max()) / scale
    d_late = float((g_early - g_late).abs().max()) / scale
    sync_g, sync_w = same_on_all_ranks(g_early), same_on_all_ranks(w_early)
    # SyncBN: world ranks x B samples against ONE device with the global batch (the single-device reference,
    # net/batchnormlayer.py:154-159).  The exchanged arena holds the SUM of the per-rank gradients of the per-rank
    # MEAN costs = world x the gradient of the global mean cost (ADAM folds the 1/world in).
    g_sbn, w_sbn, r_sbn, c_sbn = run('syncbn', rank, world)
    g_full, w_full, r_full, c_full = run('full', rank, world)
    c_mean = torch.tensor([c_sbn], dtype=torch.float64, device='cuda')
    dist.all_reduce(c_mean)
    c_mean = float(c_mean[0]) / world
    gs = float(g_full.abs().max())
    d_sbn = float((g_sbn / world - g_full).abs().max()) / gs
    l2_sbn = float((g_sbn / world - g_full).norm() / g_full.norm())
    d_run = float((r_sbn - r_full).abs().max() / r_full.abs().max())
    d_cost = abs(c_mean - c_full) / abs(c_full)
    sync_sbn = same_on_all_ranks(g_sbn) and same_on_all_ranks(w_sbn) and same_on_all_ranks(r_sbn)
    # the same with the statistics exchanged over peer memory (dpp_stats_exchange) instead of NCCL
    g_p2p, w_p2p, r_p2p, c_p2p = run('syncbn_p2p', rank, world)
    d_p2p = float((g_p2p / world - g_full).abs().max()) / gs
    l2_p2p = float((g_p2p / world - g_full).norm() / g_full.norm())
    d_run_p2p = float((r_p2p - r_full).abs().max() / r_full.abs().max())
    sync_p2p = same_on_all_ranks(g_p2p) and same_on_all_ranks(w_p2p) and same_on_all_ranks(r_p2p)
    if rank == 0:
        print("dp_check world=%d  |G_early - sum(G_local)|/max|G| = %.2e   |G_early - G_late|/max|G| = %.2e   "
              "replicas identical: G %s, W %s   split ADAM vs one reference step: max |dW| = %.1e lr"
              % (world, d_sum, d_late, sync_g, sync_w, d_adam))
        print("dp_check SyncBN vs one device with the global batch of %d: cost %.3e, gradient max %.2e / rel-L2 %.2e, "
              "running statistics %.2e, replicas identical %s" % (16 * world, d_cost, d_sbn, l2_sbn, d_run, sync_sbn))
        ok = d_sum < 2e-5 and d_late < 2e-5 and sync_g and sync_w and d_adam < 1e-3      # a missed or doubled range would be ~1 lr
        # a handful of roundoff-level ReLU decisions may differ between the two runs (different summation order of the
        # statistics): bound the gradient by what such flips allow, the cost and statistics tightly
        ok_sbn = d_cost < 1e-5 and d_run < 1e-5 and l2_sbn < 2e-2 and sync_sbn
        print("dp_check SyncBN over peer memory vs one device: gradient max %.2e / rel-L2 %.2e, running statistics %.2e, "
              "replicas identical %s" % (d_p2p, l2_p2p, d_run_p2p, sync_p2p))
        ok_p2p = d_run_p2p < 1e-5 and l2_p2p < 2e-2 and sync_p2p
        print("DP_CHECK", "PASS" if (ok and ok_sbn and ok_p2p) else "FAIL")
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
