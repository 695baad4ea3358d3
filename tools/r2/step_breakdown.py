"""Timing experiment (GPU): where the batch-128 training step spends its time.  Times, with CUDA events over graph
replays: the whole step, the step with backward-weights on the main stream (no second stream), the step without
backward-weights kernels (main chain alone), and the forward pass alone."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'deep-prior-pp_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from dpp_b200.engine import Engine
from dpp_b200.lib import lib
from net.resnet import ResNet, ResNetParams

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128


def build():
    net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=B, numJoints=1, nDims=30))
    eng = Engine(net, precision=1)
    eng._alloc_training()
    eng.t_in.buf.normal_()
    eng.y_in.normal_()
    eng.set_lr(1e-4)
    return eng


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eng = build()
print("B=%d whole step (2 streams): %.3f ms" % (B, timeit(lambda: eng.train_step(None))))
eng._wgrad_stream = None
eng._graphs.clear()
print("backward-weights on the main stream: %.3f ms" % timeit(lambda: eng.train_step(None)))
real = lib.dpp_conv2d_wgrad
import dpp_b200.engine as E


class NoWgrad(object):
    def __getattr__(self, name):
        if name == 'dpp_conv2d_wgrad':
            return lambda *a: 0
        return getattr(lib, name)


E.lib = NoWgrad()
eng._graphs.clear()
print("without backward-weights kernels: %.3f ms" % timeit(lambda: eng.train_step(None)))
E.lib = lib
g = torch.cuda.CUDAGraph()
eng.forward_device(deterministic=False)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    eng.forward_device(deterministic=False)
print("forward only (train mode): %.3f ms" % timeit(lambda: g.replay()))
