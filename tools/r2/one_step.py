"""n eager training steps of the ResNet (batch argv[2], default 128) incl. the augmentation kernel - the command the
profiler / sanitizer runs of round 2 wrap (tools/r2/call13.sh).  Eager (no CUDA graph) so that ncu sees plain launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'deep-prior-pp_b200')):
    sys.path.insert(0, p)
import ctypes as C
import numpy as np
import torch
import bench
from dpp_b200.engine import Engine
from dpp_b200.lib import lib
from net.resnet import ResNet, ResNetParams

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
bench.N_RESIDENT = max(B, 256)
ds, comp, mean = bench.make_workload()
net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=B, numJoints=1, nDims=30))
eng = Engine(net, precision=1)
eng._alloc_training()
eng.set_lr(1e-4)
rng = np.random.RandomState(5)
crops = torch.from_numpy(ds['x'][:, 0].copy()).cuda()
for s in range(steps):
    idxs = rng.randint(0, bench.N_RESIDENT, B)
    r, y, _ = bench.records_for(ds, comp, mean, idxs, rng)
    rd = torch.from_numpy(r.view(np.uint8).reshape(len(r), -1).copy()).cuda()
    lib.dpp_augment_fwd(C.c_void_p(crops.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(eng.t_in.buf.data_ptr()), B, 128, 128, None)
    eng.y_in.copy_(torch.from_numpy(y).cuda())
    cost = float(eng.train_step(None, use_graph=False).cpu()[0])
    print("step", s, "cost", cost)
eng.check_barriers()
print("done")
