"""Diagnostic (GPU): gradient parity of one batch-B train step, tcgen05 (precision 1) vs fp32 SIMT (precision 0)
vs the float64 oracle with the engine's ReLU decisions imposed - per parameter tensor, forward order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'deep-prior-pp_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
from test_gpu_resnet import _build, _data, _engine_relu_masks
from oracle import nets as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
D = 30
x, y = _data(B, D)
res = {}
for prec in (1, 0):
    net, onet, eng = _build(0, B, 1, D, precision=prec)
    eng.set_input_nchw(x)
    eng._alloc_training()
    eng.y_in.copy_(torch.from_numpy(y))
    w_before = eng.W.clone()
    cost = float(eng.train_step(1e-3, use_graph=False).cpu()[0])
    grads = eng.gradients()
    masks = _engine_relu_masks(net, onet, eng, w_before)
    res[prec] = (net, onet, cost, grads, masks)
    # intermediate gradients of interest
    print("precision", prec, "cost", cost)
net1, onet, c1, g1, m1 = res[1]
net0, _, c0, g0, m0 = res[0]
ndiff = sum(int((m1[k] != m0[k]).sum()) for k in m1)
print("ReLU decisions differing between precision 1 and 0:", ndiff)
adam = O.Adam(onet.params)
ocost, oout, ograds = O.train_step(onet, adam, torch.from_numpy(x), torch.from_numpy(y), 1e-3, 1, D, relu_masks=m1)
print("oracle cost", ocost)
lays = [l for l in onet.layers for _ in l.params]
print("%-12s %-9s %10s %10s %10s" % ("name", "kind", "p1-vs-p0", "p1-vs-orc", "p0-vs-orc"))
for pa, pb, og, l in zip(net1.params, net0.params, ograds, lays):
    a, b, o = g1[id(pa)], g0[id(pb)], og.numpy()
    if l.kind in ('conv', 'convpool') and a.ndim == 1:
        continue
    n = np.linalg.norm(o.ravel()) + 1e-30
    print("%-12s %-9s %10.3g %10.3g %10.3g" % (pa.name, l.kind, np.linalg.norm((a - b).ravel()) / n,
                                                  np.linalg.norm((a - o).ravel()) / n, np.linalg.norm((b - o).ravel()) / n))
