"""Experiment (not a test): probe MN-major descriptor variants of the experimental wgrad kernel."""
import os, sys, itertools, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(__file__))
import conftest  # noqa
from test_gpu_conv_tc import _setup, P
from dpp_b200.lib import lib

shape = (2, 32, 64, 64, 1, 1)
N, H, Cin, Cout, k, stride = shape
ref = None
for var in [('0', 0, 0, 0)] + [('1', l, s, ks) for l, s, ks in [(256, 32, 1024), (32, 256, 1024), (256, 64, 1024), (256, 32, 512), (1, 32, 1024), (256, 8, 1024)]]:
    os.environ['DPP_WGRAD_MN'] = var[0]
    os.environ['DPP_MN_LBO'], os.environ['DPP_MN_SBO'], os.environ['DPP_MN_KSTEP'] = str(var[1]), str(var[2]), str(var[3])
    print('running', var, flush=True)
    d, x, w, bias, bn, Ho, keep = _setup(N, H, Cin, Cout, k, stride, 2)
    g = torch.Generator(device='cuda').manual_seed(11)
    dy = torch.randn(N, Ho, Ho, Cout, device='cuda', generator=g)
    dw = torch.zeros(k * k * Cin, Cout, device='cuda'); db = torch.zeros(Cout, device='cuda')
    lib.dpp_conv2d_wgrad(C.byref(d), P(x), C.byref(bn), P(dy), P(dw), P(db), None)
    torch.cuda.synchronize()
    out = dw.cpu().numpy()
    if ref is None:
        ref = out
        print("ref absmax", np.abs(ref).max())
        continue
    err = np.abs(out - ref).max() / np.abs(ref).max()
    print("variant", var, "rel err %.3g" % err, "out absmax %.3g" % np.abs(out).max(), flush=True)
