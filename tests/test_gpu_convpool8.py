"""The specialised 8-filter conv+pool kernels (csrc/convpool8.cu, the default path; DPP_CONVPOOL_FAST=0 selects the
generic kernel): outputs and arg-max cells must equal the generic kernel's bit for bit (same accumulation order), for
every layer shape of the ScaleNet / PoseRegNet towers.  The generic kernel is anchored to torch fp32 in
test_gpu_convpool_fc.py."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [(128, 1, 5, 4), (31, 8, 5, 2), (13, 8, 3, 1), (64, 1, 5, 2), (30, 8, 5, 2), (32, 1, 5, 2), (14, 8, 5, 1),
         (10, 8, 3, 1)]


@pytest.mark.parametrize('H,Cin,k,pool', CASES)
def test_fast_convpool_equals_generic(H, Cin, k, pool, monkeypatch):
    from dpp_b200.lib import lib
    N, Cout = 5, 8
    rng = np.random.RandomState(H * 31 + Cin)
    x = torch.from_numpy(rng.uniform(-1, 1, (N, H, H, Cin)).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.randn(k * k * Cin, Cout) * 0.3).astype(np.float32)).cuda()
    b = torch.from_numpy((rng.randn(Cout) * 0.1).astype(np.float32)).cuda()
    Hp = (H - k + 1) // pool
    outs = []
    for fast in ('0', '1'):
        monkeypatch.setenv('DPP_CONVPOOL_FAST', fast)
        y = torch.full((N, Hp, Hp, Cout), -7., device='cuda')
        am = torch.full((N, Hp, Hp, Cout), 255, dtype=torch.uint8, device='cuda')
        P = lambda t: C.c_void_p(t.data_ptr())
        lib.dpp_convpool_fwd(P(x), P(w), P(b), P(y), P(am), None, N, H, H, Cin, Cout, k, 0, pool, 1, None)
        torch.cuda.synchronize()
        outs.append((y.cpu().numpy(), am.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
