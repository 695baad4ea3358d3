"""Kernel-level GPU tests through the C ABI for ConvPoolLayer (stem specialisation and the
generic PoseRegNet shapes) and HiddenLayer (SIMT and tcgen05 GEMM), against plain torch fp32 ops."""
import ctypes as C
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False        # the torch references must be fp32, not TF32
torch.backends.cuda.matmul.allow_tf32 = False

from dpp_b200.lib import lib  # noqa: E402


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _kc(w):      # (O,I,kh,kw) -> KC flipped
    return w.flip(2, 3).permute(2, 3, 1, 0).contiguous().reshape(-1, w.shape[0])


@pytest.mark.parametrize("cfg", [  # N, H, Cin, Cout, k, pad, pool, relu
    (3, 128, 1, 32, 5, 2, 2, 0),      # ResNet stem (specialised kernels)
    (128, 128, 1, 32, 5, 2, 2, 0),    # ... at the benchmarked batch: many regions per persistent CTA, double-buffered patches
    (2, 120, 1, 32, 5, 2, 2, 0),      # ... pooled size 60: partial 8 x 16 / 8 x 8 regions at the right and bottom edges
    (5, 24, 1, 32, 5, 2, 2, 0),       # ... smaller than one region
    (2, 128, 1, 8, 5, 0, 4, 1),       # PoseRegNet layer 0
    (2, 31, 8, 8, 5, 0, 2, 1),        # PoseRegNet layer 1
    (2, 13, 8, 8, 3, 0, 1, 1),        # PoseRegNet layer 2
])
def test_convpool_fwd_bwd_vs_torch(cfg):
    N, H, Cin, Cout, k, pad, pool, relu = cfg
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(N, Cin, H, H, device='cuda', generator=g)
    x[:, :, :(9 if H > 24 else 1)] = 1.0                  # flat region: exact ties in the pool windows
    w = (torch.randn(Cout, Cin, k, k, device='cuda', generator=g) * 0.2).requires_grad_(True)
    b = (torch.randn(Cout, device='cuda', generator=g) * 0.1).requires_grad_(True)
    xr = x.clone().requires_grad_(Cin > 1)
    o = F.conv2d(xr, w.flip(2, 3), padding=pad)
    if pool > 1:
        o = F.max_pool2d(o, pool, pool)
    o = o + b.view(1, -1, 1, 1)
    if relu:
        o = torch.relu(o)
    go = torch.randn(o.shape, device='cuda', generator=g)
    o.backward(go)
    Hp = o.shape[2]
    xn = x.permute(0, 2, 3, 1).contiguous()
    wk = _kc(w.detach())
    y = torch.zeros(N, Hp, Hp, Cout, device='cuda')
    am = torch.zeros(N, Hp, Hp, Cout, dtype=torch.uint8, device='cuda')
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device='cuda')
    lib.dpp_convpool_fwd(P(xn), P(wk), P(b.detach()), P(y), P(am), P(stats), N, H, H, Cin, Cout, k, pad, pool, relu, None)
    yr = o.detach().permute(0, 2, 3, 1)
    assert torch.allclose(y, yr, rtol=1e-4, atol=1e-4), float((y - yr).abs().max())
    assert np.allclose(stats[:Cout].cpu().numpy(), yr.double().sum((0, 1, 2)).cpu().numpy(), rtol=1e-6, atol=1e-3)
    dw = torch.zeros_like(wk)
    db = torch.zeros(Cout, device='cuda')
    dx = torch.zeros_like(xn) if Cin > 1 else None
    gon = go.permute(0, 2, 3, 1).contiguous()
    lib.dpp_convpool_bwd(P(xn), P(wk), P(y), P(am), P(gon), P(dw), P(db), P(dx), N, H, H, Cin, Cout, k, pad, pool, relu, None)
    torch.cuda.synchronize()
    dwr = _kc(w.grad)
    err = (dw - dwr).abs().max() / dwr.abs().max()
    print(cfg, "dW rel err", float(err), "db err", float((db - b.grad).abs().max() / b.grad.abs().max()))
    if err > 1e-3:
        d = ((dw - dwr).abs() / dwr.abs().max()).cpu().numpy()
        print("per-tap max err", d.max(axis=1).round(4)[:25])
    assert err < 1e-3
    assert torch.allclose(db, b.grad, rtol=1e-3, atol=1e-3 * float(b.grad.abs().max()))
    if Cin > 1:
        dxr = xr.grad.permute(0, 2, 3, 1)
        assert (dx - dxr).abs().max() / dxr.abs().max() < 1e-3


@pytest.mark.parametrize("precision,tol", [(0, 1e-5), (1, 2e-5), (2, 4e-3)])
@pytest.mark.parametrize("dims", [(128, 16384, 1024, 1), (128, 1024, 1024, 1), (4, 1024, 32, 0), (128, 968, 1024, 1),
                                  (128, 1024, 30, 0)])
def test_fc_fwd_bwd_vs_torch(dims, precision, tol):
    B, n_in, n_out, relu = dims
    g = torch.Generator(device='cuda').manual_seed(4)
    x = torch.randn(B, n_in, device='cuda', generator=g, requires_grad=True)
    w = (torch.randn(n_in, n_out, device='cuda', generator=g) * (1.0 / n_in) ** 0.5).requires_grad_(True)
    b = (torch.randn(n_out, device='cuda', generator=g) * 0.1).requires_grad_(True)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    o = x @ w + b
    if relu:
        o = torch.relu(o)
    go = torch.randn(o.shape, device='cuda', generator=g)
    o.backward(go)
    torch.backends.cuda.matmul.allow_tf32 = prev
    y = torch.zeros(B, n_out, device='cuda')
    lib.dpp_fc_fwd(P(x.detach()), P(w.detach()), P(b.detach()), P(y), B, n_in, n_out, relu, None, 1.0, precision, None)
    e = float((y - o.detach()).abs().max() / o.detach().abs().max())
    dw = torch.zeros(n_in, n_out, device='cuda')
    db = torch.zeros(n_out, device='cuda')
    dx = torch.zeros(B, n_in, device='cuda')
    scratch = torch.zeros(B, n_out, device='cuda')
    lib.dpp_fc_bwd(P(x.detach()), P(w.detach()), P(o.detach().contiguous()), P(go), P(dw), P(db), P(dx), P(scratch), B, n_in, n_out,
                   relu, None, 1.0, precision, None)
    torch.cuda.synchronize()
    ew = float((dw - w.grad).abs().max() / w.grad.abs().max())
    ex = float((dx - x.grad).abs().max() / x.grad.abs().max())
    print(dims, precision, "fwd", e, "dW", ew, "dx", ex)
    assert e < tol and ew < tol and ex < tol
    assert torch.allclose(db, b.grad, rtol=1e-4, atol=1e-4 * float(b.grad.abs().max()))


def test_adam_matches_oracle_on_identical_gradients():
    """trainer/optimizer.py:58-90 (ADAM v2, fp32 constants): same gradients in, same weights out -
    three steps, including zero and tiny gradients, against oracle.nets.Adam (numpy fp32)."""
    from oracle import nets as O
    rng = np.random.RandomState(3)
    n = 200003
    w0 = rng.randn(n).astype(np.float32)
    op = torch.from_numpy(w0.copy())
    oadam = O.Adam([op])
    w = torch.from_numpy(w0).cuda()
    m = torch.zeros_like(w); v = torch.zeros_like(w)
    hyper = torch.tensor([0.0, 1.0, 0.0, 1.0], device='cuda')
    for step, lr in enumerate([1e-4, 3.3e-4, 1e-3]):
        g = (rng.randn(n) * 10.0 ** rng.uniform(-9, 1, n)).astype(np.float32)
        g[::97] = 0.0
        hyper[0:1].fill_(lr)
        gd = torch.from_numpy(g).cuda()
        lib.dpp_adam_step(P(w), P(gd), P(m), P(v), P(hyper), n, None)
        lib.dpp_adam_tick(P(hyper), None)
        oadam.step([torch.from_numpy(g)], lr)
        torch.cuda.synchronize()
        a, b = w.cpu().numpy(), op.numpy()
        upd = np.abs(b - w0).max()
        err = np.abs(a - b).max()
        print("adam step", step, "max |dw| diff", err, "max update", upd)
        assert err <= 4e-7 * max(1.0, np.abs(b).max())      # a few ulp of w (pow/sqrt/div rounding)
        assert np.abs(m.cpu().numpy() - oadam.m[0]).max() <= 1e-6 * np.abs(oadam.m[0]).max()
        assert np.abs(v.cpu().numpy() - oadam.v[0]).max() <= 1e-6 * np.abs(oadam.v[0]).max()
