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
    """Forward against float64 torch (conv2d + max_pool2d + bias + ReLU).  Backward against float64 torch autograd of
    the convolution, with the pooling / ReLU routing taken from the kernel's OWN arg-max cells - after checking that
    each of those cells holds the window's maximum: among 16.7 M pooled outputs at batch 128 a few near-ties flip
    between any two summation orders, and one flipped cell is a 1e-3 change of dW, so comparing against torch's own
    choice would test the rounding of its convolution, not this kernel."""
    N, H, Cin, Cout, k, pad, pool, relu = cfg
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(N, Cin, H, H, device='cuda', generator=g)
    x[:, :, :min(9, H // 3)] = 1.0                        # flat region: exact ties in the pool windows
    w = (torch.randn(Cout, Cin, k, k, device='cuda', generator=g) * 0.2)
    b = (torch.randn(Cout, device='cuda', generator=g) * 0.1)
    xd = x.double().requires_grad_(Cin > 1)
    wd = w.double().requires_grad_(True)
    conv = F.conv2d(xd, wd.flip(2, 3), padding=pad)                       # [N, Cout, Hc, Wc] float64
    o = F.max_pool2d(conv, pool, pool) if pool > 1 else conv
    o = o + b.double().view(1, -1, 1, 1)
    if relu:
        o = torch.relu(o)
    Hp = o.shape[2]
    go = torch.randn(o.shape, device='cuda', generator=g)
    xn = x.permute(0, 2, 3, 1).contiguous()
    wk = _kc(w)
    y = torch.zeros(N, Hp, Hp, Cout, device='cuda')
    am = torch.zeros(N, Hp, Hp, Cout, dtype=torch.uint8, device='cuda')
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device='cuda')
    lib.dpp_convpool_fwd(P(xn), P(wk), P(b), P(y), P(am), P(stats), N, H, H, Cin, Cout, k, pad, pool, relu, None)
    yr = o.detach().float().permute(0, 2, 3, 1)
    assert torch.allclose(y, yr, rtol=1e-4, atol=1e-4), float((y - yr).abs().max())
    assert np.allclose(stats[:Cout].cpu().numpy(), yr.double().sum((0, 1, 2)).cpu().numpy(), rtol=1e-6, atol=1e-3)
    # the kernel's arg-max cell (cy * pool + cx inside the window) -> flat index into the convolution output
    amc = am.permute(0, 3, 1, 2).long()                                   # [N, Cout, Hp, Hp]
    ph = torch.arange(Hp, device='cuda').view(1, 1, Hp, 1)
    pw = torch.arange(Hp, device='cuda').view(1, 1, 1, Hp)
    Wc = conv.shape[3]
    flat = (ph * pool + amc // pool) * Wc + (pw * pool + amc % pool)
    picked = conv.detach().flatten(2).gather(2, flat.flatten(2)).view_as(amc)
    pooled = (F.max_pool2d(conv.detach(), pool, pool) if pool > 1 else conv.detach())
    assert float((pooled - picked).abs().max()) < 1e-5                    # every chosen cell is a maximum of its window
    if pool > 1:                                          # windows whose cells are all equal (flat region): the FIRST cell wins
        win = conv.detach().unfold(2, pool, pool).unfold(3, pool, pool).flatten(4)      # [N, Cout, Hp, Hp, pool*pool]
        flatwin = (win.max(4)[0] == win.min(4)[0])
        assert int(flatwin.sum()) > 0 and bool((amc[flatwin] == 0).all())
    gsel = go.double()
    if relu:
        gsel = gsel * (y.permute(0, 3, 1, 2) > 0)
    gconv = torch.zeros_like(conv).flatten(2).scatter_(2, flat.flatten(2), gsel.flatten(2)).view_as(conv)
    conv.backward(gconv)
    dw = torch.zeros_like(wk)
    db = torch.zeros(Cout, device='cuda')
    dx = torch.zeros_like(xn) if Cin > 1 else None
    gon = go.permute(0, 2, 3, 1).contiguous()
    lib.dpp_convpool_bwd(P(xn), P(wk), P(y), P(am), P(gon), P(dw), P(db), P(dx), N, H, H, Cin, Cout, k, pad, pool, relu, None)
    torch.cuda.synchronize()
    dwr = _kc(wd.grad.float())
    dbr = gsel.sum((0, 2, 3)).float()
    err = (dw - dwr).abs().max() / dwr.abs().max()
    print(cfg, "dW rel err", float(err), "db err", float((db - dbr).abs().max() / dbr.abs().max()))
    assert err < 2e-5
    assert torch.allclose(db, dbr, rtol=1e-4, atol=1e-4 * float(dbr.abs().max()))
    if Cin > 1:
        dxr = xd.grad.float().permute(0, 2, 3, 1)
        assert (dx - dxr).abs().max() / dxr.abs().max() < 2e-5


@pytest.mark.parametrize("precision,tol", [(0, 1e-5), (1, 2e-5), (2, 4e-3)])
@pytest.mark.parametrize("dims", [(128, 16384, 1024, 1), (128, 1024, 1024, 1), (4, 1024, 32, 0), (128, 968, 1024, 1),
                                  (128, 1024, 30, 0),
                                  # the streaming kernel's other paths: several sample tiles, a ragged last one, a
                                  # half-width MMA (64 samples per GPU in the strong-scaling split), ragged unit tiles
                                  (512, 1024, 1024, 1), (200, 1024, 256, 1), (64, 16384, 1024, 1), (16, 328, 200, 0)])
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
    # accumulate semantics of dpp_fc_bwd: a second call doubles dW; DPP_FC_DW_ASSIGN overwrites whatever dw held
    lib.dpp_fc_bwd(P(x.detach()), P(w.detach()), P(o.detach().contiguous()), P(go), P(dw), P(db), None, P(scratch), B, n_in, n_out,
                   relu, None, 1.0, precision, None)
    torch.cuda.synchronize()
    assert float((dw - 2 * w.grad).abs().max() / w.grad.abs().max()) < 2 * tol
    dw.fill_(float('nan'))
    lib.dpp_fc_bwd_ex(P(x.detach()), P(w.detach()), P(o.detach().contiguous()), P(go), P(dw), P(db), None, P(scratch), B, n_in,
                      n_out, relu, None, 1.0, precision, 1, None)
    torch.cuda.synchronize()
    assert float((dw - w.grad).abs().max() / w.grad.abs().max()) < tol


def test_fc_fwd_epilogue_with_mask_and_scale():
    """the k-split reduction of the streaming GEMM applies bias, ReLU, dropout mask and output scale itself"""
    B, n_in, n_out = 128, 2048, 1024
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn(B, n_in, device='cuda', generator=g)
    w = torch.randn(n_in, n_out, device='cuda', generator=g) * (1.0 / n_in) ** 0.5
    b = torch.randn(n_out, device='cuda', generator=g) * 0.1
    mask = (torch.rand(B, n_out, device='cuda', generator=g) < 0.7).float()
    for precision in (0, 1):
        y = torch.full((B, n_out), float('nan'), device='cuda')
        lib.dpp_fc_fwd(P(x), P(w), P(b), P(y), B, n_in, n_out, 1, P(mask), 0.7, precision, None)
        torch.cuda.synchronize()
        ref = torch.relu(x.double() @ w.double() + b.double()) * mask.double() * 0.7
        assert float((y.double() - ref).abs().max() / ref.abs().max()) < 2e-5


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
