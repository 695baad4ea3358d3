"""Kernel-level GPU tests through the C ABI: the tcgen05 implicit-GEMM convolution (3xTF32 and
TF32) against the fp32 SIMT kernels on the same random inputs, for every conv shape class of the
ResNet (1x1, 1x1 stride 2, 3x3, n-tiles of 16..256 channels), forward (BN+ReLU prologue, bias,
residual, BN statistics) and backward-data (accumulate, ReLU mask, BN-backward statistics)."""
import ctypes as C
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from dpp_b200.lib import lib, ConvDesc, BnRef, PackItem  # noqa: E402

SHAPES = [  # N, H, Cin, Cout, k, stride
    (2, 32, 16, 16, 3, 1),
    (2, 32, 64, 16, 1, 1),
    (2, 64, 32, 16, 1, 2),
    (2, 64, 32, 64, 1, 2),
    (3, 16, 32, 128, 1, 1),
    (4, 8, 64, 256, 1, 1),
    (4, 8, 256, 64, 1, 1),
    (4, 8, 64, 64, 3, 1),
    (1, 8, 64, 64, 3, 1),       # M = 64 < one tile
]


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _setup(N, H, Cin, Cout, k, stride, precision, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    Ho = (H + 2 * (k // 2) - k) // stride + 1
    x = torch.randn(N, H, H, Cin, device='cuda', generator=g) * 2 + 0.5
    w = torch.randn(k * k * Cin, Cout, device='cuda', generator=g) * (2.0 / (k * k * Cin)) ** 0.5
    bias = torch.randn(Cout, device='cuda', generator=g) * 0.1
    gamma = torch.rand(Cin, device='cuda', generator=g) + 0.5
    beta = torch.randn(Cin, device='cuda', generator=g) * 0.2
    sums = torch.cat([x.double().sum((0, 1, 2)), (x.double() ** 2).sum((0, 1, 2))]).contiguous()
    d = ConvDesc()
    d.N, d.H, d.W, d.Cin, d.Cout, d.k, d.stride, d.pad, d.Ho, d.Wo = N, H, H, Cin, Cout, k, stride, k // 2, Ho, Ho
    d.precision = precision
    keep = []
    if precision:
        f, gsz = C.c_int64(), C.c_int64()
        lib.dpp_conv_pack_size(Cin, Cout, k, precision, C.byref(f), C.byref(gsz))
        pf = torch.zeros(f.value, device='cuda')
        pg = torch.zeros(gsz.value, device='cuda')
        it = (PackItem * 1)()
        it[0].w, it[0].img_fwd, it[0].img_dgrad = w.data_ptr(), pf.data_ptr(), pg.data_ptr()
        it[0].Cin, it[0].Cout, it[0].k = Cin, Cout, k
        it[0].bn_fwd, it[0].bn_dgrad, it[0].passes = min(Cout, 128), min(Cin, 128), 2 if precision == 1 else 1
        items = torch.frombuffer(bytearray(bytes(it)), dtype=torch.uint8).cuda()
        lib.dpp_conv_pack_all(P(items), 1, None)
        d.wpack_fwd, d.wpack_dgrad = pf.data_ptr(), pg.data_ptr()
        keep += [pf, pg, items]
    bn = BnRef()
    bn.sums, bn.mean, bn.inv_std = sums.data_ptr(), None, None
    bn.gamma, bn.beta, bn.count, bn.eps, bn.relu = gamma.data_ptr(), beta.data_ptr(), float(N * H * H), 1e-4, 1
    keep += [gamma, beta, sums]
    return d, x, w, bias, bn, Ho, keep


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("precision,tol", [(1, 2e-5), (2, 4e-3)])
def test_fwd_tc_vs_simt(shape, precision, tol):
    N, H, Cin, Cout, k, stride = shape
    outs = []
    for prec in (0, precision):
        d, x, w, bias, bn, Ho, keep = _setup(N, H, Cin, Cout, k, stride, prec)
        res = torch.randn(N, Ho, Ho, Cout, device='cuda', generator=torch.Generator(device='cuda').manual_seed(5))
        y = torch.zeros(N, Ho, Ho, Cout, device='cuda')
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device='cuda')
        lib.dpp_conv2d_fwd(C.byref(d), P(x), C.byref(bn), P(w), P(bias), P(res), P(y), P(stats), None)
        torch.cuda.synchronize()
        outs.append((y.cpu().numpy(), stats.cpu().numpy()))
    (y0, s0), (y1, s1) = outs
    err = np.abs(y1 - y0).max() / np.abs(y0).max()
    print(shape, precision, "fwd rel err", err)
    assert err < tol
    assert np.allclose(s1, s0, rtol=max(tol * 10, 1e-4), atol=np.abs(s0).max() * tol)
    # statistics are those of the written tensor
    assert np.allclose(s1[:Cout], y1.astype(np.float64).sum((0, 1, 2)), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("precision,tol", [(1, 2e-5), (2, 4e-3)])
def test_dgrad_tc_vs_simt(shape, precision, tol):
    N, H, Cin, Cout, k, stride = shape
    outs = []
    for prec in (0, precision):
        d, x, w, bias, bn, Ho, keep = _setup(N, H, Cin, Cout, k, stride, prec)
        g = torch.Generator(device='cuda').manual_seed(9)
        dy = torch.randn(N, Ho, Ho, Cout, device='cuda', generator=g)
        dx = torch.randn(N, H, H, Cin, device='cuda', generator=g)        # pre-existing contents (accumulate)
        if stride != 1:
            dx.zero_()
        dzs = torch.zeros(2 * Cin, dtype=torch.float64, device='cuda')
        lib.dpp_conv2d_dgrad(C.byref(d), P(dy), P(w), P(dx), 1, C.byref(bn), P(x), P(dzs), None)
        torch.cuda.synchronize()
        outs.append((dx.cpu().numpy(), dzs.cpu().numpy()))
    (a0, s0), (a1, s1) = outs
    # ReLU-mask decisions are identical (same prologue arithmetic); compare values
    err = np.abs(a1 - a0).max() / np.abs(a0).max()
    print(shape, precision, "dgrad rel err", err)
    assert err < tol
    assert np.allclose(s1, s0, rtol=max(tol * 20, 1e-4), atol=np.abs(s0).max() * tol * 10)


@pytest.mark.parametrize("mn", ["1", "0"])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("precision,tol", [(1, 2e-5), (2, 4e-3)])
def test_wgrad_tc_vs_simt(shape, precision, tol, mn, monkeypatch):
    """mn=1: MN-major operand tiles (wgrad_tc_mn.cu); mn=0: K-major transposing kernel (conv_tc.cu)"""
    monkeypatch.setenv("DPP_WGRAD_MN", mn)
    N, H, Cin, Cout, k, stride = shape
    outs = []
    for prec in (0, precision):
        d, x, w, bias, bn, Ho, keep = _setup(N, H, Cin, Cout, k, stride, prec)
        g = torch.Generator(device='cuda').manual_seed(11)
        dy = torch.randn(N, Ho, Ho, Cout, device='cuda', generator=g)
        dw = torch.zeros(k * k * Cin, Cout, device='cuda')
        db = torch.zeros(Cout, device='cuda')
        lib.dpp_conv2d_wgrad(C.byref(d), P(x), C.byref(bn), P(dy), P(dw), P(db), None)
        torch.cuda.synchronize()
        outs.append((dw.cpu().numpy(), db.cpu().numpy()))
    (w0, b0), (w1, b1) = outs
    err = np.abs(w1 - w0).max() / np.abs(w0).max()
    print(shape, precision, "wgrad rel err", err)
    assert err < tol
    assert np.allclose(b1, b0, rtol=1e-4, atol=1e-4 * np.abs(b0).max())
