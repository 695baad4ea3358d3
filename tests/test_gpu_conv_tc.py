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


# ---------------------------------------------------------------------------------------------------
# The benchmarked shapes (batch 128 and beyond) against an INDEPENDENT reference: torch float64 on the
# GPU (no TF32 anywhere).  tiles >> 148, so every persistent CTA walks several tiles: accumulator
# double-buffer phase flips, both epilogue warpgroups, A-stage ring across tiles, resident weight images
# (kchunks <= ring) and streamed ones (kchunks > ring), odd and even tiles per CTA, split / replica
# choices of the backward-weights kernel.  Reference semantics: net/convlayer.py:230-235 (conv2d 'half',
# subsample), net/batchnormlayer.py:154-159 (batch statistics), nonlinearitylayer.py:119 (ReLU).
# ---------------------------------------------------------------------------------------------------
BIG_SHAPES = [  # N, H, Cin, Cout, k, stride           m-tiles x n-tiles (tiles per CTA), weight image
    (128, 32, 16, 16, 3, 1),      # 1024 (6-7), resident 5 chunks, BN=16
    (128, 32, 16, 64, 1, 1),      # 1024 (6-7), resident, BN=64, residual epilogue
    (128, 32, 64, 16, 1, 1),      # 1024, 2 chunks
    (128, 64, 32, 16, 1, 2),      # 1024, stride 2 gather / strided scatter in dgrad
    (128, 64, 32, 64, 1, 2),      # projection shortcut
    (128, 16, 32, 32, 3, 1),      # 256 (1-2), 9 chunks > ring of 6: streamed weights across tiles
    (128, 16, 32, 128, 1, 1),     # 256, BN=128
    (128, 16, 128, 32, 1, 1),     # 256, 4 chunks
    (128, 8, 64, 64, 3, 1),       # 64 tiles, 18 chunks streamed
    (128, 8, 256, 64, 1, 1),      # 64 tiles, 8 chunks streamed (ring 4)
    (128, 8, 64, 256, 1, 1),      # 64 x 2 n-tiles
    (512, 8, 64, 64, 3, 1),       # 256 tiles (1-2), 18 chunks streamed, BN=64
    (512, 8, 64, 256, 1, 1),      # 512 tiles (3-4), 2 n-tiles
    (37, 32, 16, 16, 3, 1),       # 296 tiles = exactly 2 per CTA
    (19, 32, 16, 64, 1, 1),       # 152 tiles: 4 CTAs with 2 tiles, 144 with 1
]


def _ref_ops(N, H, Cin, Cout, k, stride, x, w, bias, gamma, beta, res, dy, dx_prev, dz_kernel=None, eps=1e-4):
    """float64 torch restatement of the three fused operations (NHWC in / out like the kernels)."""
    import torch.nn.functional as F
    xd = x.double().permute(0, 3, 1, 2)
    m = xd.mean((0, 2, 3), keepdim=True)
    v = (xd * xd).mean((0, 2, 3), keepdim=True) - m * m
    istd = 1.0 / torch.sqrt(v.clamp_min(0) + eps)
    sc = gamma.double().view(1, -1, 1, 1) * istd
    # the kernels evaluate the prologue in fp32 (x * scale + shift); do the ReLU decision on the same fp32 values
    scf, shf = sc.float(), (beta.double().view(1, -1, 1, 1) - m * sc).float()
    pre32 = torch.addcmul(shf, x.permute(0, 3, 1, 2), scf)
    a = torch.relu(pre32).double().requires_grad_(True)
    W = w.double().reshape(k, k, Cin, Cout).permute(3, 2, 0, 1).contiguous().requires_grad_(True)
    y0 = F.conv2d(a, W, None, stride=stride, padding=k // 2)
    out = {}
    y = y0 + bias.double().view(1, -1, 1, 1)
    if res is not None:
        y = y + res.double().permute(0, 3, 1, 2)
    out['y'] = y.permute(0, 2, 3, 1).detach()
    out['stats'] = torch.cat([y.sum((0, 2, 3)), (y * y).sum((0, 2, 3))]).detach()
    if dy is not None:
        ga, gW = torch.autograd.grad(y0, (a, W), dy.double().permute(0, 3, 1, 2))
        da = ga
        if dx_prev is not None:
            da = da + dx_prev.double().permute(0, 3, 1, 2)
        on = pre32 > 0
        if dz_kernel is not None:
            # a BN output within fp32 roundoff of 0 may legitimately land on either side (the kernel folds
            # gamma * inv_std in fp32): for those few elements take the kernel's own on / off decision
            near = pre32.abs() < 1e-5
            out['near'] = int(near.sum())
            on = torch.where(near, dz_kernel.permute(0, 3, 1, 2) != 0, on)
        dz = torch.where(on, da, torch.zeros_like(da))
        xhat = (xd - m) * istd
        out['dz'] = dz.permute(0, 2, 3, 1)
        out['dzstats'] = torch.cat([dz.sum((0, 2, 3)), (dz * xhat).sum((0, 2, 3))])
        out['dw'] = gW.permute(2, 3, 1, 0).reshape(k * k * Cin, Cout)
        out['db'] = dy.double().sum((0, 1, 2))
    return out


def _relmax(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", BIG_SHAPES)
@pytest.mark.parametrize("precision,tol", [(1, 2e-5), (0, 2e-5)])
def test_bench_shapes_vs_torch_fp64(shape, precision, tol):
    """forward (+ residual, statistics), backward-data (accumulate where the layer is stride 1, ReLU mask,
    BN-backward statistics) and backward-weights of one layer at the benchmarked batch size, each compared
    with float64 torch.  precision 0 anchors the fp32 SIMT kernels (the other tests' comparison partner) to
    the same independent reference."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    N, H, Cin, Cout, k, stride = shape
    d, x, w, bias, bn, Ho, keep = _setup(N, H, Cin, Cout, k, stride, precision)
    gamma, beta = keep[-3], keep[-2]
    g = torch.Generator(device='cuda').manual_seed(21)
    res = torch.randn(N, Ho, Ho, Cout, device='cuda', generator=g)
    dy = torch.randn(N, Ho, Ho, Cout, device='cuda', generator=g)
    acc = stride == 1
    dx = torch.randn(N, H, H, Cin, device='cuda', generator=g) if acc else torch.zeros(N, H, H, Cin, device='cuda')
    # forward
    y = torch.full((N, Ho, Ho, Cout), float('nan'), device='cuda')
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device='cuda')
    lib.dpp_conv2d_fwd(C.byref(d), P(x), C.byref(bn), P(w), P(bias), P(res), P(y), P(stats), None)
    torch.cuda.synchronize()
    # backward-data
    dzs = torch.zeros(2 * Cin, dtype=torch.float64, device='cuda')
    dxo = dx.clone()
    lib.dpp_conv2d_dgrad(C.byref(d), P(dy), P(w), P(dxo), 1 if acc else 0, C.byref(bn), P(x), P(dzs), None)
    torch.cuda.synchronize()
    ref = _ref_ops(N, H, Cin, Cout, k, stride, x, w, bias, gamma, beta, res, dy, dx if acc else None, dz_kernel=dxo)
    e_y = _relmax(y, ref['y'])
    e_s = _relmax(stats, ref['stats'])
    e_dz = _relmax(dxo, ref['dz'])
    e_dzs = float((dzs - ref['dzstats']).abs().max() / ref['dzstats'].abs().max())
    # backward-weights
    dw = torch.zeros(k * k * Cin, Cout, device='cuda')
    db = torch.zeros(Cout, device='cuda')
    lib.dpp_conv2d_wgrad(C.byref(d), P(x), C.byref(bn), P(dy), P(dw), P(db), None)
    torch.cuda.synchronize()
    e_dw = _relmax(dw, ref['dw'])
    e_db = _relmax(db, ref['db'])
    print(shape, "precision", precision, "fwd %.2e stats %.2e | dgrad %.2e stats %.2e | wgrad %.2e db %.2e"
          % (e_y, e_s, e_dz, e_dzs, e_dw, e_db))
    assert not torch.isnan(y).any()
    assert e_y < tol and e_dz < tol
    assert e_s < 1e-5 and e_dzs < 1e-4        # fp32 tile partials summed in fp64
    assert e_dw < 5e-5 and e_db < 1e-5        # reduction over up to 5e5 pixels in fp32 partials


GROUP_SHAPES = [  # the grouped launches: fp32 3x3 kernel (C = 16 / 32) next to the tcgen05 groups
    (128, 32, 16, 16, 3, 1),      # fp32 3x3, 16 pixel lanes, 4 row blocks per image
    (128, 16, 32, 32, 3, 1),      # fp32 3x3, 4 pixel lanes
    (3, 20, 16, 16, 3, 1),        # fp32 3x3: ragged last row block (20 = 8 + 8 + 4), width not a power of two
    (5, 8, 32, 32, 3, 1),         # fp32 3x3: one row block per image, fewer items than SMs
    (128, 32, 64, 16, 1, 1),      # fp32 1x1 (k_wgrad1<64, 16>): 512 tiles of 256 pixels
    (128, 32, 16, 64, 1, 1),      # fp32 1x1 (k_wgrad1<16, 64>)
    (3, 20, 64, 16, 1, 1),        # fp32 1x1: 1200 pixels = 4 full tiles + a ragged one
    (128, 16, 128, 32, 1, 1),     # tcgen05 group, BN=32
    (128, 8, 64, 64, 3, 1),       # tcgen05 group, BN=64 (3x3 with 64 channels stays on the tensor cores)
]


@pytest.mark.parametrize("simt3", ["3", "1", "0"])
def test_wgrad_group_vs_torch_fp64(simt3, monkeypatch):
    """dpp_wgrad_group_*: all layers of GROUP_SHAPES in one handle (two runs: dw / db accumulate), every dw / db against
    float64 torch.  DPP_WGRAD_SIMT3: bit 0 = fp32 kernel for the 3x3 16 / 32-channel layers, bit 1 = fp32 kernel for the
    1x1 64 -> 16 / 16 -> 64 layers (default 3); 0 keeps everything on the tcgen05 kernels."""
    monkeypatch.setenv("DPP_WGRAD_SIMT3", simt3)
    from dpp_b200.lib import WgradLayer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    keepall, layers, refs = [], [], []
    for i, (N, H, Cin, Cout, k, stride) in enumerate(GROUP_SHAPES):
        d, x, w, bias, bn, Ho, keep = _setup(N, H, Cin, Cout, k, stride, 1, seed=40 + i)
        gamma, beta = keep[-3], keep[-2]
        g = torch.Generator(device='cuda').manual_seed(60 + i)
        dy = torch.randn(N, Ho, Ho, Cout, device='cuda', generator=g)
        dw = torch.zeros(k * k * Cin, Cout, device='cuda')
        db = torch.zeros(Cout, device='cuda')
        wl = WgradLayer()
        wl.d, wl.x, wl.in_bn, wl.has_in_bn = d, x.data_ptr(), bn, 1
        wl.dy, wl.dw, wl.db = dy.data_ptr(), dw.data_ptr(), db.data_ptr()
        layers.append(wl)
        ref = _ref_ops(N, H, Cin, Cout, k, stride, x, w, bias, gamma, beta, None, dy, None)
        refs.append((dw, db, ref['dw'], ref['db']))
        keepall += [x, w, bias, dy, keep]
    arr = (WgradLayer * len(layers))(*layers)
    h = C.c_void_p()
    create = lib.raw('dpp_wgrad_group_create')
    assert create(arr, len(layers), C.byref(h)) == 0
    n_launch = int(lib.dpp_wgrad_group_launches(h))
    lib.dpp_wgrad_group_run(h, None)
    lib.dpp_wgrad_group_run(h, None)
    torch.cuda.synchronize()
    lib.dpp_wgrad_group_destroy(h)
    print("launches per run:", n_launch)
    # fp32 launches: 3x3 C = 16, C = 32 (bit 0), 1x1 64 -> 16, 16 -> 64 (bit 1); tcgen05 widths left: 32, 64 (+ 16 without bit 1)
    assert n_launch == {"3": 6, "1": 5, "0": 3}[simt3]
    for (shape, (dw, db, rdw, rdb)) in zip(GROUP_SHAPES, refs):
        e_dw, e_db = _relmax(dw, 2 * rdw), _relmax(db, 2 * rdb)
        print(shape, "simt3", simt3, "dw %.2e db %.2e" % (e_dw, e_db))
        assert e_dw < 5e-5 and e_db < 1e-5
