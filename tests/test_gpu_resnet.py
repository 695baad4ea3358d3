"""GPU parity: ResNet forward / cost / gradients / ADAM step through the reference-surface
classes + libdpp_b200.so, against the CPU oracle (oracle/nets.py) on the same seeded inputs.
Tolerance: regressed outputs within 1e-4 relative (north_star), gradients within 5e-3 of each tensor's
max and 2e-3 in relative L2 (fp32 summation-order noise through 61 batch-norms at batch 4)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _build(net_type, B, J, D, seed=23455, precision=0):
    from net.resnet import ResNet, ResNetParams
    from oracle import nets as O
    net = ResNet(np.random.RandomState(seed), cfgParams=ResNetParams(type=net_type, batchSize=B, numJoints=J, nDims=D))
    onet = O.build_resnet(np.random.RandomState(seed), type=net_type, batchSize=B, numJoints=J, nDims=D)
    from dpp_b200.engine import Engine
    eng = Engine(net, precision=precision)
    net._eng = eng
    return net, onet, eng


def _data(B, D, seed=1):
    rng = np.random.RandomState(seed)
    x = rng.uniform(-1, 1, (B, 1, 128, 128)).astype(np.float32)
    x[:, :, :20] = 1.0
    y = rng.randn(B, D).astype(np.float32)
    return x, y


def test_forward_train_mode_matches_oracle():
    from oracle import nets as O
    B, D = 4, 30
    net, onet, eng = _build(0, B, 1, D)
    x, y = _data(B, D)
    eng.set_input_nchw(x)
    out = eng.forward_device(deterministic=False).cpu().numpy()
    with torch.no_grad():
        oout, _ = onet.forward(torch.from_numpy(x), deterministic=False)
    r = _rel(out, oout.numpy())
    print("forward(train) rel err", r)
    assert r < 1e-4


def test_forward_deterministic_matches_oracle():
    B, D = 4, 30
    net, onet, eng = _build(0, B, 1, D)
    x, y = _data(B, D)
    out = net.computeOutput(x[:3])          # exercises padding to the batch size
    net.setDeterministic()
    with torch.no_grad():
        oout, _ = onet.forward(torch.from_numpy(np.concatenate([x[:3], x[2:3]])), deterministic=True)
    r = _rel(out, oout.numpy()[:3])
    print("forward(det) rel err", r)
    assert out.shape == (3, D)
    assert r < 1e-4


def _engine_relu_masks(net, onet, eng, w_before):
    """The on/off decision the engine took at every BN->ReLU of the step it just ran (from its own raw
    buffers, batch statistics and the pre-step gamma/beta, through the library's dpp_bn_apply), keyed by
    the oracle's ReLU layer number."""
    import ctypes as C
    from dpp_b200.lib import lib
    from dpp_b200.engine import _ptr
    w_after = eng.W.clone()
    eng.W.copy_(w_before)
    masks, done = {}, set()
    for op in eng.ops:
        if op['kind'] == 'conv' and op['in_bn'] is not None:
            bn, raw = op['in_bn'], op['src']
        elif op['kind'] == 'bn_apply':
            bn, raw = op['bn'], op['src']
        else:
            continue
        if id(bn) in done:
            continue
        done.add(id(bn))
        i = net.layers.index(bn)
        assert onet.layers[i + 1].kind == 'relu'
        tmp = torch.empty_like(raw.buf)
        ref = eng._bnref(bn, raw, True, relu=1)
        lib.dpp_bn_apply(_ptr(raw.buf), C.byref(ref), _ptr(tmp), raw.pixels, int(raw.shape[1]), eng._stream())
        n, c, h, w = raw.shape
        masks[i + 1] = (tmp.reshape(n, h, w, c) > 0).permute(0, 3, 1, 2).contiguous().cpu()
    # the ReLU inside a HiddenLayer (no dropout behind it): on where the stored output is positive.  One flip among
    # the 128 x 1024 units of the first FC layer moves every gradient below it by ~1 / sqrt(#units) = 0.3 %.
    for op in eng.ops:
        if op['kind'] == 'fc' and op['dropout'] is None and op['layer'].cfgParams.activation_str == 'ReLU':
            masks[op['layer'].layerNum] = (op['dst'].buf > 0).cpu()
    eng.W.copy_(w_after)
    torch.cuda.synchronize()
    return masks


@pytest.mark.parametrize("use_graph,precision,B", [(False, 0, 4), (True, 0, 4), (True, 1, 4), (True, 1, 128)])
def test_train_step_matches_oracle(use_graph, precision, B):
    """precision 0 = fp32 SIMT kernels, 1 = 3xTF32 tcgen05 kernels (conv fwd/dgrad/wgrad).  B = 128 is the
    benchmarked configuration (BASELINE config 2): every persistent conv CTA walks several tiles there, the
    backward-weights kernels use their batch-128 split / replica choice and the FC GEMMs their split-K.

    Gradient parity is taken with the engine's ReLU decisions imposed on the (fp64) oracle: about 1e-6 of
    the ~5e6 BN->ReLU inputs lie within fp32 roundoff of 0 and may legitimately land on either side in two
    correct implementations; ONE such flip in the last stage changes every gradient below it by
    ~1/sqrt(#units) = 0.4 % in relative L2, which would drown the 1e-4 bar in coin tosses.  With the
    decisions pinned, the two backward passes compute the same function and must agree to fp32 accuracy.
    The un-pinned comparison is still made, with the loose bound such flips allow."""
    from oracle import nets as O
    D = 30
    net, onet, eng = _build(0, B, 1, D, precision=precision)
    x, y = _data(B, D)
    adam = O.Adam(onet.params)
    lr = 1e-3
    lays = [l for l in onet.layers for _ in l.params]
    for step in range(2 if B <= 16 else 1):
        eng.set_input_nchw(x)
        eng._alloc_training()
        eng.y_in.copy_(torch.from_numpy(y))
        w_before = eng.W.clone()
        cost = float(eng.train_step(lr, use_graph=use_graph).cpu()[0])
        grads = eng.gradients()
        masks = _engine_relu_masks(net, onet, eng, w_before)
        nflip = 0
        with torch.no_grad():
            col = {}
            onet.forward(torch.from_numpy(x), deterministic=False, collect=col)
            for ln, m in masks.items():
                nflip += int(((col[ln] > 0) != m).sum())       # (a unit that is exactly 0 counts as off on both sides)
        # un-pinned gradients (no state change: plain autograd on the oracle)
        out_free, _ = onet.forward(torch.from_numpy(x), deterministic=False)
        cost_free = O.cost_fn(onet, out_free, torch.from_numpy(y), B, 1, D, 0.0)
        g_free = torch.autograd.grad(cost_free, onet.params, allow_unused=True)
        ocost, oout, ograds = O.train_step(onet, adam, torch.from_numpy(x), torch.from_numpy(y), lr, 1, D,
                                           relu_masks=masks)
        print("step", step, "cost", cost, ocost, "relu decisions differing from the fp64 oracle:", nflip)
        assert abs(cost - ocost) < 1e-4 * abs(ocost)
        worst, worst2, worst_free = 0.0, 0.0, 0.0
        table = []
        for p, og, gf, l in zip(net.params, ograds, g_free, lays):
            g = grads[id(p)]
            if l.kind in ('conv', 'convpool') and g.ndim == 1:
                continue            # conv biases feed only BNs: true gradient is exactly 0 (roundoff only)
            og = og.numpy()
            e = float(np.abs(g - og).max() / (np.abs(og).max() + 1e-12))
            e2 = float(np.linalg.norm((g - og).ravel()) / (np.linalg.norm(og.ravel()) + 1e-30))
            worst, worst2 = max(worst, e), max(worst2, e2)
            table.append((p.name, l.kind, e, e2))
            if gf is not None:
                gf = gf.numpy()
                ef = float(np.linalg.norm((g - gf).ravel()) / (np.linalg.norm(gf.ravel()) + 1e-30))
                worst_free = max(worst_free, ef)
                assert ef < 0.25, (p.name, ef)     # loose: a handful of roundoff-level ReLU flips at batch 4 (see docstring)
        print("pinned-ReLU grad err: max-rel %.3g, rel-L2 %.3g; un-pinned rel-L2 %.3g" % (worst, worst2, worst_free))
        bad = [t for t in table if not (t[2] < 5e-4 and t[3] < 5e-4)]   # fp32 (engine) vs fp64 (oracle) through 189 layers
        if bad:
            print("gradient tensors, forward order (name kind max-rel rel-L2):")
            for t in table:
                print("   %-12s %-9s %.3g %.3g%s" % (t + (("  <-- " if t in bad else ""),)))
        assert not bad, bad[:4]
        # parameters after this ADAM step.  The first ADAM step is lr*sign(g): an entry whose gradient is
        # within roundoff of 0 may move the other way (|diff| = 2 lr); such entries must be rare, the rest
        # must agree closely.  Then hand the engine's weights to the oracle so that the next step again
        # compares the same function (skip the zero-gradient conv biases).
        nbad = ntot = 0
        for p, op_, l in zip(net.params, onet.params, lays):
            a = p.get_value()
            if not (l.kind in ('conv', 'convpool') and a.ndim == 1):
                b = op_.detach().numpy()
                diff = np.abs(a - b)
                assert diff.max() < 2.5e-3 * max(1.0, np.abs(b).max()), p.name   # |step| <= lr each
                nbad += int((diff > 2e-4).sum())
                ntot += diff.size
            with torch.no_grad():
                op_.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(op_.dtype))
        print("parameters off by > 2e-4 (0.1 lr) after the step: %d of %d" % (nbad, ntot))
        assert nbad < 2e-2 * ntot       # ADAM itself is checked exactly in test_adam_matches_oracle_on_identical_gradients
    # BN running statistics (EMA of mean and inv_std)
    for l, ol in zip(net.layers, onet.layers):
        if ol.kind == 'bn':
            assert _rel(l.params_nontrained[0].get_value(), ol.nontrained[0].numpy()) < 1e-3 or \
                np.abs(ol.nontrained[0].numpy()).max() < 1e-3
            assert _rel(l.params_nontrained[1].get_value(), ol.nontrained[1].numpy()) < 1e-4


def test_forward_tc_3xtf32_matches_oracle():
    B, D = 4, 30
    net, onet, eng = _build(0, B, 1, D, precision=1)
    x, y = _data(B, D)
    eng.set_input_nchw(x)
    out = eng.forward_device(deterministic=False).cpu().numpy()
    with torch.no_grad():
        oout, _ = onet.forward(torch.from_numpy(x), deterministic=False)
    r = _rel(out, oout.numpy())
    print("forward(train, 3xTF32 tcgen05) rel err", r)
    assert r < 1e-4


def test_type1_with_pca_tail_forward():
    B, J = 4, 14
    net, onet, eng = _build(1, B, J, 3)
    x, _ = _data(B, 3 * J)
    eng.set_input_nchw(x)
    out = eng.forward_device(deterministic=False).cpu().numpy()
    with torch.no_grad():
        oout, _ = onet.forward(torch.from_numpy(x), deterministic=False)
    assert out.shape == (B, 3 * J)
    assert _rel(out, oout.numpy()) < 1e-4


@pytest.mark.parametrize("precision", [0, 1])
def test_mean_joint_error_mm_matches_oracle(precision):
    """north_star: mean 3D joint error within 0.1 mm of the reference on the same synthetic batch.
    Synthetic NYU crops (data/synthetic.py), type-1 ResNet with the PCA prior layer (30 -> 14*3); joints in mm as
    handpose_evaluation computes them: joints_mm = out * cube_z / 2; error = mean_j ||pred_j - gt_j||."""
    from data import synthetic
    B, J = 8, 14
    ds = synthetic.generate('NYU', B, seed=23455)
    comp, mean = synthetic.random_orthonormal_pca(30, 3 * J, seed=1)
    net, onet, eng = _build(1, B, J, 3, precision=precision)
    # install the PCA prior in the last layer of both nets (main_nyu_posereg_embedding.py:148-158)
    net.layers[-1].W.set_value(comp.astype(np.float32))
    net.layers[-1].b.set_value(mean.astype(np.float32))
    with torch.no_grad():
        onet.layers[-1].params[0].copy_(torch.from_numpy(comp.astype(np.float32)))
        onet.layers[-1].params[1].copy_(torch.from_numpy(mean.astype(np.float32)))
    x = ds['x'].astype(np.float32)
    # batch statistics (the running ones of an untrained net are 0 / 1 and let the activations explode through 61
    # BatchNorms, which would turn the 1e-4 relative bar into metres)
    eng.set_input_nchw(x)
    out = eng.forward_device(deterministic=False).cpu().numpy()
    with torch.no_grad():
        oout, _ = onet.forward(torch.from_numpy(x), deterministic=False)
    oout = oout.numpy()
    half = (ds['cube'][:, 2] / 2.0).reshape(B, 1, 1)
    gt = ds['gt3Dcrop'].reshape(B, J, 3)
    pred = out.reshape(B, J, 3) * half
    opred = oout.reshape(B, J, 3) * half
    err = np.linalg.norm(pred - gt, axis=2).mean()
    oerr = np.linalg.norm(opred - gt, axis=2).mean()
    dmax = np.abs(pred - opred).max()
    print("mean joint error: engine %.4f mm, oracle %.4f mm, max joint coordinate difference %.2e mm" % (err, oerr, dmax))
    assert abs(err - oerr) < 0.1
    assert dmax < 0.1
    assert _rel(out, oout) < 1e-4


def test_block0_backward_matches_torch_gpu_autograd_and_oracle():
    """Isolates the stem: (a) engine dW0 vs torch autograd on the engine's own dy_stem, (b) the
    engine's dy_stem vs the oracle's gradient at the stem output."""
    import torch.nn.functional as F
    from oracle import nets as O
    B, D = 4, 30
    net, onet, eng = _build(0, B, 1, D)
    x, y = _data(B, D)
    eng.set_input_nchw(x)
    eng._alloc_training()
    eng.y_in.copy_(torch.from_numpy(y))
    eng.train_step(0.0, use_graph=False)
    stem_op = [o for o in eng.ops if o['kind'] == 'convpool'][0]
    dy = stem_op['dst'].grad.clone()                       # NHWC
    g_eng = eng.gradients()[id(net.layers[0].W)]
    # (a)
    w = torch.from_numpy(net.layers[0].W.get_value()).cuda().requires_grad_(True)
    xt = torch.from_numpy(x).cuda()
    o = F.max_pool2d(F.conv2d(xt, w.flip(2, 3), padding=2), 2, 2)
    o.backward(dy.permute(0, 3, 1, 2).contiguous())
    ea = float((torch.from_numpy(g_eng).cuda() - w.grad).abs().max() / w.grad.abs().max())
    # (b)
    for p in onet.params:
        p.grad = None
    col = {}
    out, _ = onet.forward(torch.from_numpy(x), deterministic=False, collect=col)
    col[0].retain_grad()
    for k_ in col:
        if col[k_].requires_grad:
            col[k_].retain_grad()
    O.cost_fn(onet, out, torch.from_numpy(y), B, 1, D).backward()
    dyo = col[0].grad.permute(0, 2, 3, 1).numpy()
    eb = float(np.abs(dy.cpu().numpy() - dyo).max() / np.abs(dyo).max())
    eb2 = float(np.linalg.norm(dy.cpu().numpy() - dyo) / np.linalg.norm(dyo))
    gw = onet.layers[0].params[0].grad.numpy()
    ec = float(np.abs(g_eng - gw).max() / np.abs(gw).max())
    print("stem dW vs torch-GPU autograd on same dy:", ea, "| dy_stem vs oracle: max", eb, "l2", eb2, "| dW vs oracle", ec)
    d = np.abs(dy.cpu().numpy() - dyo) / np.abs(dyo).max()
    print("  err even-even", d[:, ::2, ::2].max(), "odd rows", d[:, 1::2].max(), "odd cols", d[:, :, 1::2].max())
    print("  err flat rows(<9 pooled)", d[:, :9].max(), "rest", d[:, 10:].max())
    print("  per-channel max err", np.round(d.max(axis=(0, 1, 2)), 4))
    print("  per-image max err", np.round(d.max(axis=(1, 2, 3)), 4))
    # third opinion: block 0 re-computed with torch GPU ops (fp32) from the engine's stem output and the
    # engine's gradient at the block output
    torch.backends.cudnn.allow_tf32 = False
    blk = [o for o in eng.ops if o['kind'] == 'conv'][:4]          # conv1, conv2, conv3, convS
    tout = blk[3]['dst']
    s_in = stem_op['dst'].buf.permute(0, 3, 1, 2).contiguous().clone().requires_grad_(True)

    def bnrelu(t, bn):
        m = t.mean((0, 2, 3), keepdim=True)
        v = ((t - m) ** 2).mean((0, 2, 3), keepdim=True)
        g = torch.from_numpy(bn.gamma.get_value()).cuda().view(1, -1, 1, 1)
        b = torch.from_numpy(bn.beta.get_value()).cuda().view(1, -1, 1, 1)
        return torch.relu((t - m) * (g / torch.sqrt(v + 1e-4)) + b)

    def conv(t, L):
        w = torch.from_numpy(L.W.get_value()).cuda()
        b = torch.from_numpy(L.b.get_value()).cuda()
        return F.conv2d(t, w.flip(2, 3), b, stride=L.cfgParams.stride, padding=L.cfgParams.filterDim[0] // 2)
    a0 = bnrelu(s_in, blk[0]['in_bn'])
    h = conv(a0, blk[0]['layer'])
    h = conv(bnrelu(h, blk[1]['in_bn']), blk[1]['layer'])
    h = conv(bnrelu(h, blk[2]['in_bn']), blk[2]['layer'])
    outb = h + conv(a0, blk[3]['layer'])
    print("  block-0 output torch-GPU vs engine", float((outb.detach().permute(0, 2, 3, 1) - tout.buf).abs().max()))
    outb.backward(tout.grad.permute(0, 3, 1, 2).contiguous())
    ds = s_in.grad.permute(0, 2, 3, 1)
    assert float((ds - dy).abs().max() / ds.abs().max()) < 1e-5    # block-0 backward == torch-GPU fp32 autograd
    print("  dy_stem engine vs torch-GPU block-0 autograd:", float((ds - dy).abs().max() / ds.abs().max()),
          "| torch-GPU vs oracle-CPU:", float(np.abs(ds.cpu().numpy() - dyo).max() / np.abs(dyo).max()))
    for op in eng.ops:
        if op['kind'] in ('conv', 'fc') and op['dst'].grad is not None:
            ln = op['layer'].layerNum
            if col[ln].grad is None:
                continue
            gdst = op['dst']
            from_skip = [o for o in eng.ops if o['kind'] == 'conv' and o['residual'] is gdst]
            if from_skip and gdst.bn is None:
                gdst = from_skip[0]['dst']
            ge = gdst.grad.cpu().numpy()
            go_ = col[ln].grad.numpy()
            go_ = go_.transpose(0, 2, 3, 1) if go_.ndim == 4 else go_
            dd = np.abs(ge - go_) / (np.abs(go_).max() + 1e-30)
            per_img = dd.reshape(dd.shape[0], -1).max(axis=1)
            if ln in (3, 9, 10, 19, 46, 55, 92, 138, 183, 186, 187, 188) or per_img.max() > 2e-3:
                print("  grad@layer", ln, op['kind'], "per-image max err", np.round(per_img, 5))
    xs = stem_op['dst'].buf.cpu().numpy()
    xo = col[0].detach().permute(0, 2, 3, 1).numpy()
    print("  stem output err", np.abs(xs - xo).max() / np.abs(xo).max())
    assert ea < 1e-4                      # stem kernels vs torch-GPU autograd on the same upstream gradient
    assert eb2 < 3e-2                     # vs the oracle: bounded by ReLU-flip noise (see test_train_step)



@pytest.mark.parametrize("B", [16, 128])
def test_fused_bn_backward_equals_separate_kernels(B, monkeypatch):
    """The last backward-data kernel into a normalised tensor applies that BatchNorm's backward behind a grid-wide
    barrier (dpp_conv2d_dgrad_bn_bwd); DPP_FUSE_BN_BWD=0 issues dpp_conv2d_dgrad + dpp_bn_bwd_apply instead.  Same
    arithmetic in the same order: the gradient arenas of the two schedules must agree to float32 roundoff (the
    backward-weights kernels accumulate with atomics, whose order varies from run to run)."""
    D = 30
    x, y = _data(B, D)
    res = {}
    for fuse in ('1', '0'):
        monkeypatch.setenv('DPP_FUSE_BN_BWD', fuse)
        net, onet, eng = _build(0, B, 1, D, precision=1)
        eng.set_input_nchw(x)
        eng._alloc_training()
        eng.y_in.copy_(torch.from_numpy(y))
        for use_graph in (False, True):
            cost = float(eng.train_step(0.0, use_graph=use_graph).cpu()[0])
        eng.check_barriers()
        res[fuse] = (cost, eng.G.clone(), eng.launches_bn_bwd())
        eng.release()
    (c1, g1, n1), (c0, g0, n0) = res['1'], res['0']
    print("batch", B, "separate BN-backward launches: fused schedule", n1, "unfused", n0)
    assert n0 == 61 and n1 <= 4                       # 57 of the 61 run fused (3 stride-2 projections + the FC-side BN stay)
    assert c1 == c0
    err = float((g1 - g0).abs().max() / g0.abs().max())
    l2 = float((g1 - g0).norm() / g0.norm())
    print("gradient arena fused vs separate: max %.2e rel-L2 %.2e" % (err, l2))
    # (3xTF32 drops the lo x lo products and truncates: ~5e-7 per product, visible where a weight gradient is a small
    #  difference of large sums; the fp32 launches of the grouped path are the more exact side - measured 1.8e-5 / 2.7e-5)
    assert err < 5e-5 and l2 < 5e-5


@pytest.mark.parametrize("B", [16, 128])
def test_grouped_backward_weights_equal_per_layer_launches(B, monkeypatch):
    """dpp_wgrad_group_* runs the 63 backward-weights GEMMs in eight persistent launches (four tcgen05 launches by
    n-tile width: items of pixel chunks taken from a list, reduced straight into dW; four fp32 launches for the 3x3
    16- and 32-channel layers and the 1x1 64 -> 16 / 16 -> 64 layers, wgrad_simt3.cu); DPP_WGRAD_GROUP=0 launches
    dpp_conv2d_wgrad per layer on a second stream.  Same products, different summation order (and fp32 instead of
    3xTF32 products for twenty layers): gradient arenas agree to float32 roundoff."""
    D = 30
    x, y = _data(B, D)
    res = {}
    for grp in ('1', '0'):
        monkeypatch.setenv('DPP_WGRAD_GROUP', grp)
        net, onet, eng = _build(0, B, 1, D, precision=1)
        eng.set_input_nchw(x)
        eng._alloc_training()
        eng.y_in.copy_(torch.from_numpy(y))
        for use_graph in (False, True, True):
            cost = float(eng.train_step(0.0, use_graph=use_graph).cpu()[0])
        res[grp] = (cost, eng.G.clone(), eng.launches_wgrad())
        eng.release()
    (c1, g1, n1), (c0, g0, n0) = res['1'], res['0']
    print("batch", B, "backward-weights launches: grouped", n1, "per layer", n0)
    assert n0 == 63 and n1 == 8
    assert c1 == c0
    err = float((g1 - g0).abs().max() / g0.abs().max())
    l2 = float((g1 - g0).norm() / g0.norm())
    print("gradient arena grouped vs per-layer: max %.2e rel-L2 %.2e" % (err, l2))
    # (3xTF32 drops the lo x lo products and truncates: ~5e-7 per product, visible where a weight gradient is a small
    #  difference of large sums; the fp32 launches of the grouped path are the more exact side - measured 1.8e-5 / 2.7e-5)
    assert err < 5e-5 and l2 < 5e-5
