"""CPU oracle: torch-CPU restatement of the reference networks, cost, gradients and ADAM
(reference semantics are float32; working precision float64 by default, see DTYPE below).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Structure and initial weights are PINNED
against the reference's own constructors (oracle/ref_harness.py::describe_reference_net ->
tests/golden/reference_nets.json, tests/test_reference_pins.py: layer lists, dimensions,
parameter order, bit-identical initial weights).  The ARITHMETIC is pinned against the
reference's own layer / cost / T.grad / ADAM code evaluated eagerly with oracle/eager_theano.py
in place of Theano (tests/golden/reference_net_eval.npz: outputs, per-layer outputs, cost, all
gradients, BatchNorm EMA, one ADAM step).  Theano's primitive ops themselves stay UNPINNED
(Theano cannot run here): they follow SURVEY.md Appendix A in both places.

Reference files restated (all under /root/reference/src):
  net/resnet.py:45-346 (ResNetParams/ResNet types 0-4), :349-414 (res_block)
  net/poseregnet.py:44-165 (PoseRegNet types 0, 11)
  net/scalenet.py:49-193 (ScaleNet type 1)
  net/convlayer.py:230-251, net/convpoollayer.py:251-282, net/batchnormlayer.py:154-192,
  net/nonlinearitylayer.py:119, util/theano_helpers.py:61-69, net/hiddenlayer.py:136-154,
  net/dropoutlayer.py:93-104, net/layer.py:72-118 (initial values, draw order)
  trainer/poseregnettrainer.py:84-111 (cost/grads), trainer/optimizer.py:58-90 (ADAM)
"""
import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32


# --------------------------------------------------------------------------------------
# initial values, net/layer.py:72-118
# --------------------------------------------------------------------------------------
def init_vals(rng, shape, mode, act, method=None):
    """layer.py:72-118. ``act`` is the activation_str ('ReLU' or 'None')."""
    if method is None:
        method = 'He' if act == 'ReLU' else None      # layer.py:58-70
    if method == 'He':
        if mode == 'conv':
            bound = np.sqrt(2. / np.prod(shape[1:]))   # layer.py:85
            return np.asarray(rng.normal(loc=0.0, scale=bound, size=shape), dtype=f32)
        return np.asarray(rng.normal(loc=0.0, scale=0.01, size=shape), dtype=f32)  # :88
    if method is None or method == 'tanh':
        if mode == 'conv':
            bound = 1. / (np.prod(shape[1:]) + (shape[0] * np.prod(shape[2:])))  # :111
            return np.asarray(rng.uniform(low=-bound, high=bound, size=shape), dtype=f32)
        b = np.sqrt(6. / np.sum(shape))                 # layer.py:114-116
        return np.asarray(rng.uniform(low=-b, high=b, size=shape), dtype=f32)
    raise NotImplementedError(method)


# Working precision of the oracle's network arithmetic.  The reference computes in float32; the oracle
# defaults to float64 so that its own rounding - and CPU-dependent reduced-precision kernels inside
# oneDNN (a B200 host's torch-CPU fp32 conv backward differed from torch-GPU fp32 by 1.4e-2 on one
# layer while our kernels agreed with torch-GPU to 3e-7) - cannot be mistaken for kernel error.
# Initial values are always drawn and rounded in float32 exactly as the reference does.
DTYPE = torch.float64


def set_dtype(dt):
    global DTYPE
    DTYPE = dt


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DTYPE).requires_grad_(True)


# --------------------------------------------------------------------------------------
# layer records.  Each has: kind, layerNum, params (trainable, reference order),
# nontrained, and the static config needed by forward().
# --------------------------------------------------------------------------------------
class L(object):
    def __init__(self, kind, **kw):
        self.kind = kind
        self.params = []
        self.nontrained = []
        self.__dict__.update(kw)


def conv_out_hw(h, k, stride, border):
    # convlayer.py:141-163: 'valid' h-k+1, 'half' h; then ceil(/stride)
    o = h - k + 1 if border == 'valid' else h
    return int(np.ceil(o / float(stride)))


class OracleNet(object):
    """Graph container: ``layers`` in the reference's layer order (layerNum == index),
    ``program`` = wiring (the reference wires residual sums outside ``layers``)."""

    def __init__(self, rng, inputDim):
        self.rng = rng
        self.inputDim = tuple(inputDim)
        self.layers = []
        self.program = []     # ('layer', li, src, dst) | ('add', a, b, dst) | ('flatten', src, dst)
        self._nv = 1          # value 0 = network input
        self.out_vid = 0
        self.multi_inputs = None

    # -- builders -----------------------------------------------------------------
    def _newv(self):
        v = self._nv
        self._nv += 1
        return v

    def add_convpool(self, src, cin, hw, nf, k, pool, border, act, method, stride=1):
        l = L('convpool', layerNum=len(self.layers), cin=cin, cout=nf, k=k, pool=pool,
              border=border, act=act, stride=stride)
        W = init_vals(self.rng, (nf, cin, k, k), 'conv', act, method)
        l.params = [_t(W), _t(np.zeros((nf,), f32))]
        ho = conv_out_hw(hw[0], k, stride, border) // pool      # convpoollayer.py:175-176
        wo = conv_out_hw(hw[1], k, stride, border) // pool
        self.layers.append(l)
        dst = self._newv()
        self.program.append(('layer', l.layerNum, src, dst))
        return dst, nf, (ho, wo)

    def add_conv(self, src, cin, hw, nf, k, stride, border='half', act='None', method='He'):
        l = L('conv', layerNum=len(self.layers), cin=cin, cout=nf, k=k, stride=stride,
              border=border, act=act)
        W = init_vals(self.rng, (nf, cin, k, k), 'conv', act, method)
        l.params = [_t(W), _t(np.zeros((nf,), f32))]
        self.layers.append(l)
        dst = self._newv()
        self.program.append(('layer', l.layerNum, src, dst))
        return dst, nf, (conv_out_hw(hw[0], k, stride, border), conv_out_hw(hw[1], k, stride, border))

    def add_bn(self, src, c):
        l = L('bn', layerNum=len(self.layers), C=c, eps=1e-4, alpha=0.1)
        # batchnormlayer.py:148-152: trainable [beta, gamma]; non-trained [mean, inv_std]
        l.params = [_t(np.zeros((c,), f32)), _t(np.ones((c,), f32))]
        l.nontrained = [torch.zeros(c, dtype=DTYPE), torch.ones(c, dtype=DTYPE)]
        self.layers.append(l)
        dst = self._newv()
        self.program.append(('layer', l.layerNum, src, dst))
        return dst

    def add_relu(self, src):
        l = L('relu', layerNum=len(self.layers))
        self.layers.append(l)
        dst = self._newv()
        self.program.append(('layer', l.layerNum, src, dst))
        return dst

    def add_fc(self, src, n_in, n_out, act):
        l = L('fc', layerNum=len(self.layers), n_in=n_in, n_out=n_out, act=act)
        W = init_vals(self.rng, (n_in, n_out), 'fc', act)
        l.params = [_t(W), _t(np.zeros((n_out,), f32))]
        self.layers.append(l)
        dst = self._newv()
        self.program.append(('layer', l.layerNum, src, dst))
        return dst

    def add_dropout(self, src, p=0.3):
        l = L('dropout', layerNum=len(self.layers), p=p)
        self.rng.randint(999999)          # dropoutlayer.py:96: one draw at construction
        self.layers.append(l)
        dst = self._newv()
        self.program.append(('layer', l.layerNum, src, dst))
        return dst

    def add_add(self, a, b):
        dst = self._newv()
        self.program.append(('add', a, b, dst))
        return dst

    def add_flatten(self, src):
        dst = self._newv()
        self.program.append(('flatten', src, dst))
        return dst

    def add_concat(self, srcs):
        dst = self._newv()
        self.program.append(('concat', list(srcs), dst))
        return dst

    # -- parameter views ----------------------------------------------------------
    @property
    def params(self):
        """netbase.py:152-165: layer order, per layer [W,b] / [beta,gamma]."""
        return [p for l in self.layers for p in l.params]

    @property
    def weights(self):
        return [l.params[0] for l in self.layers if l.kind in ('conv', 'convpool', 'fc')]

    def has_dropout(self):
        return any(l.kind == 'dropout' for l in self.layers)

    # -- forward ------------------------------------------------------------------
    def forward(self, x, deterministic, masks=None, collect=None, relu_masks=None):
        """x: torch tensor (B,C,H,W) fp32 or list of tensors (ScaleNet).
        Returns (out, bn_stats) with bn_stats[layerNum] = (batch mean, batch inv_std).
        ``collect`` (optional dict) receives every layer output by layerNum."""
        vals = {}
        if isinstance(x, (list, tuple)):
            x = [xi.to(DTYPE) for xi in x]
            for i, xi in enumerate(x):
                vals[-(i + 1)] = xi
            vals[0] = x[0]
        else:
            vals[0] = x.to(DTYPE)
        if masks is not None:
            masks = [m.to(DTYPE) for m in masks]
        bn_stats = {}
        mi = 0
        for st in self.program:
            if st[0] == 'add':
                vals[st[3]] = vals[st[1]] + vals[st[2]]
                continue
            if st[0] == 'flatten':
                vals[st[2]] = vals[st[1]].flatten(1)        # .flatten(2) in Theano: (B, C*H*W)
                continue
            if st[0] == 'concat':
                vals[st[2]] = torch.cat([vals[v] for v in st[1]], dim=1)
                continue
            l = self.layers[st[1]]
            a = vals[st[2]]
            if l.kind in ('conv', 'convpool'):
                W, b = l.params
                pad = l.k // 2 if l.border == 'half' else 0
                # theano conv2d is a true convolution (filter_flip=True): flip both axes
                o = F.conv2d(a, W.flip(2, 3), stride=l.stride, padding=pad)
                if l.kind == 'convpool' and l.pool > 1:
                    o = F.max_pool2d(o, l.pool, l.pool)     # ignore_border=True == floor
                o = o + b.view(1, -1, 1, 1)                 # bias after pooling (:276-277)
                if l.act == 'ReLU':
                    o = torch.clamp_min(o, 0)
            elif l.kind == 'bn':
                beta, gamma = l.params
                if deterministic:
                    mean, inv_std = l.nontrained
                else:
                    mean = a.mean(dim=(0, 2, 3))
                    var = ((a - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
                    inv_std = 1.0 / torch.sqrt(var + l.eps)
                    bn_stats[l.layerNum] = (mean.detach(), inv_std.detach())
                o = (a - mean.view(1, -1, 1, 1)) * (gamma * inv_std).view(1, -1, 1, 1) \
                    + beta.view(1, -1, 1, 1)
            elif l.kind == 'relu':
                if relu_masks is not None and l.layerNum in relu_masks:
                    # test aid: ReLU with the on/off decisions of another implementation imposed, so
                    # that boundary pre-activations (|z| ~ roundoff) cannot flip between the two
                    o = a * relu_masks[l.layerNum].to(DTYPE)
                else:
                    o = torch.clamp_min(a, 0)               # T.maximum(x, 0)
            elif l.kind == 'fc':
                W, b = l.params
                o = a @ W + b
                if l.act == 'ReLU':
                    if relu_masks is not None and l.layerNum in relu_masks:
                        o = o * relu_masks[l.layerNum].to(DTYPE)    # imposed decisions (see 'relu' above)
                    else:
                        o = torch.clamp_min(o, 0)
            elif l.kind == 'dropout':
                if deterministic:
                    o = (1.0 - l.p) * a                     # dropoutlayer.py:104
                else:
                    o = masks[mi] * a
                    mi += 1
            else:
                raise NotImplementedError(l.kind)
            vals[st[3]] = o
            if collect is not None:
                collect[l.layerNum] = o
        return vals[self.out_vid], bn_stats

    def apply_bn_ema(self, bn_stats):
        """batchnormlayer.py:164-172: r = 0.9 r + 0.1 stat, for mean and inv_std."""
        for ln, (m, s) in bn_stats.items():
            l = self.layers[ln]
            a = float(f32(l.alpha))
            l.nontrained[0] = (1. - a) * l.nontrained[0] + a * m
            l.nontrained[1] = (1. - a) * l.nontrained[1] + a * s


# --------------------------------------------------------------------------------------
# graph constructors
# --------------------------------------------------------------------------------------
def _res_block(net, src, cin, hw, out_f, stride):
    """resnet.py:349-414."""
    nb = out_f // 4
    if cin == out_f:
        v = net.add_bn(src, cin)
        v = net.add_relu(v)
        v, c, hw1 = net.add_conv(v, cin, hw, nb, 1, 1)
        v = net.add_bn(v, c)
        v = net.add_relu(v)
        v, c, hw1 = net.add_conv(v, c, hw1, nb, 3, 1)
        v = net.add_bn(v, c)
        v = net.add_relu(v)
        v, c, hw1 = net.add_conv(v, c, hw1, out_f, 1, 1)
        return net.add_add(src, v), out_f, hw1
    v = net.add_bn(src, cin)
    a = net.add_relu(v)
    v, c, hw1 = net.add_conv(a, cin, hw, nb, 1, stride)
    v = net.add_bn(v, c)
    v = net.add_relu(v)
    v, c, hw1 = net.add_conv(v, c, hw1, nb, 3, 1)
    v = net.add_bn(v, c)
    v = net.add_relu(v)
    v3, c, hw1 = net.add_conv(v, c, hw1, out_f, 1, 1)
    sc, c, hw2 = net.add_conv(a, cin, hw, out_f, 1, stride)      # layers[-8] = first ReLU
    assert hw1 == hw2
    return net.add_add(v3, sc), out_f, hw1


def build_resnet(rng, type=0, nChan=1, wIn=128, hIn=128, batchSize=128, numJoints=16, nDims=3):
    """resnet.py:100-340."""
    net = OracleNet(rng, (batchSize, nChan, hIn, wIn))
    n = (47 - 2) // 9                                       # resnet.py:124 (py2 int division)
    stages = [32, 64, 128, 128, 128] if type == 3 else [32, 64, 128, 256, 256]
    v, c, hw = net.add_convpool(0, nChan, (hIn, wIn), stages[0], 5, 2, 'half', 'None', 'He')
    for s in range(1, 5):
        v, c, hw = _res_block(net, v, c, hw, stages[s], 2)
        for _ in range(1, n):
            v, c, hw = _res_block(net, v, c, hw, stages[s], 1)
    v = net.add_bn(v, c)
    v = net.add_relu(v)
    v = net.add_flatten(v)
    nfeat = c * hw[0] * hw[1]
    out = numJoints * nDims
    if type in (0, 1):
        v = net.add_fc(v, nfeat, 1024, 'ReLU')
        v = net.add_fc(v, 1024, 1024, 'ReLU')
        if type == 1:
            v = net.add_fc(v, 1024, 30, 'None')
            v = net.add_fc(v, 30, out, 'None')
        else:
            v = net.add_fc(v, 1024, out, 'None')
    elif type in (2, 3, 4):
        v = net.add_fc(v, nfeat, 1024, 'ReLU')
        v = net.add_dropout(v)
        v = net.add_fc(v, 1024, 1024, 'ReLU')
        v = net.add_dropout(v)
        if type == 4:
            v = net.add_fc(v, 1024, 30, 'None')
            v = net.add_fc(v, 30, out, 'None')
        else:
            v = net.add_fc(v, 1024, out, 'None')
    else:
        raise NotImplementedError()
    net.out_vid = v
    return net


def build_poseregnet(rng, type=0, nChan=1, wIn=128, hIn=128, batchSize=128, numJoints=16, nDims=3):
    """poseregnet.py:60-143 (ConvPool defaults: border 'valid', init by activation)."""
    net = OracleNet(rng, (batchSize, nChan, hIn, wIn))
    v, c, hw = net.add_convpool(0, nChan, (hIn, wIn), 8, 5, 4, 'valid', 'ReLU', None)
    v, c, hw = net.add_convpool(v, c, hw, 8, 5, 2, 'valid', 'ReLU', None)
    v, c, hw = net.add_convpool(v, c, hw, 8, 3, 1, 'valid', 'ReLU', None)
    v = net.add_flatten(v)
    v = net.add_fc(v, c * hw[0] * hw[1], 1024, 'ReLU')
    v = net.add_dropout(v)
    v = net.add_fc(v, 1024, 1024, 'ReLU')
    v = net.add_dropout(v)
    if type == 0:
        v = net.add_fc(v, 1024, numJoints * nDims, 'None')
    elif type == 11:
        v = net.add_fc(v, 1024, 30, 'None')
        v = net.add_fc(v, 30, numJoints * nDims, 'None')
    else:
        raise NotImplementedError()
    net.out_vid = v
    return net


def build_scalenet(rng, type=1, nChan=1, wIn=128, hIn=128, batchSize=128, numJoints=1, nDims=3, resizeFactor=2):
    """scalenet.py:49-193 (type 1): three towers of three 'valid' ConvPool layers on inputs -1, -2, -3 (the crop
    and its centre crops), flatten + concatenate, FC1024 - dropout - FC1024 - dropout - FC(J*D)."""
    if type != 1:
        raise NotImplementedError()
    net = OracleNet(rng, (batchSize, nChan, hIn, wIn))
    f = resizeFactor
    net.multi_inputs = [(batchSize, nChan, hIn, wIn), (batchSize, nChan, hIn // f, wIn // f),
                        (batchSize, nChan, hIn // f ** 2, wIn // f ** 2)]
    towers = [[(5, 4), (5, 2), (3, 1)], [(5, 2), (5, 2), (3, 1)], [(5, 2), (5, 1), (3, 1)]]
    flats, nfeat = [], 0
    for t, spec in enumerate(towers):
        v, c, hw = -(t + 1), nChan, net.multi_inputs[t][2:]
        for k, pool in spec:
            v, c, hw = net.add_convpool(v, c, hw, 8, k, pool, 'valid', 'ReLU', None)
        flats.append((v, c * hw[0] * hw[1]))
        nfeat += c * hw[0] * hw[1]
    # the reference flattens when the first hidden layer is wired, after all ConvPool layers exist
    v = net.add_concat([net.add_flatten(fv) for fv, _ in flats])
    v = net.add_fc(v, nfeat, 1024, 'ReLU')
    v = net.add_dropout(v)
    v = net.add_fc(v, 1024, 1024, 'ReLU')
    v = net.add_dropout(v)
    v = net.add_fc(v, 1024, numJoints * nDims, 'None')
    net.out_vid = v
    return net


def append_pca_layer(net, components, mean):
    """main_nyu_posereg_embedding.py:148-158: HiddenLayer(30 -> 3J, linear) appended after
    training; W = pca.components_, b = pca.mean_.  (The constructor's own random init draws
    from the rng before being overwritten.)"""
    n_in, n_out = components.shape
    v = net.add_fc(net.out_vid, n_in, n_out, 'None')
    l = net.layers[-1]
    l.params = [_t(np.asarray(components, f32)), _t(np.asarray(mean, f32))]
    net.out_vid = v
    return net


# --------------------------------------------------------------------------------------
# cost / gradients / ADAM
# --------------------------------------------------------------------------------------
def cost_fn(net, out, y, batch_size, numJoints, nDims, weightreg=0.0):
    """poseregnettrainer.py:84-107."""
    if numJoints == 1 and nDims == 1:
        cost = ((out.reshape(batch_size, nDims) - y) ** 2).mean(dim=1)
    elif numJoints == 1:
        cost = ((out.reshape(batch_size, nDims) - y.to(out.dtype)) ** 2).sum(dim=1)
    else:
        cost = ((out.reshape(batch_size, numJoints, nDims) - y) ** 2).sum(dim=2).mean(dim=1)
    cost = cost.mean()
    if not net.has_dropout():
        reg = 0.
        for W in net.weights:
            reg = reg + float(weightreg) * (W ** 2).sum()
        cost = cost + reg
    return cost


class Adam(object):
    """optimizer.py:58-90, with floatX=float32 constant folding:
    gamma = 1-1e-8 -> 1.0f so beta1_t == 0.9f; every constant is fp32."""

    def __init__(self, params):
        self.params = params
        self.m = [np.zeros(tuple(p.shape), f32) for p in params]
        self.v = [np.zeros(tuple(p.shape), f32) for p in params]
        self.t = f32(1.0)

    def step(self, grads, lr):
        b1, b2, eps, one = f32(0.9), f32(0.999), f32(1e-8), f32(1.0)
        lr = f32(lr)
        b1t = b1                                   # beta1 * 1.0f ** (t-1)
        c1 = one - np.power(b1, self.t, dtype=f32)
        c2 = one - np.power(b2, self.t, dtype=f32)
        for i, (p, g) in enumerate(zip(self.params, grads)):
            g = g.detach().numpy().astype(f32)
            m = b1t * self.m[i] + (one - b1t) * g
            v = b2 * self.v[i] + (one - b2) * (g * g)
            mh = m / c1
            vh = v / c2
            w = p.detach().numpy().astype(f32) - (lr * mh) / (np.sqrt(vh) + eps)
            self.m[i], self.v[i] = m.astype(f32), v.astype(f32)
            with torch.no_grad():
                p.copy_(torch.from_numpy(w.astype(f32)).to(p.dtype))
        self.t = f32(self.t + one)


def train_step(net, adam, x, y, lr, numJoints, nDims, weightreg=0.0, masks=None, relu_masks=None):
    """One ``train_model`` call (poseregnettrainer.py:146-160): cost on batch statistics,
    T.grad through them, ADAM, BN running-stat EMA; all from the old shared values."""
    for p in net.params:
        p.grad = None
    out, stats = net.forward(x, deterministic=False, masks=masks, relu_masks=relu_masks)
    cost = cost_fn(net, out, y, x.shape[0] if not isinstance(x, (list, tuple)) else x[0].shape[0],
                   numJoints, nDims, weightreg)
    grads = torch.autograd.grad(cost, net.params, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, net.params)]
    adam.step(grads, lr)
    net.apply_bn_ema(stats)
    return float(cost.detach()), out.detach(), grads


def lr_of_ep(learning_rate, ep):
    """nettrainer.py:54."""
    if ep <= 1:
        return f32(learning_rate / 10.)
    if 1 < ep <= 2:
        return f32(learning_rate / 3.)
    return f32(learning_rate * np.exp(-0.04 * ep))
