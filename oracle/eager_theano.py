"""An EAGER numeric stand-in for the parts of Theano 0.9 the reference's hot path touches.  TEST INFRASTRUCTURE ONLY
(build container; used by oracle/ref_harness.py to run the reference's own layer / network / trainer code).

Theano cannot be installed here (Python 3.12, no network), so the reference's graphs cannot be compiled.  But the
reference builds its graphs by ordinary Python calls (``conv2d(...)``, ``T.mean(...)``, ``x.dimshuffle(...)``,
``T.grad(cost, params)``): if those calls COMPUTE instead of recording, running the reference's constructors on a
concrete input evaluates the reference's own code - layer wiring, BatchNorm formula, bias / pooling / activation
order, flatten order, cost expression, ADAM update expressions - eagerly.  Tensors are torch-CPU float64 tensors
(autograd gives ``T.grad``); ``theano.shared`` variables are leaf tensors.

What this is NOT: Theano.  The primitive ops below follow Theano 0.9's DOCUMENTED semantics:
  * ``conv2d`` is a true convolution (filters flipped), ``border_mode`` 'valid' | 'half' (pad k//2) | 'full',
    ``subsample`` = stride;
  * ``pool_2d(ds, ignore_border=True, mode='max')`` = non-overlapping max pooling, floor;
  * ``T.var`` is the biased variance; ``T.nnet.batch_normalization(x, gamma, beta, mean, std)`` =
    (x - mean) * (gamma / std) + beta;
  * ``T.switch`` / ``ifelse`` on a 0-d shared flag selects a branch.
So a pin obtained through this module covers the reference's Python code around the primitives, not the primitives'
own implementation, and it computes in float64 (the reference ran float32)."""
import types

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64
FEED = {}                      # name -> numpy array for T.tensor4(name) / T.matrix(name) ... placeholders


def _raw(x):
    if isinstance(x, ET):
        return x.t
    if isinstance(x, (bool, int, float, np.floating, np.integer)):
        return float(x)
    return torch.as_tensor(np.asarray(x), dtype=DT)


class ET(object):
    """eager tensor"""
    __array_priority__ = 1000.

    def __init__(self, t, name=None):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t), dtype=DT)
        self.t = t
        self.name = name
        self.default_update = None

    ndim = property(lambda self: self.t.dim())
    shape = property(lambda self: tuple(self.t.shape))
    dtype = 'float32'
    broadcastable = property(lambda self: tuple(False for _ in self.t.shape))

    def eval(self):
        return self.t.detach().numpy()

    def __add__(self, o): return ET(self.t + _raw(o))
    def __radd__(self, o): return ET(_raw(o) + self.t)
    def __sub__(self, o): return ET(self.t - _raw(o))
    def __rsub__(self, o): return ET(_raw(o) - self.t)
    def __mul__(self, o): return ET(self.t * _raw(o))
    def __rmul__(self, o): return ET(_raw(o) * self.t)
    def __truediv__(self, o): return ET(self.t / _raw(o))
    def __rtruediv__(self, o): return ET(_raw(o) / self.t)
    def __pow__(self, o): return ET(self.t ** _raw(o))
    def __rpow__(self, o): return ET(_raw(o) ** self.t)
    def __neg__(self): return ET(-self.t)
    def __getitem__(self, idx): return ET(self.t[idx])
    def __bool__(self): return bool(self.t.item() != 0)

    def flatten(self, ndim=1):
        return ET(self.t.reshape(tuple(self.t.shape[:ndim - 1]) + (-1,)))

    def reshape(self, shape, ndim=None):
        return ET(self.t.reshape(tuple(int(s) for s in shape)))

    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (list, tuple)):
            pattern = tuple(pattern[0])
        t = self.t.permute(*[p for p in pattern if p != 'x']) if self.t.dim() else self.t
        for i, p in enumerate(pattern):
            if p == 'x':
                t = t.unsqueeze(i)
        return ET(t)

    def sum(self, axis=None, keepdims=False): return sum_(self, axis, keepdims)
    def mean(self, axis=None, keepdims=False): return mean(self, axis, keepdims)
    def max(self, axis=None, keepdims=False): return max_(self, axis, keepdims)
    def norm(self, L, axis=None): return ET(torch.linalg.vector_norm(self.t, ord=L))
    T = property(lambda self: ET(self.t.t()))


class Shared(ET):
    _count = 0

    def __init__(self, value=None, name=None, borrow=False, **kw):
        arr = np.asarray(value)
        self._np_dtype = arr.dtype
        t = torch.tensor(arr.astype(np.float64), dtype=DT)
        if arr.dtype.kind == 'f':
            t.requires_grad_(True)
        ET.__init__(self, t, name)
        self.auto_name = 'auto_%d' % Shared._count
        Shared._count += 1

    def get_value(self, borrow=False):
        return self.t.detach().numpy().astype(self._np_dtype)

    def set_value(self, v, borrow=False):
        v = torch.as_tensor(np.asarray(v, dtype=np.float64), dtype=DT)
        with torch.no_grad():
            if tuple(v.shape) == tuple(self.t.shape):
                self.t.copy_(v)
            else:
                self.t = v.clone().requires_grad_(self.t.requires_grad)


# ---- reductions / elementwise -----------------------------------------------------------------------------------
def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else (int(axis),)


def mean(x, axis=None, keepdims=False, **kw):
    a = _axes(axis)
    return ET(x.t.mean() if a is None else x.t.mean(dim=a, keepdim=keepdims))


def var(x, axis=None, keepdims=False, **kw):
    a = _axes(axis)
    return ET(x.t.var(unbiased=False) if a is None else x.t.var(dim=a, unbiased=False, keepdim=keepdims))


def sum_(x, axis=None, keepdims=False, **kw):
    a = _axes(axis)
    return ET(x.t.sum() if a is None else x.t.sum(dim=a, keepdim=keepdims))


def max_(x, axis=None, keepdims=False, **kw):
    a = _axes(axis)
    return ET(x.t.max() if a is None else x.t.amax(dim=a, keepdim=keepdims))


def switch(cond, a, b):
    if isinstance(cond, ET) and cond.t.dim() == 0:
        return _et(a) if bool(cond) else _et(b)
    return ET(torch.where(_raw(cond) != 0, _raw(a), _raw(b)))


def _et(x):
    return x if isinstance(x, ET) else ET(_raw(x) if not isinstance(_raw(x), float) else torch.tensor(_raw(x), dtype=DT))


def batch_normalization(inputs, gamma, beta, mean, std, mode='low_mem'):
    return ET((inputs.t - _raw(mean)) * (_raw(gamma) / _raw(std)) + _raw(beta))


def conv2d(input, filters, input_shape=None, filter_shape=None, border_mode='valid', subsample=(1, 1),
           filter_flip=True, image_shape=None, **kw):
    w = filters.t
    kh, kw_ = int(w.shape[2]), int(w.shape[3])
    if border_mode in ('half', 'same'):
        pad = (kh // 2, kw_ // 2)
    elif border_mode == 'valid':
        pad = (0, 0)
    elif border_mode == 'full':
        pad = (kh - 1, kw_ - 1)
    else:
        pad = tuple(int(p) for p in border_mode)
    if filter_flip:
        w = w.flip(2, 3)                      # theano conv2d is a true convolution
    return ET(F.conv2d(input.t, w, stride=tuple(int(s) for s in subsample), padding=pad))


def pool_2d(input, ds=None, ignore_border=None, st=None, padding=(0, 0), mode='max', ws=None, **kw):
    ds = tuple(int(d) for d in (ds if ds is not None else ws))
    if ignore_border is not True or mode != 'max' or st not in (None, ds):
        raise NotImplementedError("pool_2d stand-in: only non-overlapping max pooling with ignore_border=True")
    if ds == (1, 1):
        return ET(input.t)
    return ET(F.max_pool2d(input.t, ds, stride=ds))


def grad(cost, wrt, **kw):
    single = isinstance(wrt, ET)
    ws = [wrt] if single else list(wrt)          # e.g. the dict_values of NetBase.params (netbase.py:165)
    gs = torch.autograd.grad(cost.t, [w.t for w in ws], retain_graph=True, allow_unused=True)
    out = [ET(torch.zeros_like(w.t) if g is None else g) for g, w in zip(gs, ws)]
    return out[0] if single else out


def _placeholder(name=None, **kw):
    if name not in FEED:
        raise KeyError("eager theano: no value fed for symbolic input %r (set oracle.eager_theano.FEED)" % (name,))
    v = FEED[name]
    if isinstance(v, list):                      # several placeholders share a name (poseregnettrainer.py: 'y' twice)
        v = v.pop(0)
    return ET(torch.as_tensor(np.asarray(v, dtype=np.float64), dtype=DT), name)


class _Streams(object):
    def __init__(self, seed=None, **kw):
        self.gen = torch.Generator().manual_seed(int(seed) if seed is not None else 0)

    def binomial(self, size=None, n=1, p=0.5, dtype=None, **kw):
        shape = tuple(int(s) for s in size)
        return ET((torch.rand(shape, generator=self.gen, dtype=DT) < p).to(DT))


def build_modules():
    """Returns {module name: module} to place in sys.modules while reference code runs."""
    theano = types.ModuleType('theano')
    theano.shared = lambda value=None, name=None, borrow=False, **kw: Shared(value, name, borrow)
    theano.clone = lambda x, share_inputs=True, **kw: ET(x.t, getattr(x, 'name', None))
    theano.config = types.SimpleNamespace(floatX='float32')

    def _no_function(*a, **k):
        raise NotImplementedError("eager theano: theano.function is not available (nothing is compiled)")
    theano.function = _no_function
    T = types.ModuleType('theano.tensor')
    for n in ('tensor4', 'tensor3', 'matrix', 'vector', 'scalar', 'ftensor4', 'fmatrix', 'fvector', 'ivector', 'lscalar',
              'dtensor4'):
        setattr(T, n, _placeholder)
    T.mean, T.var, T.sum, T.max = mean, var, sum_, max_
    T.sqrt = lambda x: ET(torch.sqrt(_et(x).t))
    T.inv = lambda x: ET(1. / _et(x).t)
    T.sqr = lambda x: ET(_et(x).t ** 2)
    T.exp = lambda x: ET(torch.exp(_et(x).t))
    T.log = lambda x: ET(torch.log(_et(x).t))
    T.abs_ = lambda x: ET(torch.abs(_et(x).t))
    T.tanh = lambda x: ET(torch.tanh(_et(x).t))
    T.maximum = lambda a, b: ET(torch.maximum(_et(a).t, _et(b).t.expand_as(_et(a).t) if _et(b).t.dim() == 0 else _et(b).t))
    T.minimum = lambda a, b: ET(torch.minimum(_et(a).t, _et(b).t.expand_as(_et(a).t) if _et(b).t.dim() == 0 else _et(b).t))
    T.dot = lambda a, b: ET(_et(a).t @ _et(b).t)
    T.concatenate = lambda xs, axis=0: ET(torch.cat([x.t for x in xs], dim=axis))
    T.reshape = lambda x, shape, ndim=None: x.reshape(shape)
    T.flatten = lambda x, outdim=1: x.flatten(outdim)
    T.cast = lambda x, dtype: _et(x)
    T.switch = switch
    T.eq = lambda a, b: ET((_et(a).t == _raw(b)).to(DT))
    T.neq = lambda a, b: ET((_et(a).t != _raw(b)).to(DT))
    T.gt = lambda a, b: ET((_et(a).t > _raw(b)).to(DT))
    T.lt = lambda a, b: ET((_et(a).t < _raw(b)).to(DT))
    T.zeros_like = lambda x: ET(torch.zeros_like(x.t))
    T.ones_like = lambda x: ET(torch.ones_like(x.t))
    T.grad = grad
    nnet = types.ModuleType('theano.tensor.nnet')
    nnet.conv2d = conv2d
    nnet.batch_normalization = batch_normalization
    nnet.relu = lambda x, alpha=0: ET(torch.where(x.t > 0, x.t, alpha * x.t))
    nnet.sigmoid = lambda x: ET(torch.sigmoid(x.t))
    nnet.softmax = lambda x: ET(torch.softmax(x.t, dim=-1))
    T.nnet = nnet
    signal = types.ModuleType('theano.tensor.signal')
    pool = types.ModuleType('theano.tensor.signal.pool')
    pool.pool_2d = pool_2d
    signal.pool = pool
    T.signal = signal
    ifelse = types.ModuleType('theano.ifelse')
    ifelse.ifelse = switch
    sandbox = types.ModuleType('theano.sandbox')
    rng_mrg = types.ModuleType('theano.sandbox.rng_mrg')
    rng_mrg.MRG_RandomStreams = _Streams
    neighbours = types.ModuleType('theano.sandbox.neighbours')
    sandbox.rng_mrg, sandbox.neighbours = rng_mrg, neighbours
    theano.tensor, theano.ifelse, theano.sandbox = T, ifelse, sandbox
    return {'theano': theano, 'theano.tensor': T, 'theano.tensor.nnet': nnet, 'theano.tensor.signal': signal,
            'theano.tensor.signal.pool': pool, 'theano.ifelse': ifelse, 'theano.sandbox': sandbox,
            'theano.sandbox.rng_mrg': rng_mrg, 'theano.sandbox.neighbours': neighbours}
