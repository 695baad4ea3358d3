"""Executes the reference's OWN Python sources from /root/reference in this (Python 3.12 / NumPy 2 / cv2 4.13)
container, to pin the oracle and to generate golden vectors.  TEST INFRASTRUCTURE ONLY; build-container only
(/root/reference does not exist on the GPU box - nothing under tests/ -m gpu, smoke() or bench.py imports this).

The reference is Python 2.7.  Nothing is copied into the repository: the sources are read where they lie, given a
mechanical py2 -> py3 pass IN MEMORY and exec'd:
  * ``print x`` statements -> ``print(x)`` (regex on the statement lines),
  * every ``a / b`` is rewritten (AST) to ``_py2div(a, b)``: floor division when both operands are integers
    (Python / NumPy ints), true division otherwise - Python 2's classic division, which is load-bearing in
    comToTransform / cropArea3D (``hb * dsize[0] / wb``, SURVEY 8a "py2 hazards"),
  * ``xrange`` -> ``range``; modules that are absent here and unused by the functions we call (progressbar, cPickle,
    sharedmem, theano, matplotlib ...) are stubbed; ``numpy.float`` (removed alias of ``float``) and the array-returning
    ``scipy.stats.mode`` of older SciPy are shimmed.
What this cannot emulate: NumPy 1.x value-based casting.  Under NumPy 2 (NEP 50) ``float32_scalar * python_float``
stays float32 where the reference-era NumPy promoted to float64, so on float32 CoMs the reference-as-run-here can
differ from the reference-as-published in the last bit of a bound; tests/test_reference_pins.py therefore pins the
oracle on float64 CoMs (both NumPy generations agree there) bit-exactly and reports the float32 cases separately."""
import ast
import builtins
import os
import re
import sys
import types

REF_SRC = os.environ.get('DPP_REFERENCE_SRC', '/root/reference/src')


def available():
    return os.path.isdir(REF_SRC)


def _py2div(a, b):
    import numpy as np
    ints = (int, np.integer)
    if isinstance(a, ints) and isinstance(b, ints) and not isinstance(a, bool) and not isinstance(b, bool):
        return a // b
    if isinstance(a, np.ndarray) and isinstance(b, (np.ndarray,) + ints) and a.dtype.kind in 'iu' and \
            (not isinstance(b, np.ndarray) or b.dtype.kind in 'iu'):
        return a // b
    return a / b


class _Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id='_py2div', ctx=ast.Load()), args=[node.left, node.right],
                                              keywords=[]), node)
        return node

    def visit_AugAssign(self, node):
        self.generic_visit(node)
        return node      # ``x /= y`` on arrays is in-place true division in both Pythons for float arrays


_PRINT = re.compile(r'^(\s*)print (?!\()(.*)$', re.M)


def py3_source(path):
    src = open(path).read()
    src = _PRINT.sub(lambda m: '%sprint(%s)' % (m.group(1), m.group(2)), src)
    src = re.sub(r'^(\s*)print$', r'\1print()', src, flags=re.M)
    return src


def load_module(modname, relpath, extra_globals=None):
    """exec one reference source file as module ``modname`` (registered in sys.modules)."""
    path = os.path.join(REF_SRC, relpath)
    tree = ast.parse(py3_source(path), filename=path)
    tree = ast.fix_missing_locations(_Div().visit(tree))
    mod = types.ModuleType(modname)
    mod.__file__ = path
    mod.__dict__['_py2div'] = _py2div
    mod.__dict__['xrange'] = range
    if extra_globals:
        mod.__dict__.update(extra_globals)
    sys.modules[modname] = mod
    exec(compile(tree, path, 'exec'), mod.__dict__)
    return mod


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def reference_modules():
    """Returns dict(transformations, handdetector, importers, basetypes) - the reference's modules, exec'd."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference sources not found at %s" % REF_SRC)
    import numpy
    if not hasattr(numpy, 'float'):
        numpy.float = float                      # removed alias the reference uses (numpy.float == float)
    saved = {k: sys.modules.get(k) for k in ('data', 'data.transformations', 'data.basetypes', 'data.importers', 'data.dataset', 'util',
                                             'util.handdetector', 'progressbar', 'cPickle', 'net', 'trainer')}
    try:
        import pickle
        _stub('progressbar')
        _stub('cPickle', **pickle.__dict__)
        pkg_d, pkg_u = types.ModuleType('data'), types.ModuleType('util')
        pkg_d.__path__, pkg_u.__path__ = [], []
        sys.modules['data'], sys.modules['util'] = pkg_d, pkg_u
        _loaded['transformations'] = load_module('data.transformations', 'data/transformations.py')
        _loaded['basetypes'] = load_module('data.basetypes', 'data/basetypes.py')
        _loaded['handdetector'] = load_module('util.handdetector', 'util/handdetector.py')
        _loaded['importers'] = load_module('data.importers', 'data/importers.py')
        _loaded['dataset'] = load_module('data.dataset', 'data/dataset.py')
        # scipy < 1.11 returned arrays from stats.mode (the reference indexes [0][0], handdetector.py:128-130)
        import scipy.stats
        _loaded['handdetector'].stats = types.SimpleNamespace(mode=lambda a: scipy.stats.mode(a, keepdims=True))
    finally:
        for k, v in saved.items():               # give the product's own ``data`` / ``util`` packages back
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return _loaded


def _extract_def(src, name):
    """Source text of ``def name(`` (first occurrence) up to the next line that is not indented deeper."""
    lines = src.split('\n')
    start = next(i for i, l in enumerate(lines) if re.match(r'\s*def %s\(' % re.escape(name), l))
    indent = len(lines[start]) - len(lines[start].lstrip())
    end = start + 1
    while end < len(lines):
        l = lines[end]
        if l.strip() and (len(l) - len(l.lstrip())) <= indent:
            break
        end += 1
    import textwrap
    return textwrap.dedent('\n'.join(lines[start:end]))


def reference_function(relpath, name, globs):
    """Extract ONE function / method (e.g. 'augmentCrop' of NetTrainer) from a reference file whose module-level
    imports cannot be satisfied here (theano) and which has multi-line py2 print statements elsewhere; exec it
    stand-alone with ``globs`` as its globals."""
    path = os.path.join(REF_SRC, relpath)
    text = _PRINT.sub(lambda m: '%sprint(%s)' % (m.group(1), m.group(2)), _extract_def(open(path).read(), name))
    tree = ast.fix_missing_locations(_Div().visit(ast.parse(text, filename=path)))
    g = dict(globs)
    g.update(_py2div=_py2div, xrange=range, __builtins__=builtins)
    exec(compile(tree, path, 'exec'), g)
    return g[name]


# ------------------------------------------------------------------------------------------------------------
# the reference's network classes under a stand-in ``theano``
# ------------------------------------------------------------------------------------------------------------
class _Shared(object):
    """theano.shared stand-in: keeps the value; every symbolic use yields an inert placeholder."""
    _count = 0

    def __init__(self, value=None, name=None, borrow=False, **kw):
        import numpy as np
        self.value = np.asarray(value)
        self.name = name
        self.auto_name = 'auto_%d' % _Shared._count
        _Shared._count += 1
        self.broadcastable = (False,) * self.value.ndim

    def get_value(self, borrow=False):
        return self.value

    def set_value(self, v, borrow=False):
        import numpy as np
        self.value = np.asarray(v)

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        from unittest import mock
        return mock.MagicMock()

    def _sym(self, *a, **k):
        from unittest import mock
        return mock.MagicMock()
    __add__ = __radd__ = __mul__ = __rmul__ = __sub__ = __rsub__ = __truediv__ = __rtruediv__ = __pow__ = __neg__ = _sym


_NET_FILES = ['util/theano_helpers.py', 'net/layerparams.py', 'net/layer.py', 'net/convlayer.py', 'net/convpoollayer.py',
              'net/hiddenlayer.py', 'net/poollayer.py', 'net/dropoutlayer.py', 'net/batchnormlayer.py',
              'net/nonlinearitylayer.py', 'net/netbase.py', 'net/resnet.py', 'net/poseregnet.py', 'net/scalenet.py']


class _reference_net_modules(object):
    """Context manager: the reference's net/*.py exec'd with ``theano_modules`` (name -> module) standing in for
    Theano; restores sys.modules (the product's own ``net`` / ``util`` packages) and the NumPy / inspect shims."""

    def __init__(self, theano_modules):
        self.fake = dict(theano_modules)

    def __enter__(self):
        import inspect
        import pickle
        import numpy as np
        if not available():
            raise RuntimeError("reference sources not found at %s" % REF_SRC)
        self.fake['cPickle'] = pickle
        names = [r[:-3].replace('/', '.') for r in _NET_FILES]
        self.saved = {k: sys.modules.get(k) for k in list(self.fake) + names + ['net', 'util']}

        class _Cast(dict):                            # numpy.cast[dtype](value), removed in NumPy 2
            def __missing__(self, k):
                return lambda v: np.asarray(v, dtype=k)[()]
        self.had_cast = 'cast' in np.__dict__
        self.old_getargspec = getattr(inspect, 'getargspec', None)
        sys.modules.update(self.fake)
        np.cast = _Cast()
        inspect.getargspec = inspect.getfullargspec            # removed in Python 3.11
        for pkg in ('net', 'util'):
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
        return {n: load_module(n, r) for n, r in zip(names, _NET_FILES)}

    def __exit__(self, *exc):
        import inspect
        import numpy as np
        for k, v in self.saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if not self.had_cast:
            del np.cast
        if self.old_getargspec is None:
            del inspect.getargspec
        return False


def describe_reference_net(kind, **cfg):
    """Builds one of the reference's networks (``kind`` in 'ResNet', 'PoseRegNet', 'ScaleNet') by running the
    reference's own constructors (net/*.py) with rng = RandomState(23455).  Theano is replaced by an inert stand-in:
    the constructors' NUMERIC work - layer wiring and numbering, dimension arithmetic, weight initialisation and the
    order of the random draws (net/layer.py:60-118) - runs for real, the symbolic graph calls return placeholders.
    Returns a JSON-able description: per layer class, layerNum, inputDim, outputDim and per parameter name, shape,
    sha1 of the float32 bytes, sum and first values; plus the order of ``net.params``."""
    import hashlib
    from unittest import mock
    import numpy as np
    theano = mock.MagicMock()
    theano.shared = lambda value=None, name=None, borrow=False, **kw: _Shared(value, name, borrow)
    theano.config.floatX = 'float32'
    fake = {'theano': theano, 'theano.tensor': theano.tensor, 'theano.tensor.nnet': theano.tensor.nnet,
            'theano.tensor.signal': theano.tensor.signal, 'theano.tensor.signal.pool': theano.tensor.signal.pool,
            'theano.ifelse': theano.ifelse, 'theano.sandbox': theano.sandbox,
            'theano.sandbox.rng_mrg': theano.sandbox.rng_mrg, 'theano.sandbox.neighbours': theano.sandbox.neighbours}
    with _reference_net_modules(fake) as mods:
        mod = mods['net.' + kind.lower()]
        params = getattr(mod, kind + 'Params')(**cfg)
        net = getattr(mod, kind)(np.random.RandomState(23455), cfgParams=params)

        def pdesc(p):
            v = np.ascontiguousarray(p.get_value(), dtype=np.float32)
            return dict(name=p.name, shape=list(v.shape), sha1=hashlib.sha1(v.tobytes()).hexdigest(),
                        sum=float(v.astype(np.float64).sum()), head=[float(x) for x in v.ravel()[:4]])
        layers = []
        for l in net.layers:
            layers.append(dict(cls=type(l).__name__, layerNum=int(l.layerNum),
                               inputDim=[int(x) for x in l.cfgParams.inputDim],
                               outputDim=[int(x) for x in l.cfgParams.outputDim],
                               params=[pdesc(p) for p in l.params],
                               params_nontrained=[pdesc(p) for p in getattr(l, 'params_nontrained', [])],
                               weights=[p.name for p in l.weights]))
        return dict(kind=kind, cfg=cfg, layers=layers, net_params=[p.name for p in net.params],
                    outputDim=[int(x) for x in params.outputDim])


def run_reference_net(kind, cfg, inputs, deterministic=True, y=None, learning_rate=None, weightreg_factor=0.0,
                      bn_state=None):
    """EVALUATES the reference's own network code on concrete inputs with oracle/eager_theano.py standing in for
    Theano (read that module's header for what this does and does not prove).

    inputs: list of float arrays (one per network input, NCHW); deterministic: the value every ``flag_on`` switch
    (BatchNormLayer, DropoutLayer) is created with - False = training graph (batch statistics, dropout masks).
    bn_state: optional {param name: array} written into the shared variables before the forward pass is built is not
    possible in an eager run, so non-default running statistics are given here and installed at construction.
    With ``y`` (targets) the reference's cost and gradients are evaluated too (trainer/poseregnettrainer.py:70-111,
    ``setupFunctions``), and with ``learning_rate`` one ADAM step (trainer/optimizer.py:58-90).
    Returns dict(out, layer_out {layerNum: array}, masks [dropout masks], bn_updates [(name, new value)],
    cost, grads {param name: array}, new_params {param name: array})."""
    import numpy as np
    from oracle import eager_theano as E
    fake = E.build_modules()
    theano, T = fake['theano'], fake['theano.tensor']
    clones = []
    flag = 0.0 if deterministic else 1.0
    state = dict(bn_state or {})

    def shared(value=None, name=None, borrow=False, **kw):
        if name == 'flag_on':
            value = np.float32(flag)
        if name in state:
            value = np.asarray(state[name], dtype=np.asarray(value).dtype)
        return E.Shared(value, name, borrow)

    def clone(x, share_inputs=True, **kw):
        c = E.ET(x.t, getattr(x, 'name', None))
        clones.append(c)
        return c
    theano.shared, theano.clone = shared, clone
    E.FEED.clear()
    names = ['x'] if len(inputs) == 1 else ['x%d' % i for i in range(len(inputs))]
    for n, a in zip(names, inputs):
        E.FEED[n] = np.asarray(a, np.float64)
    res = {}
    with _reference_net_modules(fake) as mods:
        mod = mods['net.' + kind.lower()]
        params = getattr(mod, kind + 'Params')(**cfg)
        net = getattr(mod, kind)(np.random.RandomState(23455), cfgParams=params)
        res['out'] = net.output.eval().copy()
        res['layer_out'] = {int(l.layerNum): l.output.eval().copy() for l in net.layers}
        res['masks'] = [l.mask.eval().copy() for l in net.layers if type(l).__name__ == 'DropoutLayer']
        res['bn_updates'] = [(c.name, c.default_update.eval().copy()) for c in clones if c.default_update is not None]
        if y is not None:
            nj = 14                                   # the evaluation-only expressions (:113-125) need some values
            E.FEED['y'] = [np.asarray(y, np.float64), np.zeros((params.batch_size, nj, 3))]
            E.FEED['pca'] = np.zeros((int(np.prod(np.shape(y)[1:])), nj * 3))
            E.FEED['mean'] = np.zeros(nj * 3)
            E.FEED['learning_rate'] = np.float64(learning_rate if learning_rate is not None else 0.0)
            E.FEED['momentum'] = np.float64(0.0)
            E.FEED[None] = np.float64(0.0)
            setup = reference_function('trainer/poseregnettrainer.py', 'setupFunctions', {'theano': theano, 'T': T})

            class _Cfg(object):
                pass
            tr = types.SimpleNamespace(poseNet=net, cfgParams=_Cfg())
            tr.cfgParams.batch_size = params.batch_size
            tr.cfgParams.weightreg_factor = weightreg_factor
            setup(tr)
            res['cost'] = float(tr.cost.eval())
            res['param_order'] = [p.name for p in tr.params]
            res['grads'] = {p.name: g.eval().copy() for p, g in zip(tr.params, tr.grads)}
            if learning_rate is not None:
                adam = reference_function('trainer/optimizer.py', 'ADAM', {'theano': theano, 'T': T, 'numpy': np})
                opt = types.SimpleNamespace(params=tr.params, grads=tr.grads, updates=[], shared=[])
                updates = adam(opt, tr.learning_rate)
                by_id = {id(p): p.name for p in tr.params}
                res['new_params'] = {by_id[id(s)]: v.eval().copy() for s, v in updates if id(s) in by_id}
                res['adam_t_next'] = float(updates[-1][1].eval())
    return res


def reference_checkpoint_roundtrip(kind, cfg, save_path=None, load_path=None, seed=23455):
    """Checkpoint compatibility against the reference's own NetBase.save / NetBase.load (net/netbase.py:405-477), run
    on a reference network built with the inert theano stand-in.  ``save_path``: the reference net (weights from
    RandomState(seed)) writes its pickle there.  ``load_path``: a pickle written by the product is loaded by the
    reference code; returns {param name: value} of the reference net after loading (and its own description string)."""
    from unittest import mock
    import numpy as np
    theano = mock.MagicMock()
    theano.shared = lambda value=None, name=None, borrow=False, **kw: _Shared(value, name, borrow)
    theano.config.floatX = 'float32'
    fake = {'theano': theano, 'theano.tensor': theano.tensor, 'theano.tensor.nnet': theano.tensor.nnet,
            'theano.tensor.signal': theano.tensor.signal, 'theano.tensor.signal.pool': theano.tensor.signal.pool,
            'theano.ifelse': theano.ifelse, 'theano.sandbox': theano.sandbox,
            'theano.sandbox.rng_mrg': theano.sandbox.rng_mrg, 'theano.sandbox.neighbours': theano.sandbox.neighbours}
    with _reference_net_modules(fake) as mods:
        mod = mods['net.' + kind.lower()]
        params = getattr(mod, kind + 'Params')(**cfg)
        net = getattr(mod, kind)(np.random.RandomState(seed), cfgParams=params)
        if save_path is not None:
            net.save(save_path)
        if load_path is not None:
            net.load(load_path)
        vals = {}
        for l in net.layers:
            for p in list(l.params) + list(getattr(l, 'params_nontrained', [])):
                vals[p.name] = np.array(p.get_value(), copy=True)
        return dict(values=vals, network=str(net))


def run_reference_adam(param_values, grads_per_step, learning_rates):
    """The reference's Optimizer.ADAM (trainer/optimizer.py:58-90) evaluated with oracle/eager_theano.py over several
    steps.  An eager graph is one-shot, so every step rebuilds the update expressions with the optimiser state of the
    previous step (t, first / second moments) installed into the shared variables ADAM() creates, in creation
    order.  Returns the parameter values after every step (list of lists, float64)."""
    import numpy as np
    from oracle import eager_theano as E
    fake = E.build_modules()
    theano, T = fake['theano'], fake['theano.tensor']
    saved = {k: sys.modules.get(k) for k in fake}
    sys.modules.update(fake)
    if not hasattr(np, 'cast'):
        class _Cast(dict):
            def __missing__(self, k):
                return lambda v: np.asarray(v, dtype=k)[()]
        np.cast = _Cast()
        added_cast = True
    else:
        added_cast = False
    try:
        adam = reference_function('trainer/optimizer.py', 'ADAM', {'theano': theano, 'T': T, 'numpy': np})
        values = [np.asarray(p, np.float64).copy() for p in param_values]
        state = None                                       # (t, [m0, v0, m1, v1, ...])
        history = []
        for grads, lr in zip(grads_per_step, learning_rates):
            created = []

            def shared(value=None, name=None, borrow=False, **kw):
                k = len(created)
                if state is not None:
                    value = np.asarray(state[0] if k == 0 else state[1][k - 1], dtype=np.asarray(value).dtype)
                s = E.Shared(value, name, borrow)
                created.append(s)
                return s
            theano.shared = shared
            params = [E.Shared(v.astype(np.float32), 'p%d' % i) for i, v in enumerate(values)]
            for p, v in zip(params, values):                   # keep float64 state between the rebuilt steps
                p.set_value(v)
            created.clear()
            opt = types.SimpleNamespace(params=params, grads=[E.ET(np.asarray(g, np.float64)) for g in grads],
                                        updates=[], shared=[])
            updates = adam(opt, E.ET(np.float64(lr)))
            new = {id(s): v.eval().copy() for s, v in updates}
            values = [new[id(p)] for p in params]
            state = (new[id(created[0])], [new[id(s)] for s in created[1:]])
            history.append([v.copy() for v in values])
        return history
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if added_cast:
            del np.cast


def reference_netbase_behaviour(kind, cfg):
    """NetBase's bookkeeping methods on a reference network (inert theano): the deterministic switch, hasDropout, the
    params / weights views with filters, weightVals round trip.  Returns a plain dict of observations."""
    from unittest import mock
    import numpy as np
    theano = mock.MagicMock()
    theano.shared = lambda value=None, name=None, borrow=False, **kw: _Shared(value, name, borrow)
    theano.config.floatX = 'float32'
    fake = {'theano': theano, 'theano.tensor': theano.tensor, 'theano.tensor.nnet': theano.tensor.nnet,
            'theano.tensor.signal': theano.tensor.signal, 'theano.tensor.signal.pool': theano.tensor.signal.pool,
            'theano.ifelse': theano.ifelse, 'theano.sandbox': theano.sandbox,
            'theano.sandbox.rng_mrg': theano.sandbox.rng_mrg, 'theano.sandbox.neighbours': theano.sandbox.neighbours}
    with _reference_net_modules(fake) as mods:
        mod = mods['net.' + kind.lower()]
        net = getattr(mod, kind)(np.random.RandomState(23455), cfgParams=getattr(mod, kind + 'Params')(**cfg))
        return observe_netbase(net)


def observe_netbase(net):
    """the observations of ``reference_netbase_behaviour``, for a reference OR a product network"""
    import numpy as np
    obs = {'hasDropout': bool(net.hasDropout()), 'det0': bool(net.isDeterministic())}
    net.setDeterministic()
    obs['det1'] = bool(net.isDeterministic())
    net.unsetDeterministic()
    obs['det2'] = bool(net.isDeterministic())
    obs['params'] = [p.name for p in net.params]
    obs['weights'] = [p.name for p in net.weights]
    obs['all_params'] = [p.name for p in net.all_params]
    first = list(net.params)[0]            # py2's dict.values() was a list (netbase.py:165)
    net.params_filter = [first]
    obs['params_filtered'] = [p.name for p in net.params]
    net.params_filter = []
    # (weightVals is not observed: the reference's recGetWeightVals needs Python 2's list-returning dict.values())
    try:
        net.params_filter = [object()]
        obs['bad_filter'] = 'accepted'
    except Exception as e:
        obs['bad_filter'] = type(e).__name__
    return obs
