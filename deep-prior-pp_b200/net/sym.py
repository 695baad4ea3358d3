"""Symbolic handles and shared variables standing in for Theano's (reference: every
``theano.shared`` / ``T.tensor4`` / ``layer.output`` use in src/net/*.py and src/trainer/*.py).

``Sym`` records the wiring (which layer consumes which tensor, residual sums, flatten) so that
``dpp_b200.engine`` can lower a network object built with the reference's own constructor code
to calls into libdpp_b200.so.  ``SharedVariable`` keeps the ``get_value/set_value/name/auto_name``
surface the trainer and the entry scripts poke (trainer/nettrainer.py:915-917,
net/netbase.py:150-165, main_nyu_posereg_embedding.py:152-153); values live in host NumPy until an
engine binds the variable to its device arena, afterwards the device copy is the master.
"""
import numpy

_counter = [0]


class SharedVariable(object):
    def __init__(self, value, name=None, kind='plain'):
        self.name = name
        self.auto_name = 'auto_%d' % _counter[0]
        _counter[0] += 1
        self.kind = kind                 # 'plain' | 'convW' | 'fcW'  (device layout tag)
        self._host = numpy.array(value, dtype=numpy.float32, copy=True)
        self._binding = None             # (engine, slot) once on the device

    # -- theano.compile.SharedVariable surface --
    def get_value(self, borrow=False, return_internal_type=False):
        if self._binding is not None:
            eng, slot = self._binding
            return eng.download_param(slot)
        return self._host if borrow else self._host.copy()

    def set_value(self, value, borrow=False):
        value = numpy.asarray(value, dtype=numpy.float32)
        if value.shape != self._host.shape:
            # theano allows re-shaping a shared variable; the engine does not once bound
            if self._binding is not None:
                raise ValueError("shape change of a device-bound variable: %s -> %s" % (self._host.shape, value.shape))
            self._host = value.copy()
            return
        if self._binding is not None:
            eng, slot = self._binding
            eng.upload_param(slot, value)
        else:
            self._host = value.copy()

    @property
    def shape(self):
        return self._host.shape

    def dimshuffle(self, *a):
        return self

    def __repr__(self):
        return "<SharedVariable %s %s>" % (self.name, self._host.shape)


def shared(value, name=None, borrow=False, kind='plain'):
    return SharedVariable(value, name=name, kind=kind)


class Sym(object):
    """Node of the recorded graph.  op in {'input','layer','relu','add','flatten','reshape','concat'}."""

    def __init__(self, op, inputs=(), layer=None, shape=None, name=None):
        self.op = op
        self.inputs = tuple(inputs)
        self.layer = layer
        self.shape = shape
        self.name = name

    def flatten(self, ndim=1):
        assert ndim == 2
        shp = None
        if self.shape is not None:
            shp = (self.shape[0], int(numpy.prod(self.shape[1:])))
        return Sym('flatten', (self,), shape=shp)

    def reshape(self, shape, ndim=None):
        return Sym('reshape', (self,), shape=tuple(shape))

    def __add__(self, other):
        return Sym('add', (self, other), shape=self.shape)

    __radd__ = __add__

    def __repr__(self):
        return "<Sym %s %s %s>" % (self.op, self.name, self.shape)


def tensor4(name=None):
    return Sym('input', name=name)


def concatenate(tensor_list, axis=1):
    """T.concatenate of flattened (B, n_i) tensors along the feature axis (scalenet.py:174-178)."""
    assert axis == 1
    shp = None
    if all(t.shape is not None for t in tensor_list):
        shp = (tensor_list[0].shape[0], int(sum(t.shape[1] for t in tensor_list)))
    return Sym('concat', tuple(tensor_list), shape=shp)


def matrix(name=None):
    return Sym('input', name=name)
