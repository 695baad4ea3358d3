"""Activation markers (reference: src/util/theano_helpers.py:61-69).

The reference's ``ReLU`` is a Python function building ``T.maximum(x, 0)``; layer parameter
classes store it in ``activation`` and derive ``activation_str`` from ``__name__``.  Here it is the
same kind of object - a plain function whose ``__name__`` is 'ReLU' - but applied to a symbolic
handle it only records the op; the arithmetic is the fused ReLU inside the CUDA kernels.
"""


def ReLU(x):
    from net.sym import Sym
    if isinstance(x, Sym):
        return Sym('relu', (x,), shape=x.shape)
    raise TypeError("ReLU is a graph marker here; arithmetic runs in libdpp_b200.so")


def sigmoid(x):
    raise NotImplementedError("sigmoid is not on the DeepPrior++ hot path")


def tanh(x):
    raise NotImplementedError("tanh is not on the DeepPrior++ hot path")
