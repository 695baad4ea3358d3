"""Layer base: parameter lists and initial values (reference: src/net/layer.py:35-124).
The random draws are made from the caller's ``rng`` in exactly the reference's order and
distributions so a net built with the same seed has the same initial weights."""
import numpy

floatX = 'float32'


class Layer(object):
    def __init__(self, rng):
        self.weights = []
        self.params = []
        self.params_nontrained = []
        self.rng = rng

    def orthogonalize(self, init_vals):
        """layer.py:49-55: replace the filters by an orthonormal set - the leading left singular vectors of the
        (fan_in x n_filters) matrix of the drawn values."""
        flat = numpy.reshape(init_vals, (init_vals.shape[0], -1))
        left = numpy.linalg.svd(flat.T)[0]
        basis = left.T[0:init_vals.shape[0]].T
        return numpy.reshape(basis.swapaxes(0, 1), init_vals.shape)

    def getOptimalInitMethod(self, act_str):
        # layer.py:58-70
        if act_str == 'ReLU':
            return 'He'
        elif act_str == 'sigmoid':
            return 'sigmoid'
        elif act_str in 'tanh':
            return 'tanh'
        elif act_str is None or str(act_str) == 'None':
            return None
        raise NotImplementedError("Unknown activation function: {}".format(act_str))

    def getInitVals(self, shape, mode, act_fn=None, method=None, orthogonal=False):
        # layer.py:72-124
        if act_fn is None and method is None:
            raise UserWarning("act_fn and method not defined! At least one must be specified.")
        if act_fn is not None and method is None:
            method = self.getOptimalInitMethod(act_fn)
        if method == 'He':
            if mode == 'conv':
                W_bound = numpy.sqrt(2. / numpy.prod(shape[1:]))
                init_vals = numpy.asarray(self.rng.normal(loc=0.0, scale=W_bound, size=shape), dtype=floatX)
            elif mode == 'fc':
                init_vals = numpy.asarray(self.rng.normal(loc=0.0, scale=0.01, size=shape), dtype=floatX)
            else:
                raise NotImplementedError()
        elif method == 'Xavier':
            if mode == 'conv':
                W_bound = numpy.sqrt(3. / numpy.prod(shape[1:]))
            elif mode == 'fc':
                W_bound = numpy.sqrt(1. / shape[0])
            else:
                raise NotImplementedError()
            init_vals = numpy.asarray(self.rng.uniform(low=-W_bound, high=W_bound, size=shape), dtype=floatX)
        elif method == 'sigmoid':
            if mode == 'conv':
                W_bound = 4. * numpy.sqrt(6. / (numpy.prod(shape[1:]) + (shape[0] * numpy.prod(shape[2:]))))
                init_vals = numpy.asarray(self.rng.uniform(low=-W_bound, high=W_bound, size=shape), dtype=floatX)
            elif mode == 'fc':
                b = numpy.sqrt(6. / numpy.sum(shape))
                init_vals = 4. * numpy.asarray(self.rng.uniform(low=-b, high=b, size=shape), dtype=floatX)
            else:
                raise NotImplementedError()
        elif method == 'tanh' or method is None:
            if mode == 'conv':
                W_bound = 1. / (numpy.prod(shape[1:]) + (shape[0] * numpy.prod(shape[2:])))
                init_vals = numpy.asarray(self.rng.uniform(low=-W_bound, high=W_bound, size=shape), dtype=floatX)
            elif mode == 'fc':
                b = numpy.sqrt(6. / numpy.sum(shape))
                init_vals = numpy.asarray(self.rng.uniform(low=-b, high=b, size=shape), dtype=floatX)
            else:
                raise NotImplementedError()
        else:
            raise NotImplementedError("Unknown method!")
        if orthogonal:
            init_vals = self.orthogonalize(init_vals)
        return init_vals
