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
        """layer.py:72-124: initial weights for a 'conv' (O, I, kh, kw) or 'fc' (n_in, n_out) tensor.  One draw from
        ``self.rng`` per tensor - normal for 'He', uniform otherwise - so a net built from the same seed reproduces the
        reference's weights bit for bit (tests/test_reference_pins.py)."""
        if act_fn is None and method is None:
            raise UserWarning("act_fn and method not defined! At least one must be specified.")
        if method is None and act_fn is not None:
            method = self.getOptimalInitMethod(act_fn)
        if mode not in ('conv', 'fc'):
            raise NotImplementedError()
        conv = (mode == 'conv')
        fan_in = numpy.prod(shape[1:]) if conv else None
        fan_both = (fan_in + shape[0] * numpy.prod(shape[2:])) if conv else None      # fan-in + fan-out of a filter bank
        glorot_fc = None if conv else numpy.sqrt(6. / numpy.sum(shape))
        gain = 1.
        if method == 'He':
            sigma = numpy.sqrt(2. / fan_in) if conv else 0.01
            draw = self.rng.normal(loc=0.0, scale=sigma, size=shape)
        else:
            if method == 'Xavier':
                bound = numpy.sqrt(3. / fan_in) if conv else numpy.sqrt(1. / shape[0])
            elif method == 'sigmoid':
                bound = 4. * numpy.sqrt(6. / fan_both) if conv else glorot_fc
                gain = 1. if conv else 4.             # fully connected: the drawn values are scaled, not the bound
            elif method == 'tanh' or method is None:
                bound = 1. / fan_both if conv else glorot_fc
            else:
                raise NotImplementedError("Unknown method!")
            draw = self.rng.uniform(low=-bound, high=bound, size=shape)
        init_vals = numpy.asarray(draw, dtype=floatX)
        if gain != 1.:
            init_vals = gain * init_vals
        return self.orthogonalize(init_vals) if orthogonal else init_vals
