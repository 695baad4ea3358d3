"""BatchNormLayer (reference: src/net/batchnormlayer.py:40-222): per-channel batch statistics in
training, stored mean / INV_STD (EMA alpha 0.1) in deterministic mode, eps 1e-4, trainable
[beta, gamma].  Arithmetic is fused into the neighbouring conv kernels (statistics in the
producer's epilogue, normalise+ReLU in the consumer's prologue) - see dpp_bn_ref."""
import numpy
from net.layerparams import LayerParams
from net.layer import Layer
from net.sym import Sym, shared


class BatchNormLayerParams(LayerParams):
    def __init__(self, inputDim=None, outputDim=None, epsilon=1e-4, alpha=0.1, mode='low_mem',
                 learn_beta=True, learn_gamma=True):
        super(BatchNormLayerParams, self).__init__(inputDim, outputDim)
        self._learn_beta = learn_beta
        self._learn_gamma = learn_gamma
        self._epsilon = epsilon
        self._alpha = alpha
        self._mode = mode
        self._outputDim = self._inputDim

    epsilon = property(lambda self: self._epsilon, lambda self, v: setattr(self, '_epsilon', v))
    alpha = property(lambda self: self._alpha, lambda self, v: setattr(self, '_alpha', v))
    mode = property(lambda self: self._mode, lambda self, v: setattr(self, '_mode', v))


class _Flag(object):
    def __init__(self, v):
        self.v = numpy.float32(v)

    def set_value(self, v):
        self.v = numpy.float32(v)

    def get_value(self):
        return self.v


class BatchNormLayer(Layer):
    def __init__(self, rng, inputVar, cfgParams, copyLayer=None, layerNum=None):
        super(BatchNormLayer, self).__init__(rng)
        self.cfgParams = cfgParams
        self.layerNum = layerNum
        self.inputVar = inputVar
        inputDim = cfgParams.inputDim
        self.flag_on = _Flag(1.0)
        axes = (0,) + tuple(range(2, len(inputDim)))
        shape = tuple([size for axis, size in enumerate(inputDim) if axis not in axes])
        if copyLayer is not None:
            self.beta = copyLayer.beta
            self.gamma = copyLayer.gamma
        else:
            self.beta = shared(numpy.zeros(shape, dtype='float32'), name='beta{}'.format(layerNum))
            self.gamma = shared(numpy.ones(shape, dtype='float32'), name='gamma{}'.format(layerNum))
        self.mean = shared(numpy.zeros(shape, dtype='float32'), name='mean{}'.format(layerNum))
        self.inv_std = shared(numpy.ones(shape, dtype='float32'), name='inv_std{}'.format(layerNum))
        if copyLayer is not None:
            self.mean.set_value(copyLayer.mean.get_value())
            self.inv_std.set_value(copyLayer.inv_std.get_value())
        self.weights = []
        self.params = []
        if cfgParams._learn_beta is True:
            self.params.append(self.beta)
        if cfgParams._learn_gamma is True:
            self.params.append(self.gamma)
        self.params_nontrained = [self.mean, self.inv_std]
        self.output = Sym('layer', (inputVar,), layer=self, shape=cfgParams.outputDim,
                          name='output_layer_{}'.format(layerNum))
        self.output_pre_act = self.output

    def unsetDeterministic(self):
        self.flag_on.set_value(1.0)

    def setDeterministic(self):
        self.flag_on.set_value(0.0)

    def isDeterministic(self):
        return bool(numpy.allclose(self.flag_on.get_value(), 0.0))

    def __str__(self):
        return "epsilon {}, alpha {}".format(self.cfgParams.epsilon, self.cfgParams.alpha)
