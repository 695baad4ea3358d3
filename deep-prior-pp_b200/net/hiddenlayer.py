"""HiddenLayer (reference: src/net/hiddenlayer.py:40-169): act(dot(x, W) + b), W (n_in, n_out).
Arithmetic: dpp_fc_fwd / dpp_fc_bwd."""
import numpy
from net.layerparams import LayerParams
from net.layer import Layer
from net.sym import Sym, shared


class HiddenLayerParams(LayerParams):
    def __init__(self, inputDim=None, outputDim=None, activation=None, hasBias=True, init_method=None):
        super(HiddenLayerParams, self).__init__(inputDim, outputDim)
        self._activation = activation
        self._hasbias = hasBias
        self._init_method = init_method

    @property
    def activation(self):
        return self._activation

    @activation.setter
    def activation(self, value):
        self._activation = value

    @property
    def hasBias(self):
        return self._hasbias

    @hasBias.setter
    def hasBias(self, value):
        self._hasbias = value

    def getMemoryRequirement(self):
        return ((self.inputDim[1] * self.outputDim[1]) + self.outputDim[1]) * 4


class HiddenLayer(Layer):
    def __init__(self, rng, inputVar, cfgParams, copyLayer=None, layerNum=None):
        super(HiddenLayer, self).__init__(rng)
        assert isinstance(cfgParams, HiddenLayerParams)
        if not cfgParams.hasBias:
            raise NotImplementedError("bias-free HiddenLayer is unused on the hot path")
        self.inputVar = inputVar
        self.cfgParams = cfgParams
        self.layerNum = layerNum
        n_in = int(cfgParams.inputDim[1])
        n_out = int(cfgParams.outputDim[1])
        if copyLayer is None:
            wInitVals = self.getInitVals((n_in, n_out), 'fc', act_fn=cfgParams.activation_str,
                                         method=cfgParams._init_method)
            self.W = shared(wInitVals, name='W{}'.format(layerNum), kind='fcW')
            self.b = shared(numpy.zeros((n_out,), dtype='float32'), name='b{}'.format(layerNum))
        else:
            self.W = copyLayer.W
            self.b = copyLayer.b
        self.output = Sym('layer', (inputVar,), layer=self, shape=(cfgParams.outputDim[0], n_out),
                          name='output_layer_{}'.format(layerNum))
        self.output_pre_act = self.output
        self.params = [self.W, self.b]
        self.weights = [self.W]

    def __str__(self):
        return "inputDim {}, outputDim {}, activiation {}, hasBias {}".format(
            self.cfgParams.inputDim, self.cfgParams.outputDim, self.cfgParams.activation_str, self.cfgParams.hasBias)
