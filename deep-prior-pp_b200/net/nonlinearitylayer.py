"""NonlinearityLayer (reference: src/net/nonlinearitylayer.py:38-133): output = activation(input).
ReLU only on the hot path; fused into the consumer kernel's prologue."""
from net.layerparams import LayerParams
from net.layer import Layer
from net.sym import Sym


class NonlinearityLayerParams(LayerParams):
    def __init__(self, inputDim=None, outputDim=None, activation=None):
        super(NonlinearityLayerParams, self).__init__(inputDim, outputDim)
        self._outputDim = self._inputDim
        self._activation = activation

    @property
    def activation(self):
        return self._activation

    @activation.setter
    def activation(self, value):
        self._activation = value

    def getMemoryRequirement(self):
        return 0                      # no weights (nonlinearitylayer.py:68-73)


class NonlinearityLayer(Layer):
    def __init__(self, rng, inputVar, cfgParams, copyLayer=None, layerNum=None):
        super(NonlinearityLayer, self).__init__(rng)
        self.cfgParams = cfgParams
        self.layerNum = layerNum
        self.inputVar = inputVar
        if cfgParams.activation_str != 'ReLU':
            raise NotImplementedError("only ReLU is on the hot path")
        self.output = Sym('layer', (inputVar,), layer=self, shape=cfgParams.outputDim,
                          name='output_layer_{}'.format(layerNum))
        self.output_pre_act = self.output
        self.params = []
        self.weights = []

    def __str__(self):
        return "inputDim {}, outputDim {}, activation {}".format(self.cfgParams.inputDim, self.cfgParams.outputDim,
                                                                  self.cfgParams.activation_str)
