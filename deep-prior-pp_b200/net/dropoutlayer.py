"""DropoutLayer (reference: src/net/dropoutlayer.py:38-138): train: mask*x with mask ~
Bernoulli(1-p); deterministic: (1-p)*x (non-inverted dropout), p = 0.3.
The reference draws masks from Theano's MRG31k3p stream, which cannot be reproduced without
Theano; here masks come from a per-layer torch Philox generator seeded with the same
``rng.randint(999999)`` draw (so the construction-time rng stream stays aligned), or are injected
by the caller for parity tests (``engine.set_dropout_masks``).  Folded into dpp_fc_fwd's epilogue."""
import numpy
from net.layerparams import LayerParams
from net.layer import Layer
from net.sym import Sym
from net.batchnormlayer import _Flag


class DropoutLayerParams(LayerParams):
    def __init__(self, inputDim=None, outputDim=None, p=0.3):
        super(DropoutLayerParams, self).__init__(inputDim, outputDim)
        self._p = p

    @property
    def p(self):
        return self._p

    @p.setter
    def p(self, value):
        self._p = value


class DropoutLayer(Layer):
    def __init__(self, rng, inputVar, cfgParams, copyLayer=None, layerNum=None):
        super(DropoutLayer, self).__init__(rng)
        self.inputVar = inputVar
        self.cfgParams = cfgParams
        self.layerNum = layerNum
        assert 0. < cfgParams.p < 1.
        self.prob_drop = cfgParams.p
        self.prob_keep = 1.0 - cfgParams.p
        self.flag_on = _Flag(1.0)
        self.mask_seed = int(rng.randint(999999))      # dropoutlayer.py:96
        self.output = Sym('layer', (inputVar,), layer=self, shape=cfgParams.outputDim,
                          name='output_layer_{}'.format(layerNum))
        self.output_pre_act = self.output
        self.params = []
        self.weights = []

    def unsetDeterministic(self):
        self.flag_on.set_value(1.0)

    def setDeterministic(self):
        self.flag_on.set_value(0.0)

    def isDeterministic(self):
        return bool(numpy.allclose(self.flag_on.get_value(), 0.0))

    def __str__(self):
        return "inputDim {}, outputDim {}, p {}".format(self.cfgParams.inputDim, self.cfgParams.outputDim,
                                                        self.cfgParams.p)
