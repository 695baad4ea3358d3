"""ConvPoolLayer (reference: src/net/convpoollayer.py:39-305): conv2d -> pool_2d(max,
ignore_border) -> +bias -> activation.  Arithmetic: dpp_convpool_fwd/bwd in libdpp_b200.so."""
import numpy
from net.convlayer import ConvLayerParams
from net.layerparams import tracked
from net.layer import Layer
from net.sym import Sym, shared


class ConvPoolLayerParams(ConvLayerParams):
    def __init__(self, inputDim=None, nFilters=None, filterDim=None, activation=None, poolsize=(1, 1), poolType=0,
                 filter_shape=None, image_shape=None, outputDim=None, stride=(1, 1), border_mode='valid',
                 hasBias=True, init_method=None):
        self._poolsize = poolsize
        self._poolType = poolType
        super(ConvPoolLayerParams, self).__init__(inputDim=inputDim, nFilters=nFilters, filterDim=filterDim,
                                                  activation=activation, hasBias=hasBias, filter_shape=filter_shape,
                                                  image_shape=image_shape, outputDim=outputDim, stride=stride,
                                                  border_mode=border_mode, init_method=init_method)

    poolsize = tracked('poolsize')
    poolType = property(lambda self: self._poolType)

    def update(self):
        """convpoollayer.py:110-146: the convolution's dimensions, floor-divided by the pooling window (theano's
        pool_2d with ignore_border=True); a 1x1 window means 'no pooling' (poolType -1)"""
        n, c, h, w = self._conv_dims()
        ph, pw = self._poolsize
        self._outputDim = (n, c, h // ph, w // pw)
        if (ph, pw) == (1, 1):
            self._poolType = -1


class ConvPoolLayer(Layer):
    def __init__(self, rng, inputVar, cfgParams, copyLayer=None, layerNum=None):
        super(ConvPoolLayer, self).__init__(rng)
        assert isinstance(cfgParams, ConvPoolLayerParams)
        if cfgParams.poolType not in (0, -1):
            raise NotImplementedError("only max pooling / no pooling are on the hot path")
        filter_shape = cfgParams.filter_shape
        assert cfgParams.image_shape[1] == filter_shape[1]
        self.cfgParams = cfgParams
        self.layerNum = layerNum
        self.inputVar = inputVar
        if copyLayer is not None:
            self.W = copyLayer.W
        else:
            wInitVals = self.getInitVals(filter_shape, 'conv', act_fn=cfgParams.activation_str, orthogonal=False,
                                         method=cfgParams._init_method)
            self.W = shared(wInitVals, name='convW{}'.format(layerNum), kind='convW')
        if cfgParams.hasBias is True:
            if copyLayer is not None:
                self.b = copyLayer.b
            else:
                self.b = shared(numpy.zeros((filter_shape[0],), dtype='float32'), name='convB{}'.format(layerNum))
        self.output = Sym('layer', (inputVar,), layer=self, shape=cfgParams.outputDim,
                          name='output_layer_{}'.format(layerNum))
        self.output_pre_act = self.output
        self.params = [self.W, self.b] if cfgParams.hasBias else [self.W]
        self.weights = [self.W]

    def __str__(self):
        return "inputDim {}, outputDim {}, filterDim {}, nFilters {}, activation {}, stride {}, border_mode {}, " \
               "hasBias {}, pool_type {}, pool_size {}".format(
                   self.cfgParams.inputDim, self.cfgParams.outputDim, self.cfgParams.filterDim,
                   self.cfgParams.nFilters, self.cfgParams.activation_str, self.cfgParams.stride,
                   self.cfgParams.border_mode, self.cfgParams.hasBias, self.cfgParams.poolType,
                   self.cfgParams.poolsize)
