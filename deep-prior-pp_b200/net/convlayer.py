"""ConvLayer (reference: src/net/convlayer.py:39-266): conv2d(border 'half'|'valid', stride) + bias.
Arithmetic: dpp_conv2d_fwd/dgrad/wgrad in libdpp_b200.so (include/dpp_b200.h)."""
import numpy
from net.layerparams import LayerParams
from net.layer import Layer
from net.sym import Sym, shared


class ConvLayerParams(LayerParams):
    def __init__(self, inputDim=None, nFilters=None, filterDim=None, activation=None, hasBias=True,
                 filter_shape=None, image_shape=None, outputDim=None, stride=(1, 1), border_mode='valid',
                 init_method=None):
        super(ConvLayerParams, self).__init__(inputDim, outputDim)
        self._nFilters = nFilters
        self._filterDim = filterDim
        self._filter_shape = filter_shape
        self._image_shape = image_shape
        self._activation = activation
        self._hasbias = hasBias
        self._stride = stride
        self._border_mode = 'half' if border_mode == 'same' else border_mode
        self._init_method = init_method
        self.update()

    filter_shape = property(lambda self: self._filter_shape)
    image_shape = property(lambda self: self._image_shape)

    @property
    def stride(self):
        return self._stride

    @stride.setter
    def stride(self, value):
        self._stride = value
        self.update()

    @property
    def border_mode(self):
        return self._border_mode

    @border_mode.setter
    def border_mode(self, value):
        self._border_mode = 'half' if value == 'same' else value
        self.update()

    @property
    def nFilters(self):
        return self._nFilters

    @nFilters.setter
    def nFilters(self, value):
        self._nFilters = value
        self.update()

    @property
    def filterDim(self):
        return self._filterDim

    @filterDim.setter
    def filterDim(self, value):
        self._filterDim = value
        self.update()

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

    def _conv_dims(self):
        # convlayer.py:133-163
        self._filter_shape = (self._nFilters, self._inputDim[1], self._filterDim[0], self._filterDim[1])
        self._image_shape = self._inputDim
        if self._border_mode == 'valid':
            o = (self._inputDim[0], self._nFilters, self._inputDim[2] - self._filterDim[0] + 1,
                 self._inputDim[3] - self._filterDim[1] + 1)
        elif self._border_mode == 'full':
            o = (self._inputDim[0], self._nFilters, self._inputDim[2] + self._filterDim[0] - 1,
                 self._inputDim[3] + self._filterDim[1] - 1)
        elif self._border_mode == 'half':
            o = (self._inputDim[0], self._nFilters, self._inputDim[2], self._inputDim[3])
        else:
            raise ValueError("Unknown border mode")
        o = list(o)
        o[2] = int(numpy.ceil(o[2] / float(self._stride[0])))
        o[3] = int(numpy.ceil(o[3] / float(self._stride[1])))
        return o

    def update(self):
        self._outputDim = tuple(self._conv_dims())

    def getMemoryRequirement(self):
        return (numpy.prod(self.filter_shape) + self.filter_shape[0]) * 4


class ConvLayer(Layer):
    def __init__(self, rng, inputVar, cfgParams, copyLayer=None, layerNum=None):
        super(ConvLayer, self).__init__(rng)
        assert isinstance(cfgParams, ConvLayerParams)
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
               "hasBias {}".format(self.cfgParams.inputDim, self.cfgParams.outputDim, self.cfgParams.filterDim,
                                   self.cfgParams.nFilters, self.cfgParams.activation_str, self.cfgParams.stride,
                                   self.cfgParams.border_mode, self.cfgParams.hasBias)
