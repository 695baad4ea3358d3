"""ConvLayer (reference: src/net/convlayer.py:39-266): conv2d(border 'half'|'valid', stride) + bias.
Arithmetic: dpp_conv2d_fwd/dgrad/wgrad in libdpp_b200.so (include/dpp_b200.h)."""
import numpy
from net.layerparams import LayerParams, tracked
from net.layer import Layer
from net.sym import Sym, shared


def _border(mode):
    return 'half' if mode == 'same' else mode          # the entry code says 'same', theano's conv2d 'half'


class ConvLayerParams(LayerParams):
    """convlayer.py:39-170.  Reassigning a tracked attribute recomputes filter / image / output shapes."""
    stride = tracked('stride')
    border_mode = tracked('border_mode', convert=_border)
    nFilters = tracked('nFilters')
    filterDim = tracked('filterDim')
    activation = tracked('activation', refresh=False)
    hasBias = tracked('hasbias', refresh=False)
    filter_shape = property(lambda self: self._filter_shape)
    image_shape = property(lambda self: self._image_shape)

    def __init__(self, inputDim=None, nFilters=None, filterDim=None, activation=None, hasBias=True,
                 filter_shape=None, image_shape=None, outputDim=None, stride=(1, 1), border_mode='valid',
                 init_method=None):
        super(ConvLayerParams, self).__init__(inputDim, outputDim)
        self._nFilters, self._filterDim = nFilters, filterDim
        self._filter_shape, self._image_shape = filter_shape, image_shape
        self._activation, self._hasbias = activation, hasBias
        self._stride, self._border_mode = stride, _border(border_mode)
        self._init_method = init_method
        self.update()

    def _conv_dims(self):
        """convlayer.py:133-163: [N, nFilters, ceil(H' / stride), ceil(W' / stride)] with H' by border mode; also
        refreshes filter_shape / image_shape"""
        n, cin, h, w = self._inputDim
        kh, kw = self._filterDim
        self._filter_shape = (self._nFilters, cin, kh, kw)
        self._image_shape = self._inputDim
        grow = {'valid': (1 - kh, 1 - kw), 'full': (kh - 1, kw - 1), 'half': (0, 0)}
        if self._border_mode not in grow:
            raise ValueError("Unknown border mode")
        dh, dw = grow[self._border_mode]
        return [n, self._nFilters, int(numpy.ceil((h + dh) / float(self._stride[0]))),
                int(numpy.ceil((w + dw) / float(self._stride[1])))]

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
