"""ResNet (reference: src/net/resnet.py:45-414): 5x5 stem ConvPool, 4 stages of 5 pre-activation
bottleneck blocks (widths 64/128/256/256; type 3: 64/128/128/128), final BN+ReLU, FC tail by
``type``.  Layer order and numbering are the reference's (checkpoint keys depend on them):
stem = 0, projection block = 10 layers, identity block = 9 layers."""
import numpy
from net.netbase import NetBase, NetBaseParams
from net.sym import tensor4
from net.convpoollayer import ConvPoolLayer, ConvPoolLayerParams
from net.convlayer import ConvLayer, ConvLayerParams
from net.hiddenlayer import HiddenLayer, HiddenLayerParams
from net.dropoutlayer import DropoutLayer, DropoutLayerParams
from net.batchnormlayer import BatchNormLayer, BatchNormLayerParams
from net.nonlinearitylayer import NonlinearityLayer, NonlinearityLayerParams
from util.theano_helpers import ReLU


class ResNetParams(NetBaseParams):
    def __init__(self, type=0, nChan=1, wIn=128, hIn=128, batchSize=128, numJoints=16, nDims=3):
        super(ResNetParams, self).__init__()
        self.batch_size = batchSize
        self.numJoints = numJoints
        self.nDims = nDims
        self.numInputs = 1
        self.inputDim = (batchSize, nChan, hIn, wIn)
        self.type = type
        if type not in (0, 1, 2, 3, 4):
            raise NotImplementedError("not implemented")
        self.numOutputs = 1
        self.outputDim = (batchSize, numJoints * nDims)


class ResNet(NetBase):
    def __init__(self, rng, inputVar=None, cfgParams=None):
        self._params_filter = []
        self._weights_filter = []
        if cfgParams is None:
            raise Exception("Cannot create a Net without config parameters (ie. cfgParams==None)")
        if inputVar is None:
            inputVar = tensor4('x')
        elif isinstance(inputVar, str):
            raise NotImplementedError()
        self.inputVar = inputVar
        self.cfgParams = cfgParams
        self.rng = rng
        self.layers = []
        L = self.layers
        batchSize = cfgParams.batch_size
        t = cfgParams.type
        if t not in (0, 1, 2, 3, 4):
            raise NotImplementedError()
        depth = 47
        assert (depth - 2) % 9 == 0, 'depth should be 9n+2 (e.g., 164 or 1001)'
        n = (depth - 2) // 9                                        # resnet.py:124 (py2 '/')
        nStages = [32, 64, 128, 128, 128] if t == 3 else [32, 64, 128, 256, 256]

        L.append(ConvPoolLayer(rng, self.inputVar,
                               ConvPoolLayerParams(inputDim=cfgParams.inputDim, nFilters=nStages[0], filterDim=(5, 5),
                                                   stride=(1, 1), poolsize=(2, 2), border_mode='same',
                                                   activation=None, init_method='He'), layerNum=len(L)))
        rout = L[-1].output
        for s in range(1, 5):
            rout = self.add_res_layers(rng, rout, L[-1].cfgParams.outputDim, nStages[s], n, 2)
        L.append(BatchNormLayer(rng, rout, BatchNormLayerParams(inputDim=L[-1].cfgParams.outputDim), layerNum=len(L)))
        L.append(NonlinearityLayer(rng, L[-1].output,
                                   NonlinearityLayerParams(inputDim=L[-1].cfgParams.outputDim, activation=ReLU),
                                   layerNum=len(L)))

        def fc(inp, in_dim, n_out, act):
            L.append(HiddenLayer(rng, inp, HiddenLayerParams(inputDim=in_dim, outputDim=(batchSize, n_out),
                                                             activation=act), layerNum=len(L)))

        def drop():
            L.append(DropoutLayer(rng, L[-1].output,
                                  DropoutLayerParams(inputDim=L[-1].cfgParams.outputDim,
                                                     outputDim=L[-1].cfgParams.outputDim), layerNum=len(L)))

        od = L[-1].cfgParams.outputDim
        fc(L[-1].output.flatten(2), (od[0], int(numpy.prod(od[1:]))), 1024, ReLU)
        if t in (2, 3, 4):
            drop()
        fc(L[-1].output, L[-1].cfgParams.outputDim, 1024, ReLU)
        if t in (2, 3, 4):
            drop()
        if t in (1, 4):
            fc(L[-1].output, L[-1].cfgParams.outputDim, 30, None)       # embedding bottleneck
        fc(L[-1].output, L[-1].cfgParams.outputDim, cfgParams.numJoints * cfgParams.nDims, None)
        self.output = L[-1].output
        self.load(self.cfgParams.loadFile)

    def add_res_layers(self, rng, inputVar, inputDim, outputFilters, count, stride):
        rout = res_block(self.layers, rng, inputVar, inputDim, outputFilters, stride)
        for i in range(1, count):
            rout = res_block(self.layers, rng, rout, self.layers[-1].cfgParams.outputDim, outputFilters, 1)
        return rout


def _bn_relu(layers, rng, inp, dim):
    layers.append(BatchNormLayer(rng, inp, BatchNormLayerParams(inputDim=dim), layerNum=len(layers)))
    layers.append(NonlinearityLayer(rng, layers[-1].output,
                                    NonlinearityLayerParams(inputDim=layers[-1].cfgParams.outputDim, activation=ReLU),
                                    layerNum=len(layers)))


def _conv(layers, rng, inp, dim, nf, k, stride=1):
    layers.append(ConvLayer(rng, inp, ConvLayerParams(inputDim=dim, nFilters=nf, filterDim=(k, k),
                                                      stride=(stride, stride), border_mode='same', activation=None,
                                                      init_method='He'), layerNum=len(layers)))


def res_block(layers, rng, inputVar, inputDim, outputFilters, stride, nBottleneckFilters=None):
    """resnet.py:349-414.  Identity block when the channel count already matches (the stride
    argument is then ignored, as in the reference); otherwise projection block whose shortcut conv
    reads the first ReLU output (``layers[-8]`` at that point)."""
    if nBottleneckFilters is None:
        nBottleneckFilters = outputFilters // 4
    identity = (inputDim[1] == outputFilters)
    s = 1 if identity else stride
    _bn_relu(layers, rng, inputVar, inputDim)
    _conv(layers, rng, layers[-1].output, layers[-1].cfgParams.outputDim, nBottleneckFilters, 1, s)
    _bn_relu(layers, rng, layers[-1].output, layers[-1].cfgParams.outputDim)
    _conv(layers, rng, layers[-1].output, layers[-1].cfgParams.outputDim, nBottleneckFilters, 3)
    _bn_relu(layers, rng, layers[-1].output, layers[-1].cfgParams.outputDim)
    _conv(layers, rng, layers[-1].output, layers[-1].cfgParams.outputDim, outputFilters, 1)
    if identity:
        return inputVar + layers[-1].output
    _conv(layers, rng, layers[-8].output, layers[-8].cfgParams.outputDim, outputFilters, 1, stride)
    return layers[-2].output + layers[-1].output
