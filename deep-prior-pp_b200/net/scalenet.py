"""ScaleNet - the multi-scale CoM-refinement network of the inference cascade (reference:
src/net/scalenet.py:49-193; used by util/handdetector.py:634-676 ``refineCoM`` and
util/realtimehandposepipeline.py:117-131).  Three towers of three 'valid' ConvPool layers (8 filters each) on the
128x128 crop and its 64x64 / 32x32 centre crops, flattened and concatenated (968 + 968 + 512 = 2448 features),
then FC1024 - dropout - FC1024 - dropout - FC(numJoints*nDims).  Forward only on the hot path;
the layer list, layer numbering and parameter order are the reference's, so its pickles load."""
from net.convpoollayer import ConvPoolLayer, ConvPoolLayerParams
from net.hiddenlayer import HiddenLayer, HiddenLayerParams
from net.dropoutlayer import DropoutLayer, DropoutLayerParams
from net.netbase import NetBase, NetBaseParams
from net.sym import tensor4, concatenate
from util.theano_helpers import ReLU


class ScaleNetParams(NetBaseParams):
    def __init__(self, type=0, nChan=1, wIn=128, hIn=128, batchSize=128, numJoints=16, nDims=3, resizeFactor=2,
                 shared_conv=False):
        super(ScaleNetParams, self).__init__()
        self.batch_size = batchSize
        self.numJoints = numJoints
        self.nDims = nDims
        self.shared_conv = shared_conv
        if type != 1:
            raise NotImplementedError("not implemented")
        self.numInputs = 3
        self.inpConv = 3          # ConvPool layers per tower
        f = resizeFactor
        self.inputDim = [(batchSize, nChan, hIn, wIn), (batchSize, nChan, hIn // f, wIn // f),
                         (batchSize, nChan, hIn // f ** 2, wIn // f ** 2)]
        # (filter edge, pool edge) of the three layers of each tower (scalenet.py:55-107)
        towers = [[(5, 4), (5, 2), (3, 1)], [(5, 2), (5, 2), (3, 1)], [(5, 2), (5, 1), (3, 1)]]
        for t, spec in enumerate(towers):
            dim = self.inputDim[t]
            for k, pool in spec:
                self.layers.append(ConvPoolLayerParams(inputDim=dim, nFilters=8, filterDim=(k, k), poolsize=(pool, pool),
                                                       activation=ReLU))
                dim = self.layers[-1].outputDim
        lout = 0
        for t in range(self.numInputs):
            od = self.layers[(t + 1) * self.inpConv - 1].outputDim
            lout += od[1] * od[2] * od[3]
        self.layers.append(HiddenLayerParams(inputDim=(batchSize, lout), outputDim=(batchSize, 1024), activation=ReLU))
        self.layers.append(DropoutLayerParams(inputDim=self.layers[-1].outputDim, outputDim=self.layers[-1].outputDim))
        self.layers.append(HiddenLayerParams(inputDim=self.layers[-1].outputDim, outputDim=(batchSize, 1024),
                                             activation=ReLU))
        self.layers.append(DropoutLayerParams(inputDim=self.layers[-1].outputDim, outputDim=self.layers[-1].outputDim))
        self.layers.append(HiddenLayerParams(inputDim=self.layers[-1].outputDim,
                                             outputDim=(batchSize, numJoints * nDims), activation=None))
        self.outputDim = self.layers[-1].outputDim


class ScaleNet(NetBase):
    def __init__(self, rng, inputVar=None, cfgParams=None, twin=None):
        if cfgParams is None:
            raise Exception("Cannot create a Net without config parameters (ie. cfgParams==None)")
        if inputVar is not None:
            raise Exception("Do not give inputVar, created inline")
        self._params_filter = []
        self._weights_filter = []
        self.rng = rng
        self.cfgParams = cfgParams
        self.inputVar = [tensor4('x{}'.format(i)) for i in range(cfgParams.numInputs)]
        for k, v in enumerate(self.inputVar):
            v.shape = tuple(cfgParams.inputDim[k])
        self.layers = []
        n_conv = cfgParams.numInputs * cfgParams.inpConv
        for i, layerParam in enumerate(cfgParams.layers):
            if i < n_conv and i % cfgParams.inpConv == 0:
                inp = self.inputVar[i // cfgParams.inpConv]          # a tower starts at its own input
            elif i == n_conv:
                # the towers' last outputs, flattened and concatenated, feed the first hidden layer
                inp = concatenate([self.layers[(t + 1) * cfgParams.inpConv - 1].output.flatten(2)
                                   for t in range(cfgParams.numInputs)], axis=1)
            else:
                inp = self.layers[-1].output
            cl = None if twin is None else twin.layers[i]
            if cl is None and cfgParams.shared_conv is True and cfgParams.inpConv - 1 < i < n_conv:
                cl = self.layers[i % cfgParams.inpConv]
            constructor = globals()[layerParam.__class__.__name__[:-6]]
            self.layers.append(constructor(rng, inputVar=inp, cfgParams=layerParam, copyLayer=cl, layerNum=i))
        self.output = self.layers[-1].output
        self.load(self.cfgParams.loadFile)
