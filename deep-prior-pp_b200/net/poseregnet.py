"""PoseRegNet (reference: src/net/poseregnet.py:44-165): 3x ConvPool (8 filters, valid) +
FC1024-drop-FC1024-drop-FC; type 11 adds the 30-D bottleneck.  What the three
main_*_posereg_embedding.py scripts build."""
from net.netbase import NetBase, NetBaseParams
from net.sym import tensor4
from net.convpoollayer import ConvPoolLayerParams
from net.hiddenlayer import HiddenLayerParams
from net.dropoutlayer import DropoutLayerParams
from util.theano_helpers import ReLU


class PoseRegNetParams(NetBaseParams):
    def __init__(self, type=0, nChan=1, wIn=128, hIn=128, batchSize=128, numJoints=16, nDims=3):
        super(PoseRegNetParams, self).__init__()
        self.batch_size = batchSize
        self.numJoints = numJoints
        self.nDims = nDims
        self.inputDim = (batchSize, nChan, hIn, wIn)
        if type not in (0, 11):
            raise NotImplementedError("not implemented")
        L = self.layers
        L.append(ConvPoolLayerParams(inputDim=(batchSize, nChan, hIn, wIn), nFilters=8, filterDim=(5, 5),
                                     poolsize=(4, 4), activation=ReLU))
        L.append(ConvPoolLayerParams(inputDim=L[-1].outputDim, nFilters=8, filterDim=(5, 5), poolsize=(2, 2),
                                     activation=ReLU))
        L.append(ConvPoolLayerParams(inputDim=L[-1].outputDim, nFilters=8, filterDim=(3, 3), poolsize=(1, 1),
                                     activation=ReLU))
        l3out = L[-1].outputDim
        L.append(HiddenLayerParams(inputDim=(l3out[0], l3out[1] * l3out[2] * l3out[3]), outputDim=(batchSize, 1024),
                                   activation=ReLU))
        L.append(DropoutLayerParams(inputDim=L[-1].outputDim, outputDim=L[-1].outputDim))
        L.append(HiddenLayerParams(inputDim=L[-1].outputDim, outputDim=(batchSize, 1024), activation=ReLU))
        L.append(DropoutLayerParams(inputDim=L[-1].outputDim, outputDim=L[-1].outputDim))
        if type == 11:
            L.append(HiddenLayerParams(inputDim=L[-1].outputDim, outputDim=(batchSize, 30), activation=None))
        L.append(HiddenLayerParams(inputDim=L[-1].outputDim, outputDim=(batchSize, numJoints * nDims),
                                   activation=None))
        self.outputDim = L[-1].outputDim


class PoseRegNet(NetBase):
    def __init__(self, rng, inputVar=None, cfgParams=None):
        if cfgParams is None:
            raise Exception("Cannot create a Net without config parameters (ie. cfgParams==None)")
        if inputVar is None:
            inputVar = tensor4('x')
        elif isinstance(inputVar, str):
            inputVar = tensor4(inputVar)
        super(PoseRegNet, self).__init__(rng, inputVar, cfgParams)
