"""RealtimeHandposePipeline - the inference cascade of the deployment demo (reference:
src/util/realtimehandposepipeline.py:66-131 constructor + initNets, :296-333 detect, :335-368 estimatePose,
:197-198 pose de-normalisation; driven by src/test_realtimepipeline.py:61-67).

Kept: the constructor signature, ``initNets``, ``detect(frame)`` (tracking branch: the last CoM is refined by the
ScaleNet and the hand is re-cropped around it), ``estimatePose(crop, com3D)``.  Added: ``processBatch`` - the same
cascade over a whole batch of frames resident in HBM (``dpp_b200.cascade.Cascade``; BASELINE config 5).
Not here: camera capture, the producer/consumer processes, the cv2 windows and keyboard handling
(:133-294, :370-560) and the contour-based first detection ``HandDetector.detect`` (CPU, serial: out of scope,
SURVEY 8a) - the first CoM comes from the caller (``lastcom``)."""
import copy
import numpy

from net.poseregnet import PoseRegNet, PoseRegNetParams
from net.resnet import ResNet, ResNetParams
from net.scalenet import ScaleNet, ScaleNetParams
from util.handdetector import HandDetector


class RealtimeHandposePipeline(object):
    # states of pipeline
    STATE_IDLE = 0
    STATE_INIT = 1
    STATE_RUN = 2

    # different hands
    HAND_LEFT = 0
    HAND_RIGHT = 1

    # different detectors
    DETECTOR_COM = 0

    def __init__(self, poseNet, config, di, verbose=False, comrefNet=None):
        self.importer = di
        self.poseNet = poseNet
        self.comrefNet = comrefNet
        self.initialconfig = copy.deepcopy(config)
        self.config = config                      # the reference keeps it in a multiprocessing Manager dict
        self.verbose = verbose
        self.hand = self.HAND_LEFT
        self.state = self.STATE_RUN
        self.tracking = True
        self.lastcom = (0, 0, 0)
        self._cascade = None

    def initNets(self):
        """realtimehandposepipeline.py:111-131: build the nets from their parameter objects and force the first
        (engine-building) forward pass."""
        if isinstance(self.poseNet, PoseRegNetParams):
            self.poseNet = PoseRegNet(numpy.random.RandomState(23455), cfgParams=self.poseNet)
            self.poseNet.computeOutput(numpy.zeros(self.poseNet.cfgParams.inputDim, dtype='float32'))
        elif isinstance(self.poseNet, ResNetParams):
            self.poseNet = ResNet(numpy.random.RandomState(23455), cfgParams=self.poseNet)
            self.poseNet.computeOutput(numpy.zeros(self.poseNet.cfgParams.inputDim, dtype='float32'))
        elif not hasattr(self.poseNet, 'computeOutput'):
            raise RuntimeError("Unknown pose estimation method!")
        if self.comrefNet is not None:
            if isinstance(self.comrefNet, ScaleNetParams):
                self.comrefNet = ScaleNet(numpy.random.RandomState(23455), cfgParams=self.comrefNet)
                self.comrefNet.computeOutput([numpy.zeros(sz, dtype='float32') for sz in self.comrefNet.cfgParams.inputDim])
            elif not hasattr(self.comrefNet, 'computeOutput'):
                raise RuntimeError("Unknown refine method!")

    # -- per-frame surface ------------------------------------------------------------------------------------
    def detect(self, frame):
        """:296-333, tracking branch.  Returns (crop normalised to [-1, 1], M, com3D)."""
        hd = HandDetector(frame, self.config['fx'], self.config['fy'], importer=self.importer, refineNet=self.comrefNet)
        if not self.tracking or numpy.allclose(self.lastcom, 0):
            raise NotImplementedError("first detection (cv2 contours) is outside the B200 path: set lastcom")
        loc, _ = hd.track(self.lastcom, self.config['cube'], doHandSize=False)
        self.lastcom = loc
        dim = self.poseNet.layers[0].cfgParams.inputDim
        if numpy.allclose(loc, 0):
            return numpy.zeros((dim[2], dim[3]), dtype='float32'), numpy.eye(3), loc
        crop, M, com = hd.cropArea3D(com=loc, size=self.config['cube'], dsize=(dim[2], dim[3]))
        com3D = self.importer.jointImgTo3D(com)
        sc = (self.config['cube'][2] / 2.)
        crop[crop == 0] = numpy.float32(numpy.float64(com3D[2]) + sc)
        crop -= com3D[2]
        crop /= numpy.float32(sc)
        return crop, M, com3D

    def estimatePose(self, crop, com3D):
        """:335-368."""
        if self.hand == self.HAND_LEFT:
            inp = crop[None, None, :, :].astype('float32')
        else:
            inp = crop[None, None, :, ::-1].astype('float32')
        jts = self.poseNet.computeOutput(numpy.ascontiguousarray(inp))
        jj = jts[0].reshape((-1, 3))
        if self.config.get('invX') is True:
            jj[:, 1] *= (-1.)
        if self.config.get('invY') is True:
            jj[:, 0] *= (-1.)
        if self.hand == self.HAND_RIGHT:
            jj[:, 0] *= (-1.)
        return jj

    def processFrame(self, frame):
        """detect + estimatePose + de-normalisation (:163-198 without the display): joints (J,3) in mm."""
        crop, M, com3D = self.detect(frame)
        pose = self.estimatePose(crop, com3D)
        return pose * numpy.float32(self.config['cube'][2]) / numpy.float32(2.) + com3D

    # -- batched surface ---------------------------------------------------------------------------------------
    def processBatch(self, frames, lastcoms, ndvalue=None):
        """The cascade over a batch of frames (numpy (n,H,W) or torch CUDA) and their previous CoMs (n,3).
        Returns the dict of ``dpp_b200.cascade.Cascade.run`` (pose (n,J,3) in mm, com, com3D, M)."""
        from dpp_b200.cascade import Cascade
        if self._cascade is None:
            self._cascade = Cascade(self.poseNet, self.comrefNet, self.importer, self.config['fx'], self.config['fy'],
                                    self.config['cube'])
        res = self._cascade.run(frames, lastcoms, ndvalue=ndvalue, right_hand=(self.hand == self.HAND_RIGHT))
        if self.config.get('invX') is True:
            res['pose_norm'][:, :, 1] *= (-1.)
        if self.config.get('invY') is True:
            res['pose_norm'][:, :, 0] *= (-1.)
        if self.config.get('invX') is True or self.config.get('invY') is True:
            res['pose'] = (res['pose_norm'] * numpy.float32(self.config['cube'][2]) / numpy.float32(2.)
                           + res['com3D'][:, None, :]).astype('float32')
        return res
