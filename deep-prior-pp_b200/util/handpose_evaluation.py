"""HandposeEvaluation - the numeric error metrics of the reference's evaluation class (reference:
src/util/handpose_evaluation.py:44-91 constructor, :92-181 mean / std / max / per-joint errors, :197-228 frames
within a distance).  Every metric is a reduction of ``sqrt(square(gt - joints).sum(axis=2))``; that (n, J) error
matrix and its per-frame nan-mean / nan-max are computed once on the device by ``dpp_joint_errors``
(csrc/recrop.cu) and kept there; the getters reduce those device tensors, so a validation loop never moves the
joints back to the host.  Plotting (matplotlib) and the dataset-specific subclasses' skeleton drawings are out
of scope (SURVEY 8f row f4 names the metrics only)."""
import numpy


class HandposeEvaluation(object):
    def __init__(self, gtjoints, joints, dolegend=True, linewidth=1):
        if not isinstance(gtjoints, (numpy.ndarray, list)) or not isinstance(joints, (numpy.ndarray, list)):
            raise ValueError("Params must be list or ndarray")
        if len(gtjoints) != len(joints):
            print("Error: groundtruth has {} elements, eval data has {}".format(len(gtjoints), len(joints)))
            raise ValueError("Params must be the same size")
        if len(gtjoints) == len(joints) == 0:
            print("Error: groundtruth has {} elements, eval data has {}".format(len(gtjoints), len(joints)))
            raise ValueError("Params must be of non-zero size")
        if gtjoints[0].shape != joints[0].shape:
            print("Error: groundtruth has {} dims, eval data has {}".format(gtjoints[0].shape, joints[0].shape))
            raise ValueError("Params must be of same dimensionality")
        self.gtjoints = numpy.asarray(gtjoints)
        self.joints = numpy.asarray(joints)
        assert (self.gtjoints.shape == self.joints.shape)
        self.linewidth = linewidth
        self.dolegend = dolegend
        self.subfolder = './eval/'
        self.jointNames = None
        self._dev = None

    # -- device-side error matrix --------------------------------------------------------------------------------
    def _errors(self):
        """(err (n,J), frame nan-mean (n,), frame nan-max (n,)) torch CUDA tensors, computed once."""
        if self._dev is None:
            from dpp_b200.cascade import joint_errors
            self._dev = joint_errors(self.joints.astype('float32'), self.gtjoints.astype('float32'))
        return self._dev

    @staticmethod
    def _nanmean(t):
        ok = ~t.isnan()
        return float((t.nan_to_num(0.) * ok).sum() / ok.sum())

    # -- handpose_evaluation.py:92-181 --------------------------------------------------------------------------
    def getMeanError(self):
        return self._nanmean(self._errors()[1])

    def getStdError(self):
        err = self._errors()[0]
        ok = ~err.isnan()
        cnt = ok.sum(dim=1)
        mean = (err.nan_to_num(0.) * ok).sum(dim=1) / cnt
        var = (((err - mean[:, None]).nan_to_num(0.) * ok) ** 2).sum(dim=1) / cnt        # numpy.nanstd: ddof = 0
        return self._nanmean(var.sqrt())

    def getMeanErrorOverSeq(self):
        return self._errors()[1].cpu().numpy()

    def getMedianError(self):
        err = self._errors()[0].reshape(-1)
        return float(err[~err.isnan()].double().quantile(0.5))      # scipy nanmedian of the flattened errors

    def getMaxError(self):
        m = self._errors()[2]
        return float(m[~m.isnan()].max())

    def getMaxErrorOverSeq(self):
        return self._errors()[2].cpu().numpy()

    def getJointMeanError(self, jointID):
        return self._nanmean(self._errors()[0][:, jointID])

    def getJointStdError(self, jointID):
        e = self._errors()[0][:, jointID]
        e = e[~e.isnan()]
        return float(((e - e.mean()) ** 2).mean().sqrt())

    def getJointErrorOverSeq(self, jointID):
        return self._errors()[0][:, jointID].cpu().numpy()

    def getJointDiffOverSeq(self, jointID):
        return self.gtjoints[:, jointID, :] - self.joints[:, jointID, :]

    def getJointMaxError(self, jointID):
        e = self._errors()[0][:, jointID]
        return float(e[~e.isnan()].max())

    # -- handpose_evaluation.py:197-228 -------------------------------------------------------------------------
    def getNumFramesWithinMaxDist(self, dist):
        return int((self._errors()[2] <= dist).sum())

    def getNumFramesWithinMeanDist(self, dist):
        return int((self._errors()[1] <= dist).sum())

    def getNumFramesWithinMedianDist(self, dist):
        return int((self._errors()[0].double().quantile(0.5, dim=1) <= dist).sum())      # numpy.median (NaN propagates)

    def getJointNumFramesWithinMaxDist(self, dist, jointID):
        return int((self._errors()[0][:, jointID] <= dist).sum())

    # -- plots (matplotlib / vtk): out of scope ------------------------------------------------------------------
    def plotEvaluation(self, basename, methodName='Our method', baseline=None):
        raise NotImplementedError("evaluation plots (matplotlib) are outside the B200 path; the metrics are available "
                                  "through the get* methods")

    plotResult = plotJoints = plotResult3D = plotEvaluation


class ICVLHandposeEvaluation(HandposeEvaluation):
    """handpose_evaluation.py:684-737 without the plot styling: 16 joints (C, T1-3, I1-3, M1-3, R1-3, P1-3)."""

    def __init__(self, gt, joints, dolegend=True, linewidth=1):
        super(ICVLHandposeEvaluation, self).__init__(gt, joints, dolegend, linewidth)
        self.jointNames = ['C'] + ['%s%d' % (f, k) for f in 'TIMRP' for k in (1, 2, 3)]
        self.plotMaxJointDist = 80
        self.fps = 10.0


class NYUHandposeEvaluation(HandposeEvaluation):
    """handpose_evaluation.py:740-850 without the plot styling; ``joints`` selects the 14 evaluation joints
    ('eval', the entry scripts' setting) or all 36."""

    def __init__(self, gt, joint, joints='eval', dolegend=True, linewidth=1):
        super(NYUHandposeEvaluation, self).__init__(gt, joint, dolegend, linewidth)
        if joints not in ('eval', 'all'):
            raise ValueError("Unknown joint parameter")
        self.plotMaxJointDist = 80
        self.fps = 25.0


class MSRAHandposeEvaluation(HandposeEvaluation):
    """handpose_evaluation.py:853-910 without the plot styling: 21 joints (wrist + 4 per finger)."""

    def __init__(self, gt, joints, dolegend=True, linewidth=1):
        super(MSRAHandposeEvaluation, self).__init__(gt, joints, dolegend, linewidth)
        self.jointNames = ['C'] + ['%s%d' % (f, k) for f in 'TIMRP' for k in (1, 2, 3, 4)]
        self.plotMaxJointDist = 80
        self.fps = 20.0
