"""PoseRegNetTrainer (reference: src/trainer/poseregnettrainer.py:44-264): cost = mean over the
batch of the summed squared embedding error (:92-99), gradients of every parameter, ADAM update,
validation cost / error / PCA back-projected joint errors (:113-129).  ``train_model`` and the
validation functions are engine calls instead of compiled Theano functions."""
import ctypes as C
import numpy

from trainer.nettrainer import NetTrainer, NetTrainerParams
from trainer.optimizer import Optimizer


class PoseRegNetTrainerParams(NetTrainerParams):
    def __init__(self):
        super(PoseRegNetTrainerParams, self).__init__()


class PoseRegNetTrainer(NetTrainer):
    def __init__(self, poseNet=None, cfgParams=None, rng=None, subfolder='./eval/', numChunks=1):
        super(PoseRegNetTrainer, self).__init__(cfgParams, 5, subfolder, numChunks)
        self.poseNet = poseNet
        self.rng = rng if rng is not None else self.rng
        if not isinstance(cfgParams, PoseRegNetTrainerParams):
            raise ValueError("cfgParams must be an instance of PoseRegNetTrainerParams")
        self.setupFunctions()

    def setupFunctions(self):
        cfg = self.poseNet.cfgParams
        if not (cfg.numJoints == 1 and cfg.nDims > 1):
            raise NotImplementedError("only the embedding branch (numJoints == 1, poseregnettrainer.py:94-95) is built")
        if self.cfgParams.weightreg_factor != 0.0 and not self.poseNet.hasDropout():
            raise NotImplementedError("weight regularisation is 0 in the reference's entry scripts")
        self.params = self.poseNet.params
        self.grads = None

    def compileFunctions(self, compileDebugFcts=False):
        self.setupTrain()
        self.compileDebugFcts = compileDebugFcts
        self.setupValidate()

    # -- helpers ---------------------------------------------------------------------------------
    def _load_batch(self, xsrc, index):
        """minibatch ``index`` of a device-resident set (this rank's rows of it in a data-parallel job)"""
        eng = self._engine()
        B = eng.B
        eng.t_in.buf.view(B, -1).copy_(xsrc[index * B:(index + 1) * B].reshape(B, -1))
        return eng

    def setupTrain(self):
        opt = Optimizer(self.grads, self.params)
        self.updates = opt.ADAM()

        def train_model(index, learning_rate):
            self._to_device()
            eng = self._load_batch(self._dev['train_x'], index)
            eng._alloc_training()
            B = eng.B
            eng.y_in.copy_(self._dev['train_y'][index * B:(index + 1) * B])
            cost = eng.train_step(float(learning_rate), use_graph=getattr(self, 'use_graph', True))
            # the mean over the GLOBAL minibatch (poseregnettrainer.py:92-99): equal shares, so the mean of the means
            return numpy.float32(self._allreduce_scalar(float(cost.cpu()[0])))
        self.train_model = train_model

        def test_model_on_train(index):
            return self._eval(self._dev['train_x'], self._dev['train_y'], index, 'errors')
        self.test_model_on_train = test_model_on_train

    def _eval(self, xsrc, ysrc, index, what):
        """deterministic/probabilistic forward on one batch + cost / error observers"""
        import torch
        self._to_device()
        eng = self._load_batch(xsrc, index)
        out = eng.forward_device(deterministic=self.poseNet.isDeterministic())
        B = eng.B
        y = ysrc[index * B:(index + 1) * B]
        if what == 'cost':
            return self._allreduce_scalar(float(((out - y) ** 2).sum(dim=1).mean().cpu()))
        if what == 'errors':
            return self._allreduce_scalar(float(torch.sqrt(((out - y) ** 2).sum(dim=1)).mean().cpu()))
        pca = self._static('pca_data')
        mean = self._static('mean_data')
        y3 = self._dev['val_y3D'][index * B:(index + 1) * B]
        j = (out @ pca + mean).reshape(B, -1, 3) - y3.reshape(B, -1, 3)
        e = torch.sqrt((j ** 2).sum(dim=2))
        if what == 'avg':
            return self._allreduce_scalar(float(e.mean(dim=1).mean().cpu()))
        return self._allreduce_scalar(float(e.max(dim=1)[0].max().cpu()), op='max')

    def _static(self, key):
        import torch
        k = '_st_' + key
        if k not in self._dev:
            self._dev[k] = torch.from_numpy(numpy.ascontiguousarray(getattr(self, key + 'DB'), dtype='float32')).to(
                self._engine().dev)
        return self._dev[k]

    def setupValidate(self):
        self.validation_cost = lambda index: self._eval(self._dev['val_x'], self._dev['val_y'], index, 'cost')
        self.validation_observer.append(self.validation_cost)
        self.validation_error = lambda index: self._eval(self._dev['val_x'], self._dev['val_y'], index, 'errors')
        self.validation_observer.append(self.validation_error)
        if hasattr(self, 'val_data_y3DDB'):
            self.validation_error_avg = lambda index: self._eval(self._dev['val_x'], self._dev['val_y'], index, 'avg')
            self.validation_error_max = lambda index: self._eval(self._dev['val_x'], self._dev['val_y'], index, 'max')
            self.validation_observer.append(self.validation_error_avg)
            self.validation_observer.append(self.validation_error_max)

    def augment_poses(self, macro_params, macro_idx, last, tidxs, idxs, new_data):
        """poseregnettrainer.py:221-264 with the reference signature (per-sample slow path; the
        trainer's own epoch pipeline batches the records instead)."""
        a = macro_params['args']
        for idx, i in zip(tidxs, idxs):
            img = self.train_data_xDB[i, 0].copy()
            com = a['di'].joint3DToImg(self.train_data_comDB[i])
            cube = self.train_data_cubeDB[i].copy()
            M = self.train_data_MDB[i].copy()
            gt3Dcrop = self.train_gt3DcropDB[i].copy()
            imgD, _, curLabel, cube, com2D, M, _ = self.augmentCrop(
                img, gt3Dcrop, com, cube, M, a['aug_modes'], a['hd'], a['normZeroOne'],
                sigma_com=a.get('sigma_com'), sigma_sc=a.get('sigma_sc'), rot_range=a.get('rot_range'))
            new_data['train_data_x'][idx] = imgD
            if a.get('proj') is not None:
                new_data['train_data_y'][idx] = a['proj'].transform(curLabel.reshape(1, -1))[0]
            else:
                new_data['train_data_y'][idx] = curLabel.reshape(new_data['train_data_y'][idx].shape)
