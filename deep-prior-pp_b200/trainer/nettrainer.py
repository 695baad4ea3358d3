"""NetTrainer (reference: src/trainer/nettrainer.py:47-997): training-data residency, the
epoch/minibatch loop with the learning-rate schedule, validation and best-weights snapshot, and
``augmentCrop``.

B200 redesign of the data path: the reference splits host RAM into macro batches
that are copied to the GPU one at a time while 8 worker processes augment the next one with cv2.
Here the whole (aligned) training set lives in HBM once; per epoch the host only prepares one
112-byte ``dpp_aug_rec`` per sample (in ``para_num_proc`` worker processes when ``para_augment`` is
set, mirroring nettrainer.py:666-689) and ``dpp_augment_fwd`` regenerates the augmented set on the
device from the ORIGINAL crops.  What is kept: ceil(N/B) minibatches per epoch, the tail-padding
rule with ``RandomState(N)`` (:365-413), fresh augmentation per epoch swapped in when the last
minibatch of an epoch is requested (:528-529), ``lr_of_ep`` (:54), validation every
``validation_frequency`` iterations on full batches only (:796,859), ``net_last.pkl`` snapshots,
NaN abort, best-weights restore."""
import os
import time
import multiprocessing
import numpy

from net.convpoollayer import ConvPoolLayer
from net.convlayer import ConvLayer


class NetTrainerParams(object):
    def __init__(self):
        self.batch_size = 128
        self.momentum = 0.9
        self.learning_rate = 0.01
        self.weightreg_factor = 0.001
        self.use_early_stopping = True
        # nettrainer.py:54
        self.lr_of_ep = lambda ep: numpy.float32(self.learning_rate / 10.) if ep <= 1 else \
            numpy.float32(self.learning_rate / 3.) if 1 < ep <= 2 else \
            numpy.float32(self.learning_rate * numpy.exp(-0.04 * ep))
        self.snapshot_last = 5
        self.snapshot_freq = None
        self.para_augment = False
        self.para_num_proc = 8
        self.augment_fun_params = {'fun': None, 'args': {}}
        self.para_load = False
        self.load_fun_params = {'fun': None, 'args': {}}
        self.force_macrobatch_reload = False
        self.pad_random = True
        self.validation_frequency = 1000
        self.pre_epoch_fn = None
        self.post_epoch_fn = None
        self.pre_minibatch_fn = None
        self.post_minibatch_fn = None


def _draw_all(rng, n, n_modes, sigma_com, sigma_sc, rot_range):
    """the n draws of one epoch in the reference's order (nettrainer.py:954-957): mode, CoM offset, rotation, scale"""
    ri, rn, ru = rng.randint, rng.randn, rng.uniform
    draws = []
    for _ in range(n):
        mode = ri(0, n_modes)
        off = rn(3) * sigma_com
        rot = ru(-rot_range, rot_range)
        sc = abs(1. + rn() * sigma_sc)
        draws.append((mode, off, rot, sc))
    return draws


def _records_chunk(args):
    """worker: build augmentation records + labels for a slice of samples (one vectorised pass,
    ``HandDetector.aug_records_batch``: bit-identical to the per-sample ``aug_record``).  ``draws`` is either the
    slice's list of draws or ('replay', rng state, n, draw parameters, rows): the worker then replays the epoch's
    WHOLE draw sequence from the trainer's generator state (the sequence is serial; replaying it in every worker in
    parallel keeps its ~13 us per sample off the thread that launches the training steps) and takes its rows;
    the generator state after the sequence travels back with the result."""
    trainer_state, idxs, draws = args[:3]
    src_rows = args[3] if len(args) > 3 else idxs
    state_after = None
    if isinstance(draws, tuple) and len(draws) == 5 and draws[0] == 'replay':
        _, st, n_all, dpar, rows = draws
        rng = numpy.random.RandomState()
        rng.set_state(st)
        alld = _draw_all(rng, n_all, *dpar)
        state_after = rng.get_state()
        draws = [alld[i] for i in rows]
    out = _records_chunk_drawn(trainer_state, idxs, draws, src_rows)
    return out if state_after is None else out + (state_after,)


def _records_chunk_drawn(trainer_state, idxs, draws, src_rows):
    hd, di, aug_modes, comDB, cubeDB, MDB, gtDB, proj = trainer_state
    idxs = numpy.asarray(idxs, dtype=numpy.int64)
    if len(idxs) == 0:
        from dpp_b200.lib import AUG_REC_DTYPE
        return numpy.zeros(0, dtype=AUG_REC_DTYPE), numpy.zeros((0, 0), dtype='float32')
    com = hd._toimg(numpy.asarray(comDB, dtype='float32')[idxs])           # di.joint3DToImg, vectorised
    # idxs index the host arrays (global rows); src_rows are the rows of the device-resident crops the records read
    # (the same numbers on one device, this rank's local rows in a data-parallel job)
    recs, labels = hd.aug_records_batch(numpy.asarray(src_rows, dtype=numpy.int64),
                                        [aug_modes[d[0]] for d in draws], numpy.array([d[1] for d in draws]),
                                        numpy.array([d[2] for d in draws]), numpy.array([d[3] for d in draws]), com,
                                        numpy.asarray(cubeDB)[idxs], numpy.asarray(MDB)[idxs], numpy.asarray(gtDB)[idxs])
    labels = labels.reshape(len(idxs), -1)
    if proj is not None:
        labels = proj.transform(labels)                  # poseregnettrainer.py:262, all rows at once
    return recs, numpy.asarray(labels, dtype='float32')


class NetTrainer(object):
    def __init__(self, cfgParams, memory_factor, subfolder='./eval/', numChunks=1):
        self.subfolder = subfolder
        self.cfgParams = cfgParams
        self.rng = numpy.random.RandomState(23455)
        if not isinstance(cfgParams, NetTrainerParams):
            raise ValueError("cfgParams must be an instance of NetTrainerParams")
        # nettrainer.py:100-112: a fraction of the free device memory bounds one resident array
        try:
            import torch
            free = torch.cuda.mem_get_info()[0] if torch.cuda.is_available() else None
        except Exception:
            free = None
        if free is None:
            import psutil
            free = psutil.virtual_memory().available
        self.memorySize = (free / 1024 ** 2) / float(memory_factor)
        if cfgParams.para_load is True:
            raise NotImplementedError("para_load (chunked loading from disk) is out of scope: the set is HBM resident")
        self.currentMacroBatch = -1
        self.numChunks = numChunks
        self.trainSize = 0
        self.sampleSize = 0
        self.numTrainSamplesMB = 0
        self.numTrainSamples = 0
        self.numValSamples = 0
        self.epoch = 0
        self.managedVar = []
        self.trainingVar = []
        self.validation_observer = []
        self._pool = None
        self._dev = {}
        # record workers (nettrainer.py:666-689 starts para_num_proc augmentation processes): forked HERE, from a
        # process that has neither a CUDA context nor NCCL threads yet - the workers are plain NumPy processes
        if cfgParams.para_augment and cfgParams.para_num_proc > 1:
            self._pool = multiprocessing.get_context('fork').Pool(cfgParams.para_num_proc)
        # data parallelism (not in the single-device reference): under torchrun the global minibatch of
        # cfgParams.batch_size samples is split over the ranks (strong scaling, SURVEY 8e / BASELINE config 3)
        from dpp_b200 import dp
        self._dist, self.rank, self.world = dp.init_process_group()
        self.local_batch = dp.local_batch(cfgParams.batch_size, self.world)
        self.syncbn = {'1': True, 'nccl': True, 'p2p': 'p2p'}.get(os.environ.get('DPP_SYNCBN', '0'), False)
        self.verbose = getattr(self, 'verbose', True) and self.rank == 0

    # -- data registration ------------------------------------------------------------------
    def setData(self, train_data, train_y, val_data, val_y, max_train_size=0):
        if (train_data.shape[0] != train_y.shape[0]) or (val_data.shape[0] != val_y.shape[0]):
            raise ValueError("Number of samples must be the same as number of labels.")
        self.trainSize = max(train_data.nbytes, train_y.nbytes, max_train_size) / 1024. / 1024.
        self.numTrainSamplesMB = train_data.shape[0]
        self.numTrainSamples = self.numTrainSamplesMB
        self.numValSamples = val_data.shape[0]
        self.sampleSize = self.trainSize / self.numTrainSamplesMB
        assert self.memorySize > self.sampleSize * self.cfgParams.batch_size
        if self.getNumMacroBatches() != 1:
            raise NotImplementedError("training set larger than the HBM budget (%.0f MB > %.0f MB): sharding over "
                                      "data-parallel ranks is the supported way to scale" % (self.trainSize, self.memorySize))
        # shrink to the smallest whole number of minibatches (nettrainer.py:257-258)
        self.memorySize = self.sampleSize * numpy.ceil(self.numTrainSamplesMB / float(self.cfgParams.batch_size)) \
            * self.cfgParams.batch_size
        self.train_data_xDB = self.alignData(train_data)
        self.train_data_yDB = self.alignData(train_y)
        self.trainingVar += ['train_data_x', 'train_data_y']
        self.val_data_xDB = val_data
        self.val_data_yDB = val_y
        if self.rank == 0:
            print("{} train samples, {} val samples, batch size {}".format(train_data.shape[0], val_data.shape[0],
                                                                           self.cfgParams.batch_size))
            print("{} macro batches, {} mini batches per macro, {} full mini batches total".format(
                self.getNumMacroBatches(), self.getNumMiniBatchesPerMacroBatch(), self.getNumMiniBatches()))
            if self.world > 1:
                print("data parallel: {} ranks x {} samples per minibatch{}".format(
                    self.world, self.local_batch, ", SyncBN" if self.syncbn else ", per-replica BatchNorm statistics"))
        self._dev_dirty = True

    def addData(self, data):
        if not isinstance(data, dict):
            raise ValueError("Error: expected dictionary for data!")
        for key in data:
            setattr(self, key + 'DB', self.alignData(data[key]))
        self._dev_dirty = True

    def addStaticData(self, data):
        if not isinstance(data, dict):
            raise ValueError("Error: expected dictionary for data!")
        for key in data:
            setattr(self, key + 'DB', data[key])
        self._dev_dirty = True

    def addManagedData(self, data):
        if not isinstance(data, dict):
            raise ValueError("Error: expected dictionary for data!")
        for key in data:
            if data[key].shape[0] != self.numTrainSamplesMB:
                raise ValueError("Number of samples must be the same as number of labels.")
            setattr(self, key + 'DB', self.alignData(data[key]))
            self.trainingVar.append(key)

    def alignData(self, data, alignSize=None, out=None, fillData=None):
        """nettrainer.py:365-413: pad to the macro-batch size; the tail is filled with random real
        samples drawn with RandomState(N) so that x / y / side arrays stay aligned."""
        if alignSize is None:
            alignSize = self.getNumSamplesPerMacroBatch()
        if data.shape[0] == alignSize:
            topad = 0
        else:
            topad = alignSize - data.shape[0] % alignSize
        sz = [(0, topad)] + [(0, 0)] * (len(data.shape) - 1)
        padded = numpy.pad(data, sz, mode='constant', constant_values=0)
        if fillData is None:
            fillData = data
        if (data.shape[0] % alignSize) != 0:
            if self.cfgParams.pad_random:
                rng = numpy.random.RandomState(data.shape[0])
                for i in range(0, alignSize - (data.shape[0] % alignSize)):
                    padded[data.shape[0] + i] = fillData[rng.randint(0, fillData.shape[0])]
            else:
                for i in range(0, alignSize - (data.shape[0] % alignSize)):
                    padded[data.shape[0] + i] = padded[data.shape[0] - 1]
        return padded

    # -- batch arithmetic (nettrainer.py:415-487) ------------------------------------------------
    def getSizeMiniBatch(self):
        return self.cfgParams.batch_size * self.sampleSize

    def getSizeMacroBatch(self):
        """size of a macro batch in MB as the reference defines it (nettrainer.py:422-427)"""
        return self.getNumMacroBatches() * self.getSizeMiniBatch()

    def getNumFullMiniBatches(self):
        return self.getNumMiniBatches()

    def getNumMiniBatches(self):
        return int(numpy.ceil(self.numTrainSamples / float(self.cfgParams.batch_size)))

    def getNumMacroBatches(self):
        return int(numpy.ceil(self.trainSize / float(self.getGPUMemAligned())))

    def getNumMiniBatchesPerMacroBatch(self):
        return int(self.getGPUMemAligned() / self.sampleSize / self.cfgParams.batch_size)

    def getNumSamplesPerMacroBatch(self):
        return int(self.getNumMiniBatchesPerMacroBatch() * self.cfgParams.batch_size)

    def getNumMiniBatchesPerChunk(self):
        return int(self.getNumMiniBatchesPerMacroBatch() * self.getNumMacroBatches())

    def getNumSamplesPerChunk(self):
        return self.getNumMiniBatchesPerChunk() * self.cfgParams.batch_size

    def getGPUMemAligned(self):
        return self.sampleSize * self.cfgParams.batch_size * int(
            self.memorySize / float(self.sampleSize * self.cfgParams.batch_size))

    # -- device residency -------------------------------------------------------------------------
    def _to_device(self):
        """Upload the (aligned) training set, the validation set and the static arrays once."""
        import torch
        if not getattr(self, '_dev_dirty', True):
            return
        from dpp_b200 import dp
        dev = self._engine().dev
        f = lambda a: torch.from_numpy(numpy.ascontiguousarray(a, dtype='float32')).to(dev)
        d = self._dev
        if self.train_data_xDB.shape[1] != 1:
            raise NotImplementedError("multi-channel training crops")
        B = self.cfgParams.batch_size
        # rows of the aligned set that live on this rank: slice [rank*b, (rank+1)*b) of every global minibatch
        rows = dp.local_rows(self.train_data_xDB.shape[0], B, self.rank, self.world)
        self._rows = rows
        d['train_x_orig'] = f(self.train_data_xDB[rows].reshape((len(rows),) + self.train_data_xDB.shape[-2:]))
        d['train_x'] = d['train_x_orig'].clone()
        d['train_y'] = f(self.train_data_yDB[rows].reshape(len(rows), -1))
        nb = self.val_data_xDB.shape[0] // B * B
        vrows = dp.local_rows(nb, B, self.rank, self.world) if nb else numpy.zeros(0, dtype=numpy.int64)
        d['val_x'] = f(self.val_data_xDB[:nb][vrows].reshape((len(vrows),) + self.val_data_xDB.shape[-2:]))
        flat = lambda a: a[:nb][vrows].reshape(len(vrows), int(numpy.prod(a.shape[1:])))   # nb may be 0: fewer validation samples than a batch
        d['val_y'] = f(flat(self.val_data_yDB))
        if hasattr(self, 'val_data_y3DDB'):
            d['val_y3D'] = f(flat(self.val_data_y3DDB))
        self._dev_dirty = False

    def _engine(self):
        """the net's device executor; in a data-parallel job it is built for this rank's share of the minibatch and
        joined to the gradient (and, with DPP_SYNCBN=1, BatchNorm-statistics) exchange"""
        net = self.poseNet
        if self.world > 1:
            net._engine_opts = {'batch': self.local_batch}
        eng = net._engine()
        if self.world > 1 and (eng.world != self.world or eng.allreduce_fn is None):
            dist = self._dist
            eng.set_world(self.world, lambda t: dist.all_reduce(t), rank=self.rank, syncbn=self.syncbn, dist=dist)
        return eng

    def _allreduce_scalar(self, value, op='mean'):
        """combine a per-rank scalar (cost / error of this rank's share of a minibatch) into the global figure"""
        if self.world == 1:
            return value
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self._engine().dev)
        if op == 'max':
            self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
            return float(t[0])
        self._dist.all_reduce(t)
        return float(t[0]) / self.world

    def _sync_running_stats(self):
        """per-replica BatchNorm: the running mean / inv_std of the ranks differ slightly; average them before they
        are used (validation) or stored (snapshots), so that every rank evaluates and saves the same network"""
        if self.world > 1 and not self.syncbn:
            eng = self._engine()
            self._dist.all_reduce(eng.R)
            eng.R.mul_(1.0 / self.world)

    # -- augmentation pipeline ----------------------------------------------------------------------
    def setupDataLoading(self):
        """nettrainer.py:666-699: start the record workers and prepare the first augmented set."""
        if self.cfgParams.para_augment and self.cfgParams.para_num_proc > 1 and self._pool is None:
            self._pool = multiprocessing.get_context('fork').Pool(self.cfgParams.para_num_proc)   # a second train() call
        self._pending = self._request_augmentation()
        self._swap_augmentation()
        self._pending = self._request_augmentation()

    def unsetDataLoading(self):
        if self._pool is not None:
            self._pool.terminate()
            self._pool = None
        self._pending = None

    def _draw(self, n):
        a = self.cfgParams.augment_fun_params['args']
        none_or = lambda v, default: default if v is None else v        # nettrainer.py:942-947 tests `is None`
        sigma_com = none_or(a.get('sigma_com'), 5.)
        sigma_sc = none_or(a.get('sigma_sc'), 0.02)
        rot_range = none_or(a.get('rot_range'), 180.)
        return _draw_all(self.rng, n, len(a['aug_modes']), sigma_com, sigma_sc, rot_range)      # nettrainer.py:954-957 draw order

    def _draw_params(self):
        a = self.cfgParams.augment_fun_params['args']
        none_or = lambda v, default: default if v is None else v
        return (len(a['aug_modes']), none_or(a.get('sigma_com'), 5.), none_or(a.get('sigma_sc'), 0.02),
                none_or(a.get('rot_range'), 180.))

    def _request_augmentation(self):
        a = self.cfgParams.augment_fun_params['args']
        n_all = self.train_data_xDB.shape[0]
        self._to_device()
        idxs = [int(i) for i in self._rows]   # every rank draws for ALL samples (the stream of the single-device run)
        src = list(range(len(idxs)))          # ... and builds records only for the rows it holds
        n = len(idxs)
        replay = self._pool is not None
        if replay:
            # the draw sequence is replayed inside the workers from this generator state (see _records_chunk);
            # _swap_augmentation installs the state after the sequence before the next request is made
            st, dpar = self.rng.get_state(), self._draw_params()
            draws_of = lambda sl: ('replay', st, n_all, dpar, idxs[sl])
        else:
            draws = self._draw(n_all)
            draws = [draws[i] for i in idxs]
            draws_of = lambda sl: draws[sl]
        # the side arrays travel as the rows a job needs (a few hundred bytes per sample), not as whole arrays
        def job(sl):
            ii = numpy.asarray(idxs[sl], dtype=numpy.int64)
            g = lambda arr: numpy.asarray(arr)[ii]
            state = (a['hd'], a['di'], a['aug_modes'], g(self.train_data_comDB), g(self.train_data_cubeDB),
                     g(self.train_data_MDB), g(self.train_gt3DcropDB), a.get('proj'))
            return (state, list(range(len(ii))), draws_of(sl), src[sl])
        if replay:
            k = self.cfgParams.para_num_proc
            step = (n + k - 1) // k
            return ('async', self._pool.map_async(_records_chunk, [job(slice(s, s + step)) for s in range(0, max(n, 1), step)]))
        return ('sync', [_records_chunk(job(slice(0, n)))])

    def _swap_augmentation(self):
        """wait for the records, regenerate train_x/train_y on the device (one kernel launch)"""
        import torch
        from dpp_b200.augment import run_records_device
        kind, res = self._pending
        parts = res.get() if kind == 'async' else res
        if kind == 'async' and len(parts[0]) > 2:
            self.rng.set_state(parts[0][2])          # the generator as the epoch's draw sequence leaves it
        recs = numpy.concatenate([p[0] for p in parts])
        labels = numpy.concatenate([p[1] for p in parts])
        self._to_device()
        d = self._dev
        run_records_device(d['train_x_orig'], recs, out_dev=d['train_x'])
        d['train_y'].copy_(torch.from_numpy(labels.reshape(d['train_y'].shape)))

    def loadMiniBatch(self, mini_idx):
        """nettrainer.py:489-599 for the single-macro-batch case: when the last minibatch of an
        epoch is requested and force_macrobatch_reload is set, the freshly augmented set is swapped
        in and the next one is requested."""
        n = self.getNumMiniBatchesPerMacroBatch()
        aug = self.cfgParams.augment_fun_params['fun'] is not None
        if aug and self.cfgParams.force_macrobatch_reload and (mini_idx % n) == n - 1:
            self._swap_augmentation()
            self._pending = self._request_augmentation()
        return mini_idx % n

    # -- training loop (nettrainer.py:778-907) -------------------------------------------------------
    def train(self, n_epochs=50, storeFilters=False):
        if len(self.validation_observer) < 1:
            raise ValueError("Require at least 1 validation function, that monitors validation cost!")
        self._to_device()
        if self.cfgParams.augment_fun_params['fun'] is not None:
            self.setupDataLoading()
        wvals = []
        n_val_batches = self.val_data_xDB.shape[0] // self.cfgParams.batch_size
        best_validation_loss = numpy.inf
        bestParams = None
        bestParamsEp = -1
        start_time = time.time()
        train_costs = []
        validation_obs = [[] for x in range(1, len(self.validation_observer))]
        self.epoch = 0
        self.poseNet.setDeterministic()
        for vi in range(1, len(self.validation_observer)):
            validation_obs[vi - 1].append(numpy.nanmean([self.validation_observer[vi](i) for i in range(n_val_batches)]))
        self.poseNet.unsetDeterministic()
        while self.epoch < n_epochs:
            if self.epoch % self.cfgParams.snapshot_last == 0:
                self._sync_running_stats()
                self.poseNet.save(self.subfolder + '/net_last.pkl')        # rank 0 writes (NetBase.save)
            if self.cfgParams.snapshot_freq is not None:
                if self.epoch % self.cfgParams.snapshot_freq == 0:
                    self._sync_running_stats()
                    self.poseNet.save(self.subfolder + '/net_{}.pkl'.format(self.epoch))
            if self.cfgParams.pre_epoch_fn is not None:
                getattr(self, self.cfgParams.pre_epoch_fn)()
            self.epoch += 1
            learning_rate = self.cfgParams.lr_of_ep(self.epoch)
            for minibatch_index in range(self.getNumFullMiniBatches()):
                if self.cfgParams.pre_minibatch_fn is not None:
                    getattr(self, self.cfgParams.pre_minibatch_fn)()
                self.poseNet.unsetDeterministic()
                mini_idx = self.loadMiniBatch(minibatch_index)
                minibatch_avg_cost = self.train_model(mini_idx, learning_rate)
                if getattr(self, 'verbose', True):
                    print("minibatch {0:4d}, average cost: {1}".format(minibatch_index, minibatch_avg_cost))
                if numpy.any(numpy.isnan(minibatch_avg_cost)):
                    self.checkNaNs()
                    assert False
                train_costs.append(minibatch_avg_cost)
                if self.cfgParams.post_minibatch_fn is not None:
                    getattr(self, self.cfgParams.post_minibatch_fn)()
                iter_count = (self.epoch - 1) * self.getNumFullMiniBatches() + minibatch_index
                if (iter_count + 1) % self.cfgParams.validation_frequency == 0:
                    if storeFilters:
                        for lay in self.poseNet.layers:
                            if isinstance(lay, (ConvPoolLayer, ConvLayer)):
                                wvals.append(lay.W.get_value())
                    self._sync_running_stats()
                    self.poseNet.setDeterministic()
                    this_validation_loss = numpy.nanmean([self.validation_observer[0](i) for i in range(n_val_batches)])
                    for vi in range(1, len(self.validation_observer)):
                        validation_obs[vi - 1].append(
                            numpy.nanmean([self.validation_observer[vi](i) for i in range(n_val_batches)]))
                    self.poseNet.unsetDeterministic()
                    if self.rank == 0:
                        print("{}: epoch {}, LR {}, minibatch {}/{}, validation cost {} error {}".format(
                            time.ctime(), self.epoch, learning_rate, minibatch_index + 1, self.getNumFullMiniBatches(),
                            this_validation_loss, [vo[-1] for vo in validation_obs]))
                    if this_validation_loss < best_validation_loss:
                        best_validation_loss = this_validation_loss
                        if self.rank == 0:
                            print("Best validation loss so far, store network weights!")
                        bestParams = self.poseNet.weightVals
                        bestParamsEp = self.epoch
            if self.cfgParams.post_epoch_fn is not None:
                getattr(self, self.cfgParams.post_epoch_fn)()
        end_time = time.time()
        self._sync_running_stats()
        if self.rank == 0:
            print('Optimization complete with best validation score of %f,' % best_validation_loss)
            print('The code run for %d epochs, with %f epochs/sec' % (self.epoch, self.epoch / (end_time - start_time)))
        if bestParams is not None and self.cfgParams.use_early_stopping is True:
            self.poseNet.weightVals = bestParams
            if self.rank == 0:
                print('Best params at epoch %d' % bestParamsEp)
        if self.cfgParams.augment_fun_params['fun'] is not None:
            self.unsetDataLoading()
        return train_costs, wvals, validation_obs[0] if len(validation_obs) == 1 else validation_obs

    def checkNaNs(self):
        for param_i in self.params:
            if numpy.any(numpy.isnan(param_i.get_value())):
                print("NaN in weights", param_i.name)

    # -- augmentCrop with the reference signature (per sample; slow path) ------------------------------
    def augmentCrop(self, img, gt3Dcrop, com, cube, M, aug_modes, hd, normZeroOne=False, sigma_com=None,
                    sigma_sc=None, rot_range=None):
        """nettrainer.py:919-997.  The four random numbers are always drawn, in the reference's
        order; the pixel work runs in dpp_augment_fwd."""
        from dpp_b200.augment import run_records
        assert len(img.shape) == 2
        assert isinstance(aug_modes, list)
        if normZeroOne is True:
            raise NotImplementedError("normZeroOne=True is unused by the reference's entry scripts")
        sigma_com = 5. if sigma_com is None else sigma_com
        sigma_sc = 0.02 if sigma_sc is None else sigma_sc
        rot_range = 180. if rot_range is None else rot_range
        mode = self.rng.randint(0, len(aug_modes))
        off = self.rng.randn(3) * sigma_com
        rot = self.rng.uniform(-rot_range, rot_range)
        sc = abs(1. + self.rng.randn() * sigma_sc)
        name = aug_modes[mode]
        rec, curLabel, cube2, com2, M2 = hd.aug_record(0, name, off, rot, sc, com, cube, M, gt3Dcrop)
        imgD = run_records(numpy.ascontiguousarray(img, dtype='float32')[None], numpy.array([rec]))[0]
        rot_out = numpy.mod(rot, 360) if (name == 'rot' and not numpy.allclose(rot, 0.)) else (rot if name == 'rot' else 0.)
        return imgD, None, curLabel, numpy.asarray(cube2), com2, M2, rot_out
