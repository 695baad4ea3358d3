"""NetBase (reference: src/net/netbase.py:52-477): layer-list container, params/weights views,
batched ``computeOutput`` with last-batch padding, deterministic switch, pickle save/load with the
reference's ``'{layerNum}-values'`` schema.  Forward passes run through ``dpp_b200.engine``."""
import difflib
import gzip
import pickle
import time
import numpy

from net.sym import tensor4
from net.convpoollayer import ConvPoolLayer, ConvPoolLayerParams  # noqa: F401 (class-name lookup)
from net.convlayer import ConvLayer, ConvLayerParams  # noqa: F401
from net.hiddenlayer import HiddenLayer, HiddenLayerParams  # noqa: F401
from net.dropoutlayer import DropoutLayer, DropoutLayerParams  # noqa: F401
from net.batchnormlayer import BatchNormLayer, BatchNormLayerParams  # noqa: F401
from net.nonlinearitylayer import NonlinearityLayer, NonlinearityLayerParams  # noqa: F401


class NetBaseParams(object):
    def __init__(self):
        self.numInputs = 1
        self.numOutputs = 1
        self.layers = []
        self.inputDim = None
        self.outputDim = None
        self.loadFile = None

    def getMemoryRequirement(self):
        return sum(l.getMemoryRequirement() for l in self.layers)


def _unique(prms):
    # netbase.py:150-151: dict keyed by auto_name; py3 dicts keep insertion (= layer) order
    return list(dict((obj.auto_name, obj) for obj in prms).values())


class NetBase(object):
    def __init__(self, rng, inputVar, cfgParams, twin=None):
        self._params_filter = []
        self._weights_filter = []
        self.inputVar = inputVar
        self.cfgParams = cfgParams
        self.rng = rng
        self.layers = []
        for i, layerParam in enumerate(cfgParams.layers):
            if i == 0:
                inp = inputVar
            else:
                prev = self.layers[-1]
                # netbase.py:103-110: flatten conv->hidden, reshape hidden->conv
                if len(prev.cfgParams.outputDim) == 4 and len(layerParam.inputDim) == 2:
                    inp = prev.output.flatten(2)
                    inp.name = "input_layer_{}".format(i)
                elif len(layerParam.inputDim) == 4 and len(prev.cfgParams.outputDim) == 2:
                    inp = prev.output.reshape(layerParam.inputDim, ndim=4)
                    inp.name = "input_layer_{}".format(i)
                else:
                    inp = prev.output
            constructor = globals()[layerParam.__class__.__name__[:-6]]
            self.layers.append(constructor(rng, inputVar=inp, cfgParams=layerParam,
                                           copyLayer=None if (twin is None) else twin.layers[i], layerNum=i))
        self.output = self.layers[-1].output
        self.load(self.cfgParams.loadFile)

    def __str__(self):
        cfg = "Network configuration:\n"
        for i, l in enumerate(self.layers):
            cfg += "Layer {}: {} with {} \n".format(i, l.__class__.__name__, l)
        return cfg

    # -- parameter views (netbase.py:141-203) --------------------------------------------
    @property
    def all_params(self):
        return _unique([p for l in self.layers for p in l.params])

    @property
    def params(self):
        if not hasattr(self, '_params_filter'):
            self._params_filter = []
        blocked = [an.auto_name for an in self._params_filter]
        return _unique([p for l in self.layers for p in l.params if p.auto_name not in blocked])

    @property
    def params_filter(self):
        return self._params_filter

    @params_filter.setter
    def params_filter(self, bl):
        names = [p.auto_name for l in self.layers for p in l.params]
        for b in bl:
            if b.auto_name not in names:
                raise UserWarning("Param {} not in model!".format(b))
        self._params_filter = bl

    @property
    def all_weights(self):
        return _unique([p for l in self.layers for p in l.weights])

    @property
    def weights(self):
        if not hasattr(self, '_weights_filter'):
            self._weights_filter = []
        blocked = [an.auto_name for an in self._weights_filter]
        return _unique([p for l in self.layers for p in l.weights if p.name not in blocked])

    @property
    def weights_filter(self):
        return self._weights_filter

    @weights_filter.setter
    def weights_filter(self, bl):
        names = [p.auto_name for l in self.layers for p in l.weights]
        for b in bl:
            if b.auto_name not in names:
                raise UserWarning("Weight {} not in model!".format(b))
        self._weights_filter = bl

    # -- engine ---------------------------------------------------------------------------
    def _engine(self):
        """Build (once) the device executor for the current ``self.output`` graph.  It is
        rebuilt if layers were appended (main_nyu_posereg_embedding.py:148-158 appends the PCA
        prior layer after training)."""
        from dpp_b200.engine import Engine
        eng = getattr(self, '_eng', None)
        if eng is None or eng.output_sym is not self.output:
            if eng is not None:
                eng.release()
            eng = Engine(self, **getattr(self, '_engine_opts', {}))     # a data-parallel trainer sets {'batch': B / world}
            self._eng = eng
        return eng

    def computeOutput(self, inputs, timeit=False):
        """netbase.py:217-316: pad the sample count to a multiple of batch_size (last batch padded
        by repeating the last sample), run batch by batch, return the first nSamp rows."""
        if not isinstance(inputs, list):
            inputs = [inputs]
        assert all(i.shape[0] == inputs[0].shape[0] for i in inputs[1:])
        if not self.isDeterministic():
            print("WARNING: network is probabilistic for testing, DISABLING")
            self.setDeterministic()
        batch_size = self.cfgParams.batch_size
        nSamp = inputs[0].shape[0]
        padSize = int(batch_size * numpy.ceil(nSamp / float(batch_size)))
        eng = self._engine()
        multi_out = isinstance(self.output, list)
        if multi_out:
            raise NotImplementedError("networks with several outputs (none of the reference's nets has more than one)")
        outdims = [self.cfgParams.outputDim]
        out = [numpy.zeros((padSize,) + tuple(od[1:]), dtype='float32') for od in outdims]
        # the device executor may hold a share of the batch only (data-parallel training): walk the padded sample
        # range in ITS batch size - in deterministic mode every sample is independent, so the result is the same
        dev_batch = eng.B
        assert padSize % dev_batch == 0
        start = time.time()
        for i in range(padSize // dev_batch):
            batch = []
            for k in range(len(inputs)):
                chunk = inputs[k][i * dev_batch:(i + 1) * dev_batch]
                if chunk.shape[0] < dev_batch:
                    pad = numpy.zeros((dev_batch,) + chunk.shape[1:], dtype=inputs[k].dtype)
                    pad[0:chunk.shape[0]] = chunk
                    pad[chunk.shape[0]:] = inputs[k][-1]
                    chunk = pad
                batch.append(numpy.ascontiguousarray(chunk, dtype='float32'))
            o = eng.forward_host(batch)
            out[0][i * dev_batch:(i + 1) * dev_batch] = o.reshape((dev_batch,) + tuple(outdims[0][1:]))
        end = time.time()
        if timeit:
            print("{} in {}s, {}ms per frame".format(padSize, end - start, (end - start) * 1000. / padSize))
        return out[0][0:nSamp]

    # -- train/test switch (netbase.py:318-358) ------------------------------------------
    def unsetDeterministic(self):
        for layer in self.layers:
            if isinstance(layer, (DropoutLayer, BatchNormLayer)):
                layer.unsetDeterministic()

    def setDeterministic(self):
        for layer in self.layers:
            if isinstance(layer, (DropoutLayer, BatchNormLayer)):
                layer.setDeterministic()

    def isDeterministic(self):
        for layer in self.layers:
            if isinstance(layer, (DropoutLayer, BatchNormLayer)):
                if not layer.isDeterministic():
                    return False
        return True

    def hasDropout(self):
        return any(isinstance(layer, DropoutLayer) for layer in self.layers)

    # -- weight values (netbase.py:360-403) ------------------------------------------------
    @property
    def weightVals(self):
        return self.recGetWeightVals(self.all_params)

    @weightVals.setter
    def weightVals(self, value):
        self.recSetWeightVals(self.all_params, value)

    def recSetWeightVals(self, param, value):
        if isinstance(value, list):
            assert isinstance(param, list), "tried to assign a list of weights to params, which is not a list"
            assert len(param) == len(value), "tried to assign unequal list of weights"
            for i in range(len(value)):
                self.recSetWeightVals(param[i], value[i])
        else:
            param.set_value(value)

    def recGetWeightVals(self, param):
        if isinstance(param, list):
            return [self.recGetWeightVals(p) for p in param]
        return param.get_value()

    # -- checkpoints (netbase.py:405-477) --------------------------------------------------
    def save(self, filename):
        try:                                    # data-parallel job: the replicas are identical, rank 0 writes
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
                return
        except ImportError:
            pass
        state = dict([('class', self.__class__.__name__), ('network', self.__str__())])
        for layer in self.layers:
            key = '{}-values'.format(layer.layerNum)
            state[key] = [p.get_value() for p in layer.params]
            state[key].extend([p.get_value() for p in layer.params_nontrained])
        opener = gzip.open if filename.lower().endswith('.gz') else open
        with opener(filename, 'wb') as handle:
            pickle.dump(state, handle, 2)       # protocol 2: readable from the reference's py2 cPickle
        print('Saved model parameter to {}'.format(filename))

    def load(self, filename, raise_on_error=True):
        if filename is None:
            return
        print('Loading model parameters from {}'.format(filename))
        opener = gzip.open if filename.lower().endswith('.gz') else open
        with opener(filename, 'rb') as handle:
            saved = pickle.load(handle, encoding='latin1')   # reads py2 pickles of the reference
        if saved['network'] != self.__str__():
            print("Possibly not matching network configuration!")
            differences = list(difflib.Differ().compare(saved['network'].splitlines(), self.__str__().splitlines()))
            print("Differences are:")
            print("\n".join(differences))
        for layer in self.layers:
            key = '{}-values'.format(layer.layerNum)
            targets = layer.params + layer.params_nontrained
            if key not in saved:
                if raise_on_error:
                    raise ImportError("{} not in saved variables!".format(key))
                print("WARNING: {} not in saved variables!".format(key))
                continue
            if len(targets) != len(saved[key]):
                print("Warning: Layer parameters for layer {} do not match. Trying to fit on shape!".format(layer.layerNum))
                n_assigned = 0
                for p in targets:
                    for v in saved[key]:
                        if p.get_value().shape == v.shape:
                            p.set_value(v)
                            n_assigned += 1
                if n_assigned != len(targets):
                    if raise_on_error:
                        raise ImportError("Could not load all necessary variables!")
                    print("WARNING: Could not load all necessary variables!")
                else:
                    print("Found fitting parameters!")
            else:
                for p, v in zip(targets, saved[key]):
                    if p.get_value().shape == v.shape:
                        p.set_value(v)
                    elif raise_on_error:
                        raise ImportError("Skipping parameter for {}! Shape {} does not fit {}.".format(
                            p.name, p.get_value().shape, v.shape))
                    else:
                        print("WARNING: Skipping parameter for {}! Shape {} does not fit {}.".format(
                            p.name, p.get_value().shape, v.shape))
        print('Done')
