"""Host-side dry run of the inference cascade WITHOUT a GPU (run by tests/test_host_cascade.py in a subprocess).

TEST INFRASTRUCTURE ONLY.  It checks the Python plumbing of the product path - dpp_b200/cascade.py (records, padding of
the last batch, pointer hand-off to the C ABI), util/realtimehandposepipeline.py and the single-frame HandDetector
methods - by running tests/test_gpu_cascade.py's test bodies with
  * torch.cuda faked (tensors stay on the CPU),
  * dpp_recrop_fwd / dpp_joint_errors / dpp_sample_poses emulated from the raw pointers by tests/recrop_model.py (a NumPy transcription of
    k_recrop), tests/poses_model.py (k_sample_poses) and NumPy,
  * the device engines replaced by the oracle nets.
Nothing here is reachable from the product: the real entry points raise DppError without CUDA.  The kernels themselves
are only ever validated on the GPU (pytest -m gpu)."""
import os
import sys
import ctypes as C
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, 'deep-prior-pp_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np, torch
import recrop_model
from dpp_b200 import cascade as PC
from dpp_b200.lib import CROP_REC_DTYPE

# ---- fake torch.cuda
torch.cuda.is_available = lambda: True
torch.Tensor.cuda = lambda self, *a, **k: self
torch.Tensor.is_cuda = property(lambda self: True)
class _S: cuda_stream = 0
torch.cuda.current_stream = lambda *a, **k: _S()
torch.cuda.set_device = lambda *a, **k: None
_full, _empty = torch.full, torch.empty
def _nodev(f):
    def g(*a, **k):
        k.pop('device', None); return f(*a, **k)
    return g
torch.full, torch.empty = _nodev(_full), _nodev(_empty)
_DevT0 = torch.device
_from = torch.Tensor.to
def _to(self, *a, **k):
    k.pop('non_blocking', None)
    a = tuple(x for x in a if not isinstance(x, (_DevT0, str)))
    return _from(self, *a, **k) if (a or k) else self
torch.Tensor.to = _to

def _arr(ptr, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    buf = (C.c_char * n).from_address(ptr.value if isinstance(ptr, C.c_void_p) else ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)

class FakeLib:
    def dpp_recrop_fwd(self, frames, recs, o0, o1, o2, n, Hf, Wf, H, W, st):
        assert W % 32 == 0 and H % 8 == 0
        if n == 0: return 0
        rec = _arr(recs, (n,), np.dtype(CROP_REC_DTYPE))
        F = int(rec['src_index'].max()) + 1
        fr = _arr(frames, (F, Hf, Wf), np.float32)
        x0, x1, x2 = recrop_model.run(fr, rec, H, W, centre_crops=True)
        _arr(o0, (n, H, W), np.float32)[...] = x0
        if o1 is not None: _arr(o1, (n, H//2, W//2), np.float32)[...] = x1
        if o2 is not None: _arr(o2, (n, H//4, W//4), np.float32)[...] = x2
        return 0
    def dpp_joint_errors(self, p, g, err, fm, fx, n, J, st):
        P, G = _arr(p, (n, J, 3), np.float32), _arr(g, (n, J, 3), np.float32)
        e = np.sqrt(np.square(G - P).sum(axis=2))
        _arr(err, (n, J), np.float32)[...] = e
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            _arr(fm, (n,), np.float32)[...] = np.nanmean(e, axis=1)
            _arr(fx, (n,), np.float32)[...] = np.nanmax(e, axis=1)
        return 0
    def dpp_augment_fwd(self, crops, recs, out, n_out, H, W, st):
        import augment_model
        from dpp_b200.lib import AUG_REC_DTYPE
        if n_out == 0:
            return 0
        rec = _arr(recs, (n_out,), np.dtype(AUG_REC_DTYPE))
        src = _arr(crops, (int(rec['src_index'].max()) + 1, H, W), np.float32)
        _arr(out, (n_out, H, W), np.float32)[...] = augment_model.run(src, rec, H, W)
        return 0

    def dpp_sample_poses(self, bp, bc, bcu, mode, ridx, off, sc, cs, fx, fy, ux, uy, flip, o_p, o_c, o_cu, n, J, st):
        import poses_model
        md, ri = _arr(mode, (n,), np.int32), _arr(ridx, (n,), np.int32)
        nb = int(ri.max()) + 1
        res = poses_model.run(_arr(bp, (nb, J, 3), np.float32), _arr(bc, (nb, 3), np.float32), _arr(bcu, (nb, 3), np.float32),
                              md, ri, _arr(off, (n, 3), np.float64), _arr(sc, (n,), np.float64),
                              _arr(cs, (n, 2), np.float64), (fx, fy, ux, uy, flip))
        _arr(o_p, (n, J, 3), np.float32)[...] = res[0]
        _arr(o_c, (n, 3), np.float32)[...] = res[1]
        _arr(o_cu, (n, 3), np.float32)[...] = res[2]
        return 0
PC.lib = FakeLib()
sys.modules['dpp_b200.lib'].lib = PC.lib
import dpp_b200.augment as _AUG
_AUG.lib = PC.lib
torch.cuda.current_device = lambda: 0
_DevT = torch.device


class _CpuDevice(object):          # torch.device('cuda', 0) -> the CPU device; isinstance checks keep working
    def __new__(cls, *a, **k):
        return _DevT('cpu')


torch.device = _CpuDevice

# ---- fake engines: oracle nets
from oracle import nets as ON
class FakeT:
    def __init__(self, shape): self.shape = tuple(shape); self.buf = torch.zeros((shape[0], shape[2], shape[3], shape[1]))
class FakeEngine:
    """forward-only stand-in for dpp_b200.engine.Engine: the oracle net of the same architecture, carrying the PRODUCT
    net's current parameter values (and its appended PCA prior layer, main_nyu_posereg_embedding.py:148-158)"""

    def __init__(self, net):
        cfg = net.cfgParams
        self.dev = torch.device('cpu')
        self.B = cfg.batch_size
        dims = cfg.inputDim if isinstance(cfg.inputDim, list) else [cfg.inputDim]
        self.t_ins = [FakeT(d) for d in dims]
        self.output_sym = net.output
        self.ops = [dict(kind="conv"), dict(kind="fc"), dict(kind="concat", srcs=[1, 2, 3])]
        kind = type(net).__name__
        nd = net.layers[-1].cfgParams.outputDim[1]
        n_own = len(cfg.layers) if getattr(cfg, 'layers', None) else None      # layers the net class itself built
        extra = [] if n_own is None else net.layers[n_own:]
        if kind == 'ScaleNet':
            self.onet = ON.build_scalenet(np.random.RandomState(23455), type=cfg_type(net), batchSize=self.B, numJoints=1, nDims=3)
        elif kind == 'PoseRegNet':
            base_out = net.layers[n_own - 1].cfgParams.outputDim[1] if extra else nd
            self.onet = ON.build_poseregnet(np.random.RandomState(23455), type=cfg_type(net), batchSize=self.B,
                                            numJoints=1, nDims=base_out)
        else:
            extra = []
            self.onet = ON.build_resnet(np.random.RandomState(23455), type=cfg_type(net), batchSize=self.B,
                                        numJoints=1, nDims=nd)
        own_params = [p for l in (net.layers[:n_own] if n_own else net.layers) for p in l.params]
        assert len(own_params) == len(self.onet.params)
        with torch.no_grad():
            for po, pp in zip(self.onet.params, own_params):
                po.copy_(torch.from_numpy(np.asarray(pp.get_value(), np.float64)).to(po.dtype).reshape(po.shape))
        for l in extra:                                  # the PCA prior layer(s) appended by the entry script
            ON.append_pca_layer(self.onet, l.W.get_value(), l.b.get_value())
        self.net, self.own_params = net, own_params
        self.t_in = self.t_ins[0]
        self.y_in, self.adam = None, None
        self.mask_gen = torch.Generator().manual_seed(1)

    def _alloc_training(self):
        if self.y_in is None:
            self.y_in = torch.zeros((self.B, int(self.net.layers[-1].cfgParams.outputDim[1])), dtype=torch.float32)
            self.adam = ON.Adam(self.onet.params)

    def train_step(self, lr=None, use_graph=True):
        """one oracle train step (batch statistics, dropout masks, ADAM, BN EMA) on the staged batch; the new weights
        are written back to the product's variables like the device arena would hold them"""
        self._alloc_training()
        masks = [(torch.rand((self.B, int(l.cfgParams.outputDim[1])), generator=self.mask_gen) < l.prob_keep).double()
                 for l in self.net.layers if type(l).__name__ == 'DropoutLayer'] or None
        cfg = self.net.cfgParams
        cost, _, _ = ON.train_step(self.onet, self.adam, self._x()[0], self.y_in.clone(), float(lr), cfg.numJoints,
                                   cfg.nDims, masks=masks)
        for po, pp in zip(self.onet.params, self.own_params):
            pp.set_value(po.detach().numpy().astype(np.float32).reshape(pp.get_value().shape))
        return torch.tensor([cost], dtype=torch.float32)

    def _x(self):
        return [t.buf.permute(0, 3, 1, 2).contiguous() for t in self.t_ins]

    def forward_device(self, deterministic=True):
        with torch.no_grad():
            xs = self._x()
            o, _ = self.onet.forward(xs if len(xs) > 1 else xs[0], deterministic=deterministic)
        return o.float()                                  # the device buffers are float32

    def forward_host(self, batch, deterministic=True):
        for t, b in zip(self.t_ins, batch):
            t.buf.copy_(torch.from_numpy(b).permute(0, 2, 3, 1))
        return self.forward_device().numpy()

    def release(self):
        pass


def cfg_type(net):
    """ResNetParams keeps its type; the PoseRegNet / ScaleNet parameter classes do not (like the reference's):
    type 0 of PoseRegNet has 8 layers, type 11 has 9 (tests/golden/reference_nets.json); ScaleNet is type 1 only."""
    kind = type(net).__name__
    if kind == 'ResNet':
        return net.cfgParams.type
    if kind == 'ScaleNet':
        return 1
    return 0 if len(net.cfgParams.layers) == 8 else 11


from net import netbase
def _engine(self):
    eng = getattr(self, '_eng', None)
    if eng is None or eng.output_sym is not self.output:      # rebuilt when layers were appended, like NetBase._engine
        eng = FakeEngine(self); self._eng = eng
    return eng
netbase.NetBase._engine = _engine

import test_gpu_cascade as T
for name in ['NYU', 'ICVL', 'MSRA15']:
    T.test_recrop_kernel_bit_exact(name); print('recrop', name, 'ok')
T.test_joint_errors_match_reference_formulas(); print('joint errors ok')
T.test_handpose_evaluation_metrics(); print('evaluation ok')
for name in ['NYU', 'MSRA15']:
    T.test_dataset_stack_on_device(name)
print('dataset ok')
T.test_cascade_matches_oracle(); print('cascade ok')
import test_gpu_poses as TP
for name in ['NYU', 'ICVL', 'MSRA15']:
    TP.test_sample_random_poses_matches_oracle(name); TP.test_sample_random_poses_matches_reference_fixture(name)
print('poses ok')


# ------------------------------------------------------------------------------------------------------------------
# the reference's OWN entry scripts against the product package (build container only: needs /root/reference)
# ------------------------------------------------------------------------------------------------------------------
def run_reference_entry_script(script, workdir):
    """Executes /root/reference/src/<script> - py2 -> py3 pass in memory (oracle/ref_harness.py), nothing else changed -
    with ``deep-prior-pp_b200/`` as its ``src/``.  Everything the script does on the host runs for real: sequences from
    the importers (synthetic backend), Dataset stacks, side arrays, HandDetector.sampleRandomPoses + PCA, network and
    trainer construction, setData / addStaticData / addManagedData / compileFunctions, later save(), the PCA prior
    layer, computeOutput() and the HandposeEvaluation metrics.  Replaced: the device (see the top of this file),
    the number of epochs (2 instead of 100; the loop itself - record workers, device augmentation from the
    original crops, minibatches, validation observers, snapshots, best-parameter restore - runs with the oracle as the
    arithmetic) and matplotlib.  The run ends where the script asks for other methods' published result files (``loadBaseline``)."""
    import ast
    import pickle
    import types
    from unittest import mock
    from oracle import ref_harness as RH
    import trainer.poseregnettrainer as TP
    plt = mock.MagicMock()
    mpl = types.ModuleType('matplotlib')
    mpl.use = lambda *a, **k: None
    mpl.pyplot = plt
    stubs = {'matplotlib': mpl, 'matplotlib.pyplot': plt, 'cPickle': pickle}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    calls = []

    real_train = TP.PoseRegNetTrainer.train

    def fake_train(self, n_epochs=50, storeFilters=False):
        calls.append(n_epochs)                   # the scripts ask for 100 epochs; two are enough to walk the loop
        self.verbose = False
        return real_train(self, n_epochs=2, storeFilters=storeFilters)
    TP.PoseRegNetTrainer.train = fake_train
    cwd = os.getcwd()
    os.chdir(workdir)
    os.makedirs('eval', exist_ok=True)
    text = RH.py3_source(os.path.join(RH.REF_SRC, script))
    tree = ast.fix_missing_locations(RH._Div().visit(ast.parse(text, filename=script)))
    g = {'__name__': '__main__', '__file__': script, '_py2div': RH._py2div, 'xrange': range}
    ended = None
    try:
        exec(compile(tree, script, 'exec'), g)
        ended = 'completed'
    except AttributeError as e:
        if 'loadBaseline' not in str(e):
            raise
        ended = 'loadBaseline'
    finally:
        os.chdir(cwd)
        TP.PoseRegNetTrainer.train = real_train
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert calls, "the script never reached train()"
    hpe = g.get('hpe')
    return ended, calls, g, hpe


if __name__ == '__main__' or True:
    from oracle import ref_harness as _RH
    if _RH.available():
        import tempfile
        # 130 frames: two minibatches (the second padded with random real samples) and one full validation batch;
        # 40 frames: fewer validation samples than a batch (the observers average over zero batches, like the reference)
        for script, frames in (('main_nyu_posereg_embedding.py', '130'), ('main_icvl_posereg_embedding.py', '40')):
            os.environ['DPP_SYNTH_FRAMES'] = frames
            os.environ['DPP_SYNTHETIC'] = '1'          # the reference's scripts name ../data/<set>/: explicit opt-in to synthetic frames
            with tempfile.TemporaryDirectory() as d:
                ended, calls, g, hpe = run_reference_entry_script(script, d)
            joints = g['joints']
            print(script, 'ran to', ended, '| train(n_epochs=%d)' % calls[0], '| joints', joints.shape,
                  '| mean error %.1f mm' % hpe.getMeanError())
            assert ended == 'loadBaseline' and joints.shape[1:] == (g['train_gt3D'].shape[1], 3)
            assert g['poseNet'].cfgParams.numJoints == g['train_gt3D'].shape[1] and np.isfinite(hpe.getMeanError())
        print('entry scripts ok')
    else:
        print('entry scripts skipped (no /root/reference)')
