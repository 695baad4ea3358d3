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
    def __init__(self, net):
        cfg = net.cfgParams
        self.dev = torch.device('cpu')
        self.B = cfg.batch_size
        dims = cfg.inputDim if isinstance(cfg.inputDim, list) else [cfg.inputDim]
        self.t_ins = [FakeT(d) for d in dims]
        self.output_sym = net.output
        if len(dims) == 3:
            self.onet = ON.build_scalenet(np.random.RandomState(23455), type=1, batchSize=self.B, numJoints=1, nDims=3)
        elif type(net).__name__ == 'PoseRegNet':
            self.onet = ON.build_poseregnet(np.random.RandomState(23455), type=0, batchSize=self.B,
                                            numJoints=cfg.numJoints, nDims=cfg.nDims)
        else:
            self.onet = ON.build_resnet(np.random.RandomState(23455), type=1, batchSize=self.B, numJoints=cfg.numJoints, nDims=3)
    def _x(self): return [t.buf.permute(0, 3, 1, 2).contiguous() for t in self.t_ins]
    def forward_device(self, deterministic=True):
        with torch.no_grad():
            xs = self._x()
            o, _ = self.onet.forward(xs if len(xs) > 1 else xs[0], deterministic=True)
        return o
    def forward_host(self, batch, deterministic=True):
        for t, b in zip(self.t_ins, batch): t.buf.copy_(torch.from_numpy(b).permute(0, 2, 3, 1))
        return self.forward_device().numpy()
    def release(self): pass
from net import netbase
def _engine(self):
    eng = getattr(self, '_eng', None)
    if eng is None:
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
