"""Generates tests/golden/reference_pins.npz by EXECUTING THE REFERENCE'S OWN SOURCES (oracle/ref_harness.py: read from
/root/reference, py2 -> py3 pass in memory, nothing copied) in the build container: Python 3.12, NumPy 2.3, cv2 4.13.
Run: ``python tests/golden/make_reference_vectors.py``.  tests/test_reference_pins.py checks the oracle against it.

Functions executed (reference file:line):
  trainer/nettrainer.py:919-997        NetTrainer.augmentCrop  (with HandDetector.moveCoM / rotateHand / scaleHand /
                                       recropHand / comToTransform / comToBounds, util/handdetector.py:204-258,678-803)
  util/handdetector.py:511-533,634-676 track (doHandSize=False) + refineCoM with a recording stub net
  util/realtimehandposepipeline.py:296-368  detect (tracking branch: track + cropArea3D + normalisation), estimatePose
  util/handdetector.py:805-909         sampleRandomPoses
  data/importers.py                    NYU / ICVL / MSRA15 jointImgTo3D, joint3DToImg
  net/*.py                             ResNet / PoseRegNet / ScaleNet constructors -> tests/golden/reference_nets.json
  net/*.py, trainer/poseregnettrainer.py:70-111 (cost, T.grad), trainer/optimizer.py:58-90 (ADAM) EVALUATED with
                                       oracle/eager_theano.py in place of Theano -> tests/golden/reference_net_eval.npz
NumPy-generation caveat: see oracle/ref_harness.py - float32-scalar arithmetic is float32 under NumPy 2 where the
reference-era NumPy 1.x used float64, so float outputs can differ from the oracle (which restates NumPy 1.x) in the
last bits; every integer / index result is expected to agree exactly."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'deep-prior-pp_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

DATASETS = {'NYU': ('NYUImporter', 'NYU_CAM', 588., 587.), 'ICVL': ('ICVLImporter', 'ICVL_CAM', 241.42, 241.42),
            'MSRA15': ('MSRA15Importer', 'MSRA_CAM', 241.42, 241.42)}
AUG_MODES = ['com', 'rot', 'sc', 'none']
POSE_MODES = ['com', 'rot', 'sc', 'none', 'rot+com', 'rot+com+sc']


class _Cfg(object):
    numInputs = 3


class RecordingNet(object):
    """stands in for the refinement net: records what refineCoM feeds it, answers with a fixed linear map"""
    cfgParams = _Cfg()

    def __init__(self, fn):
        self.fn, self.seen = fn, None

    def computeOutput(self, xs):
        self.seen = [np.array(x, copy=True) for x in xs]
        return self.fn(xs)


def augment_vectors(ref, name, n, seed):
    from oracle import ref_harness as RH, augment as OA
    from data import synthetic
    cls, camname, _, _ = DATASETS[name]
    rdi = getattr(ref['importers'], cls)('/nonexistent/')
    cam = OA.Camera(**getattr(OA, camname))
    augmentCrop = RH.reference_function('trainer/nettrainer.py', 'augmentCrop', {'numpy': np})
    ds = synthetic.generate(name, n, seed=seed)
    rhd = ref['handdetector'].HandDetector(np.zeros((128, 128), np.float32) + 1., abs(rdi.fx), abs(rdi.fy), importer=rdi)

    class Self(object):
        rng = np.random.RandomState(seed + 1)
    out = dict(x=ds['x'][:, 0], gt3Dcrop=ds['gt3Dcrop'], cube=ds['cube'], M=ds['M'], rng_seed=np.int64(seed + 1))
    coms, res = [], []
    for i in range(n):
        com = cam.joint3DToImg(ds['com3D'][i])       # NumPy-1.x value (the reference under NumPy 2 is 1-2 ulp off)
        coms.append(com)
        res.append(augmentCrop(Self, ds['x'][i, 0].copy(), ds['gt3Dcrop'][i].copy(), com.copy(), ds['cube'][i].copy(),
                               ds['M'][i].copy(), AUG_MODES, rhd))
    out.update(com=np.stack(coms), out_img=np.stack([r[0] for r in res]), out_label=np.stack([r[2] for r in res]),
               out_cube=np.stack([np.asarray(r[3], np.float64) for r in res]), out_com=np.stack([r[4] for r in res]),
               out_M=np.stack([np.asarray(r[5], np.float64) for r in res]), out_rot=np.array([r[6] for r in res]))
    return {'augment_%s_%s' % (name, k): v for k, v in out.items()}


class _Val(object):
    def __init__(self, value):
        self.value = value


def reference_pipeline(ref, rdi, fx, fy, cube, comref_net, pose_net=None, right_hand=False):
    """A stand-in ``self`` carrying the reference's OWN RealtimeHandposePipeline.detect / estimatePose
    (util/realtimehandposepipeline.py:296-368, extracted and executed by oracle/ref_harness.py) in the tracking state
    the realtime loop is in after initialisation."""
    import types
    from oracle import ref_harness as RH
    g = {'numpy': np, 'HandDetector': ref['handdetector'].HandDetector}
    p = types.SimpleNamespace(STATE_IDLE=0, STATE_INIT=1, STATE_RUN=2, HAND_LEFT=0, HAND_RIGHT=1,
                              sync={'config': {'fx': fx, 'fy': fy, 'cube': tuple(cube)}}, importer=rdi,
                              comrefNet=comref_net, poseNet=pose_net, state=_Val(2), tracking=_Val(True),
                              hand=_Val(1 if right_hand else 0), lastcom=(0, 0, 0), handsizes=[], numinitframes=50,
                              verbose=False)
    p.detect = types.MethodType(RH.reference_function('util/realtimehandposepipeline.py', 'detect', g), p)
    p.estimatePose = types.MethodType(RH.reference_function('util/realtimehandposepipeline.py', 'estimatePose', g), p)
    return p


class _PoseNetStub(object):
    """what detect() asks the pose net for: its input size"""
    def __init__(self, fn=None):
        import types
        dim = (1, 1, 128, 128)
        self.cfgParams = types.SimpleNamespace(inputDim=dim)
        self.layers = [types.SimpleNamespace(cfgParams=types.SimpleNamespace(inputDim=dim))]
        self.fn, self.seen = fn, None

    def computeOutput(self, x):
        self.seen = np.array(x, copy=True)
        return self.fn(x)


def cascade_vectors(ref, name, n, seed):
    from data import synthetic
    from test_oracle_cascade import _tiny_fns
    cls, _, fx, fy = DATASETS[name]
    rdi = getattr(ref['importers'], cls)('/nonexistent/')
    fr = synthetic.generate_frames(name, n, seed=seed, edge_fraction=0.5)
    cube = fr['cube']
    lastcom = fr['lastcom'].astype(np.float32).astype(np.float64)    # float64 dtype, float32-representable values:
    refine_fn, _ = _tiny_fns(77)                                     # both NumPy generations then agree bit for bit
    keys = ('x0', 'x1', 'x2', 'loc', 'crop_raw', 'crop', 'M', 'com3D')
    acc = {k: [] for k in keys}
    for i in range(n):
        net = RecordingNet(refine_fn)
        pipe = reference_pipeline(ref, rdi, fx, fy, cube, net, _PoseNetStub())
        pipe.lastcom = lastcom[i].copy()
        crop, M, com3D = pipe.detect(fr['frames'][i])                # the reference's own detect(): track + cropArea3D
        loc = pipe.lastcom                                           # + its normalisation (:327-332)
        hd = ref['handdetector'].HandDetector(fr['frames'][i], fx, fy, importer=rdi, refineNet=None)
        raw, _, _ = hd.cropArea3D(com=loc, size=cube, dsize=(128, 128))
        for k, v in zip(keys, (net.seen[0][0, 0], net.seen[1][0, 0], net.seen[2][0, 0], loc, raw, crop, M, com3D)):
            acc[k].append(np.array(v, copy=True))
    out = dict(frames=fr['frames'], lastcom=lastcom, cube=np.array(cube), fxfy=np.array([fx, fy]), fn_seed=np.int64(77))
    out.update({k: np.stack(v) for k, v in acc.items()})
    return {'cascade_%s_%s' % (name, k): v for k, v in out.items()}


def pose_vectors(ref, name, n, seed):
    from data import synthetic
    cls = DATASETS[name][0]
    rdi = getattr(ref['importers'], cls)('/nonexistent/')
    ds = synthetic.generate(name, 8, seed=seed)
    r = ref['handdetector'].HandDetector.sampleRandomPoses(rdi, np.random.RandomState(seed + 1), ds['gt3Dcrop'],
                                                           ds['com3D'], ds['cube'], n, POSE_MODES, retall=True)
    out = dict(base_poses=ds['gt3Dcrop'], base_com=ds['com3D'], base_cube=ds['cube'], rng_seed=np.int64(seed + 1),
               out_poses=r[0], out_com=r[1], out_cube=r[2], out_rot=r[3])
    return {'poses_%s_%s' % (name, k): v for k, v in out.items()}


def geometry_vectors(ref, name, n, seed):
    cls, _, fx, fy = DATASETS[name]
    rdi = getattr(ref['importers'], cls)('/nonexistent/')
    rng = np.random.RandomState(seed)
    W, H = rdi.depth_map_size
    coms = np.stack([rng.uniform(20, W - 20, n), rng.uniform(20, H - 20, n), rng.uniform(350, 950, n)], axis=1)
    hd = ref['handdetector'].HandDetector(np.zeros((H, W), np.float32) + 1., fx, fy, importer=rdi)
    cube = (250, 250, 250)
    bounds = np.array([hd.comToBounds(c, cube) for c in coms], np.float64)
    trans = np.stack([hd.comToTransform(c, cube, (128, 128)) for c in coms])
    p3 = np.stack([rdi.jointImgTo3D(c) for c in coms])
    pts = np.stack([rng.uniform(-200, 200, n), rng.uniform(-200, 200, n), rng.uniform(350, 950, n)], axis=1)
    p2 = np.stack([rdi.joint3DToImg(p) for p in pts])
    out = dict(coms=coms, cube=np.array(cube), bounds=bounds, transform=trans, img_to_3d=p3, pts=pts, to_img=p2)
    return {'geometry_%s_%s' % (name, k): v for k, v in out.items()}


NET_CASES = [
    ('ResNet', dict(type=0, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30)),
    ('ResNet', dict(type=1, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=14, nDims=3)),
    ('ResNet', dict(type=4, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=14, nDims=3)),
    ('PoseRegNet', dict(type=0, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30)),
    ('ScaleNet', dict(type=1, nChan=1, wIn=128, hIn=128, batchSize=2, resizeFactor=2, numJoints=1, nDims=3)),
]


def net_descriptions():
    """net/resnet.py, net/poseregnet.py, net/scalenet.py constructors (and every layer class under net/) executed with
    an inert theano stand-in: layer lists, dimensions, parameter names / order and the initial weights' sha1."""
    import json
    from oracle import ref_harness as RH
    out = [RH.describe_reference_net(kind, **cfg) for kind, cfg in NET_CASES]
    path = os.path.join(HERE, 'reference_nets.json')
    with open(path, 'w') as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


def _stats(a):
    a = np.asarray(a, np.float64)
    return np.array([a.sum(), np.abs(a).sum(), np.abs(a).max(), float(a.ravel()[0]), float(a.ravel()[-1])])


def net_eval_inputs(kind, cfg, seed, train):
    """the inputs of one case, regenerated from the seed by the tests (not stored)"""
    B = cfg['batchSize']
    rng = np.random.RandomState(seed)
    x0 = rng.uniform(-1, 1, (B, 1, 128, 128)).astype(np.float32)
    xs = [x0] if kind != 'ScaleNet' else [x0, np.ascontiguousarray(x0[:, :, 32:96, 32:96]),
                                           np.ascontiguousarray(x0[:, :, 48:80, 48:80])]
    y = rng.randn(B, cfg['numJoints'] * cfg['nDims']).astype(np.float32) if train else None
    return xs, y


def net_eval_case(kind, cfg, seed, train):
    """One evaluation of the reference's own network / trainer code (oracle/eager_theano.py stands in for Theano):
    inputs, output, per-layer output statistics and - for ``train`` - cost, gradient statistics (full tensors for
    everything up to 1024 elements), parameters after one ADAM step, BatchNorm running-stat updates, dropout masks."""
    from oracle import ref_harness as RH
    xs, y = net_eval_inputs(kind, cfg, seed, train)
    r = RH.run_reference_net(kind, cfg, xs, deterministic=not train, y=y, learning_rate=1e-3 if train else None)
    out = dict(out=r['out'], layer_nums=np.array(sorted(r['layer_out'])),
               layer_stats=np.stack([_stats(r['layer_out'][k]) for k in sorted(r['layer_out'])]))
    if train:
        names = r['param_order']
        out.update(cost=np.float64(r['cost']), param_order=np.array(names),
                   grad_stats=np.stack([_stats(r['grads'][n]) for n in names]),
                   newp_stats=np.stack([_stats(r['new_params'][n]) for n in names]),
                   adam_t_next=np.float64(r['adam_t_next']))
        for n in names:
            if r['grads'][n].size <= 1024:
                out['grad__' + n] = r['grads'][n]
                out['newp__' + n] = r['new_params'][n]
        for i, m in enumerate(r['masks']):
            out['mask%d' % i] = m.astype(np.float32)
        out['bn_names'] = np.array([n for n, _ in r['bn_updates']])
        for i, (n, v) in enumerate(r['bn_updates']):
            out['bn%d' % i] = v
    return out


NET_EVAL_CASES = [
    ('resnet0_train', 'ResNet', dict(type=0, nChan=1, wIn=128, hIn=128, batchSize=4, numJoints=1, nDims=30), 11, True),
    ('resnet1_det', 'ResNet', dict(type=1, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=14, nDims=3), 12, False),
    ('poseregnet0_train', 'PoseRegNet', dict(type=0, nChan=1, wIn=128, hIn=128, batchSize=4, numJoints=1, nDims=30), 13, True),
    ('poseregnet0_det', 'PoseRegNet', dict(type=0, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30), 14, False),
    ('scalenet1_det', 'ScaleNet', dict(type=1, nChan=1, wIn=128, hIn=128, batchSize=2, resizeFactor=2, numJoints=1, nDims=3), 15, False),
]


def net_evaluations():
    out = {}
    for tag, kind, cfg, seed, train in NET_EVAL_CASES:
        for k, v in net_eval_case(kind, cfg, seed, train).items():
            out['%s__%s' % (tag, k)] = v
    path = os.path.join(HERE, 'reference_net_eval.npz')
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


def main():
    from oracle import ref_harness as RH
    net_descriptions()
    net_evaluations()
    ref = RH.reference_modules()
    out = {}
    for name in ('NYU', 'ICVL', 'MSRA15'):
        out.update(augment_vectors(ref, name, 16, 300))
        out.update(pose_vectors(ref, name, 96, 310))
        out.update(geometry_vectors(ref, name, 64, 320))
    for name in ('NYU', 'ICVL'):
        out.update(cascade_vectors(ref, name, 5, 330))
    path = os.path.join(HERE, 'reference_pins.npz')
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == '__main__':
    main()
