"""Pins of the CPU oracle against the REFERENCE'S OWN CODE.

tests/golden/reference_pins.npz was produced by executing the reference's sources (trainer/nettrainer.py augmentCrop,
util/handdetector.py, data/importers.py, data/transformations.py; oracle/ref_harness.py reads them from
/root/reference and applies a py2 -> py3 pass in memory) - see tests/golden/make_reference_vectors.py.  These tests
compare oracle/ with those outputs on the same inputs:
  * every INTEGER / INDEX result bit-exact: crop bounds, which source pixel every warped / resized pixel shows,
    z-threshold and background decisions;
  * floating-point results to a few float32 ulps: the fixture was computed under NumPy 2 (NEP 50) where
    ``float32_scalar (op) python_float`` stays float32, the oracle restates the reference-era NumPy 1.x which promoted
    to float64 - the tolerance covers exactly that last-bit difference (north_star's bar for floats is
    1e-4 relative).
When /root/reference is present (build container) the same comparisons also run LIVE on fresh seeds."""
import os
import numpy as np
import pytest

from oracle import augment as OA, cascade as OC, ref_harness as RH
from test_oracle_cascade import _tiny_fns

f32, f64 = np.float32, np.float64
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_pins.npz'))
CAMS = {'NYU': OA.NYU_CAM, 'ICVL': OA.ICVL_CAM, 'MSRA15': OA.MSRA_CAM}
AUG_MODES = ['com', 'rot', 'sc', 'none']
POSE_MODES = ['com', 'rot', 'sc', 'none', 'rot+com', 'rot+com+sc']
ULP = 2.0 ** -23


def g(prefix, name, key):
    return G['%s_%s_%s' % (prefix, name, key)]


def _check_augment(name, x, gt3Dcrop, com, cube, M, seed, ref_out):
    cam = OA.Camera(**CAMS[name])
    hd = OA.Hand(cam, use_cv2=True)
    rng = np.random.RandomState(seed)
    seen = set()
    for i in range(x.shape[0]):
        mode, off, rot, sc = OA.draw_aug_params(rng, len(AUG_MODES))
        img, lab, cube2, com2, M2 = OA.augment_crop(x[i].copy(), gt3Dcrop[i].copy(), com[i].copy(), cube[i].copy(),
                                                    M[i].copy(), AUG_MODES[mode], off, rot, sc, hd)
        r_img, r_lab, r_cube, r_com, r_M = [ref_out[k][i] for k in ('img', 'label', 'cube', 'com', 'M')]
        seen.add(AUG_MODES[mode])
        if AUG_MODES[mode] == 'sc':
            # the new cube is float32 under NumPy 2, float64 under NumPy 1.x: values differ in the last bits, the
            # warp's source pixels and the background / clamp decisions do not
            assert np.abs(img - r_img).max() <= 8 * ULP, (name, i)
            assert np.array_equal(img > 1.0 - 8 * ULP, r_img > 1.0 - 8 * ULP)    # same background pixels
        else:
            assert np.array_equal(img, r_img), (name, i, AUG_MODES[mode], int((img != r_img).sum()))
        assert np.abs(lab - r_lab).max() <= 16 * ULP, (name, i)
        np.testing.assert_allclose(np.asarray(cube2, f64), r_cube, rtol=4 * ULP)
        np.testing.assert_allclose(com2, r_com, rtol=4 * ULP)
        np.testing.assert_allclose(np.asarray(M2, f64), r_M, rtol=1e-6, atol=1e-5)
    return seen


@pytest.mark.parametrize('name', ['NYU', 'ICVL', 'MSRA15'])
def test_augment_crop_against_reference_fixture(name):
    ref_out = {k: g('augment', name, 'out_' + k) for k in ('img', 'label', 'cube', 'com', 'M')}
    seen = _check_augment(name, g('augment', name, 'x'), g('augment', name, 'gt3Dcrop'), g('augment', name, 'com'),
                          g('augment', name, 'cube'), g('augment', name, 'M'), int(g('augment', name, 'rng_seed')), ref_out)
    assert {'com', 'rot'} <= seen


@pytest.mark.parametrize('name', ['NYU', 'ICVL', 'MSRA15'])
def test_geometry_and_projections_against_reference_fixture(name):
    cam = OA.Camera(**CAMS[name])
    fx, fy = {'NYU': (588., 587.)}.get(name, (241.42, 241.42))
    coms, cube = g('geometry', name, 'coms'), tuple(g('geometry', name, 'cube'))
    for i, c in enumerate(coms):                       # float64 CoMs: both NumPy generations agree -> exact
        assert tuple(g('geometry', name, 'bounds')[i]) == tuple(f64(v) for v in OC.com_to_bounds(c, cube, fx, fy))
        assert np.array_equal(g('geometry', name, 'img_to_3d')[i], cam.jointImgTo3D(c))
    for i, p in enumerate(g('geometry', name, 'pts')):
        assert np.array_equal(g('geometry', name, 'to_img')[i], cam.joint3DToImg(p))
    # comToTransform (py2 integer division, the sz[1]/sz[0] swap) via the product's host mirror of the same function
    from util.handdetector import HandDetector
    hd = HandDetector(np.zeros((8, 8), f32) + 1., fx, fy)
    for i, c in enumerate(coms):
        assert np.array_equal(g('geometry', name, 'transform')[i], hd.comToTransform(c, cube, (128, 128)))
        assert tuple(g('geometry', name, 'bounds')[i]) == tuple(f64(v) for v in hd.comToBounds(c, cube))


def _check_cascade(name, frames, lastcom, cube, fx, fy, ref):
    cam = OA.Camera(**CAMS[name])
    refine_fn, _ = _tiny_fns(77)
    for i in range(frames.shape[0]):
        # what refineCoM feeds the refinement net: bit-exact
        b = OC.com_to_bounds(lastcom[i], cube, fx, fy)
        t = OC.refine_inputs(OC.resize_nn(OC.get_crop(frames[i], *b), (128, 128)), cube, lastcom[i])
        for k, key in enumerate(('x0', 'x1', 'x2')):
            assert np.array_equal(t[k][0, 0], ref[key][i]), (name, i, key)
        # refined CoM: float32 arithmetic on float32 scalars -> a few ulps between the NumPy generations
        loc = OC.track(frames[i], lastcom[i], cube, cam, fx, fy, refine_fn)
        np.testing.assert_allclose(loc, ref['loc'][i], rtol=4 * ULP, atol=2e-4)     # u = q*fx + ux cancels near the border
        # the pose net's crop for the reference's refined CoM: bit-exact (raw mm crop and normalised crop)
        raw, M, _ = OC.crop_area_3d(frames[i], ref['loc'][i], cube, fx, fy)
        assert np.array_equal(raw, ref['crop_raw'][i]) and np.array_equal(M, ref['M'][i])
        crop, _, com3D = OC.pipeline_detect(frames[i], ref['loc'][i], cube, cam, fx, fy)
        assert np.array_equal(crop, ref['crop'][i]), (name, i)
        np.testing.assert_allclose(com3D, ref['com3D'][i], rtol=4 * ULP, atol=2e-4)


@pytest.mark.parametrize('name', ['NYU', 'ICVL'])
def test_cascade_against_reference_fixture(name):
    fx, fy = g('cascade', name, 'fxfy')
    ref = {k: g('cascade', name, k) for k in ('x0', 'x1', 'x2', 'loc', 'crop_raw', 'crop', 'M', 'com3D')}
    _check_cascade(name, g('cascade', name, 'frames'), g('cascade', name, 'lastcom'), tuple(g('cascade', name, 'cube')),
                   float(fx), float(fy), ref)


@pytest.mark.parametrize('name', ['NYU', 'ICVL', 'MSRA15'])
def test_sample_random_poses_against_reference_fixture(name):
    cam = OA.Camera(**CAMS[name])
    r = [g('poses', name, 'out_' + k) for k in ('poses', 'com', 'cube', 'rot')]
    o = OA.sample_random_poses(cam, np.random.RandomState(int(g('poses', name, 'rng_seed'))), g('poses', name, 'base_poses'),
                               g('poses', name, 'base_com'), g('poses', name, 'base_cube'), r[0].shape[0], POSE_MODES,
                               retall=True)
    assert np.array_equal(o[3], r[3])                          # same random stream, same draw order
    assert np.array_equal(o[1], r[1])
    np.testing.assert_allclose(o[2], r[2], rtol=2 * ULP)
    assert np.abs(o[0] - r[0]).max() <= 32 * ULP               # normalised poses, |values| <= ~1


# ------------------------------------------------------------------------------------------------------------
# live: run the reference's code right now (build container only)
# ------------------------------------------------------------------------------------------------------------
live = pytest.mark.skipif(not RH.available(), reason="/root/reference is not present on this machine")


@live
@pytest.mark.parametrize('name', ['NYU', 'MSRA15'])
def test_live_reference_augment_crop(name):
    from data import synthetic
    ref = RH.reference_modules()
    rdi = getattr(ref['importers'], name + 'Importer')('/nonexistent/')
    augmentCrop = RH.reference_function('trainer/nettrainer.py', 'augmentCrop', {'numpy': np})
    n, seed = 60, 4100
    ds = synthetic.generate(name, n, seed=seed)
    cam = OA.Camera(**CAMS[name])
    rhd = ref['handdetector'].HandDetector(np.zeros((128, 128), f32) + 1., abs(rdi.fx), abs(rdi.fy), importer=rdi)

    class Self(object):
        rng = np.random.RandomState(seed + 1)
    com = np.stack([cam.joint3DToImg(c) for c in ds['com3D']])
    res = [augmentCrop(Self, ds['x'][i, 0].copy(), ds['gt3Dcrop'][i].copy(), com[i].copy(), ds['cube'][i].copy(),
                       ds['M'][i].copy(), AUG_MODES, rhd) for i in range(n)]
    ref_out = dict(img=[r[0] for r in res], label=[r[2] for r in res], cube=[np.asarray(r[3], f64) for r in res],
                   com=[r[4] for r in res], M=[np.asarray(r[5], f64) for r in res])
    assert _check_augment(name, ds['x'][:, 0], ds['gt3Dcrop'], com, ds['cube'], ds['M'], seed + 1, ref_out) == set(AUG_MODES)


@live
def test_live_reference_cascade():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    import make_reference_vectors as MK
    out = MK.cascade_vectors(RH.reference_modules(), 'NYU', 12, 977)
    ref = {k: out['cascade_NYU_' + k] for k in ('x0', 'x1', 'x2', 'loc', 'crop_raw', 'crop', 'M', 'com3D')}
    _check_cascade('NYU', out['cascade_NYU_frames'], out['cascade_NYU_lastcom'], tuple(out['cascade_NYU_cube']), 588., 587., ref)


# ------------------------------------------------------------------------------------------------------------
# network classes: layer lists, dimensions, parameter order and INITIAL WEIGHTS against the reference's constructors
# ------------------------------------------------------------------------------------------------------------
import hashlib   # noqa: E402
import json      # noqa: E402

NETS = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_nets.json')))


def _product_net(kind, cfg):
    if kind == 'ResNet':
        from net.resnet import ResNet as N, ResNetParams as P
    elif kind == 'PoseRegNet':
        from net.poseregnet import PoseRegNet as N, PoseRegNetParams as P
    else:
        from net.scalenet import ScaleNet as N, ScaleNetParams as P
    return N(np.random.RandomState(23455), cfgParams=P(**cfg))


def _oracle_net(kind, cfg):
    from oracle import nets as ON
    cfg = dict(cfg)
    build = {'ResNet': ON.build_resnet, 'PoseRegNet': ON.build_poseregnet, 'ScaleNet': ON.build_scalenet}[kind]
    return build(np.random.RandomState(23455), **cfg)


def _sha(v):
    return hashlib.sha1(np.ascontiguousarray(v, dtype=f32).tobytes()).hexdigest()


def _check_net_against(desc):
    kind, cfg = desc['kind'], desc['cfg']
    net = _product_net(kind, cfg)
    assert len(net.layers) == len(desc['layers'])
    for l, rl in zip(net.layers, desc['layers']):
        where = (kind, cfg['type'], rl['layerNum'], rl['cls'])
        assert type(l).__name__ == rl['cls'] and l.layerNum == rl['layerNum'], where
        assert list(l.cfgParams.inputDim) == rl['inputDim'] and list(l.cfgParams.outputDim) == rl['outputDim'], where
        mine = list(l.params) + list(l.params_nontrained)
        ref = rl['params'] + rl['params_nontrained']
        assert [p.name for p in mine] == [p['name'] for p in ref], where
        for p, rp in zip(mine, ref):
            v = p.get_value()
            assert list(v.shape) == rp['shape'] and _sha(v) == rp['sha1'], where + (p.name,)
        assert [p.name for p in l.weights] == rl['weights'], where
    assert [p.name for p in net.params] == desc['net_params']
    assert list(net.cfgParams.outputDim) == desc['outputDim']
    # the oracle the GPU parity tests compare against starts from the same weights (same draws, same order):
    # its trainable parameters, in order, are the reference's net.params values
    onet = _oracle_net(kind, cfg)
    ref_by_name = {p['name']: p for rl in desc['layers'] for p in rl['params']}
    ovals = [p.detach().numpy() for p in onet.params]
    assert len(ovals) == len(desc['net_params'])
    for name, ov in zip(desc['net_params'], ovals):
        rp = ref_by_name[name]
        if name.startswith('conv') and name[4] == 'W':
            assert list(ov.shape) == rp['shape']
        assert ov.size == int(np.prod(rp['shape'])) and np.isclose(float(ov.astype(f64).sum()), rp['sum'], rtol=1e-6, atol=1e-5), name
        assert _sha(ov.reshape(rp['shape'])) == rp['sha1'], (kind, cfg['type'], name)


@pytest.mark.parametrize('idx', range(len(NETS)))
def test_network_classes_against_reference_fixture(idx):
    _check_net_against(NETS[idx])


@live
@pytest.mark.parametrize('kind,types', [('ResNet', (0, 1, 2, 3, 4)), ('PoseRegNet', (0, 11)), ('ScaleNet', (1,))])
def test_live_reference_network_constructor(kind, types):
    """every network type the reference can build (its other type numbers raise NotImplementedError there too)"""
    for t in types:
        cfg = dict(type=t, nChan=1, wIn=128, hIn=128, batchSize=3, numJoints=1, nDims=30)
        if kind == 'ScaleNet':
            cfg.update(resizeFactor=2, nDims=3)
        _check_net_against(RH.describe_reference_net(kind, **cfg))


# ------------------------------------------------------------------------------------------------------------
# network ARITHMETIC: the reference's own layer / cost / T.grad / ADAM code, evaluated with oracle/eager_theano.py in
# place of Theano (tests/golden/reference_net_eval.npz), against the oracle.  Covers everything the reference
# writes in Python around Theano's primitives; the primitives themselves follow their documented semantics.
# ------------------------------------------------------------------------------------------------------------
import sys as _sys   # noqa: E402
_sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
import make_reference_vectors as MK   # noqa: E402

NE = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_net_eval.npz'))


def _rel(a, b):
    a, b = np.asarray(a, f64), np.asarray(b, f64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def _stats(a):
    return MK._stats(a)


def _check_net_eval(kind, cfg, seed, train, ref):
    """ref: dict with the keys of make_reference_vectors.net_eval_case"""
    import torch
    from oracle import nets as ON
    xs, y = MK.net_eval_inputs(kind, cfg, seed, train)
    onet = _oracle_net(kind, cfg)
    tx = [torch.from_numpy(x) for x in xs]
    tin = tx if len(tx) > 1 else tx[0]
    collect = {}
    if not train:
        with torch.no_grad():
            out, _ = onet.forward(tin, deterministic=True, collect=collect)
    else:
        masks = [torch.from_numpy(np.asarray(ref['mask%d' % i])) for i in range(8) if ('mask%d' % i) in ref] or None
        with torch.no_grad():
            fwd, _ = onet.forward(tin, deterministic=False, masks=masks, collect=collect)
        adam = ON.Adam(onet.params)
        before = [p.detach().numpy().copy() for p in onet.params]
        cost, out, grads = ON.train_step(onet, adam, tin, torch.from_numpy(y), 1e-3, cfg['numJoints'], cfg['nDims'],
                                         masks=masks)
        assert abs(cost - float(ref['cost'])) <= 1e-10 * abs(float(ref['cost']))
        names = [str(n) for n in ref['param_order']]
        assert len(names) == len(grads)
        for i, n in enumerate(names):
            g = grads[i].numpy()
            rs = ref['grad_stats'][i]
            scale = max(rs[2], 1e-30)
            if rs[2] > 1e-10:                       # conv biases in front of a BatchNorm have an exactly-zero gradient
                assert np.abs(_stats(g) - rs)[[0, 1, 2]].max() <= 1e-7 * max(rs[1], scale), n
            else:
                assert np.abs(g).max() <= 1e-10, n
            if ('grad__' + n) in ref and rs[2] > 1e-10:
                assert _rel(g.reshape(ref['grad__' + n].shape), ref['grad__' + n]) <= 1e-8, n
            # one ADAM step (the oracle folds the constants in float32 like Theano's floatX graph: 1e-6 relative)
            p_new = onet.params[i].detach().numpy()
            if rs[2] > 1e-10 and ('newp__' + n) in ref:
                want = ref['newp__' + n]
                assert np.abs(p_new.reshape(want.shape) - want).max() <= 2e-6 * max(np.abs(want).max(), 1e-3), n
                assert np.abs(p_new - before[i]).max() > 0          # the step moved the parameter
        assert float(ref['adam_t_next']) == 2.0 and float(adam.t) == 2.0
        # BatchNorm running statistics after the step (EMA of mean AND inv_std, batchnormlayer.py:164-172)
        bn_layers = [l for l in onet.layers if l.kind == 'bn']
        assert len(ref['bn_names']) == 2 * len(bn_layers)
        for i, l in enumerate(bn_layers):
            for k in (0, 1):
                v = l.nontrained[k]
                v = v.numpy() if hasattr(v, 'numpy') else np.asarray(v)
                assert _rel(v, ref['bn%d' % (2 * i + k)]) <= 1e-6, (i, k)
    out = out.numpy() if hasattr(out, 'numpy') else out
    assert _rel(out, ref['out']) <= 1e-10
    # every layer's output (statistics): the wiring, not just the end result
    nums = [int(n) for n in ref['layer_nums']]
    seen = 0
    for row, ln in enumerate(nums):
        if ln in collect:
            s = _stats(collect[ln].detach().numpy())
            rs = ref['layer_stats'][row]
            assert np.abs(s - rs)[[0, 1, 2]].max() <= 1e-8 * max(rs[1], 1e-30), (kind, ln)
            seen += 1
    assert seen >= 0.9 * len(nums)


@pytest.mark.parametrize('case', range(len(MK.NET_EVAL_CASES)))
def test_network_arithmetic_against_reference_fixture(case):
    tag, kind, cfg, seed, train = MK.NET_EVAL_CASES[case]
    ref = {k[len(tag) + 2:]: NE[k] for k in NE.files if k.startswith(tag + '__')}
    _check_net_eval(kind, cfg, seed, train, ref)


@live
def test_live_reference_train_step():
    kind, cfg = 'ResNet', dict(type=0, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30)
    _check_net_eval(kind, cfg, 901, True, MK.net_eval_case(kind, cfg, 901, True))


@live
def test_live_reference_trainer_host_arithmetic():
    """NetTrainerParams.lr_of_ep and NetTrainer's batch / macro-batch arithmetic and alignData (trainer/nettrainer.py:
    47-72, 365-487) executed from the reference source against the product's trainer/nettrainer.py."""
    import types
    from trainer.nettrainer import NetTrainer, NetTrainerParams
    ref_init = RH.reference_function('trainer/nettrainer.py', '__init__', {'numpy': np})       # NetTrainerParams.__init__
    rp = types.SimpleNamespace()
    ref_init(rp)
    mp = NetTrainerParams()
    for lr in (0.01, 1e-3, 3e-4):
        rp.learning_rate = mp.learning_rate = lr
        for ep in range(0, 60):
            a, b = rp.lr_of_ep(ep), mp.lr_of_ep(ep)
            assert type(a) is type(b) and a == b, (lr, ep, a, b)
    for k in ('batch_size', 'momentum', 'weightreg_factor', 'use_early_stopping', 'snapshot_last', 'snapshot_freq',
              'para_augment', 'para_num_proc', 'para_load', 'force_macrobatch_reload', 'pad_random',
              'validation_frequency'):
        assert getattr(rp, k) == getattr(mp, k), k
    names = ['getSizeMiniBatch', 'getSizeMacroBatch', 'getNumFullMiniBatches', 'getNumMiniBatches', 'getNumMacroBatches',
             'getNumMiniBatchesPerMacroBatch', 'getNumSamplesPerMacroBatch', 'getNumMiniBatchesPerChunk',
             'getNumSamplesPerChunk', 'getGPUMemAligned', 'alignData']
    ref_fns = {n: RH.reference_function('trainer/nettrainer.py', n, {'numpy': np}) for n in names}
    rng = np.random.RandomState(4)
    for trial in range(40):
        cfg = types.SimpleNamespace(batch_size=int(rng.choice([8, 32, 128])), pad_random=bool(trial % 3))
        rs = types.SimpleNamespace(cfgParams=cfg)
        for n, fn in ref_fns.items():
            setattr(rs, n, types.MethodType(fn, rs))
        ms = NetTrainer.__new__(NetTrainer)
        ms.cfgParams = cfg
        n_samples = int(rng.randint(1, 3000))
        for s in (rs, ms):
            s.numTrainSamples = n_samples
            s.sampleSize = 64. / 1024.
            s.trainSize = n_samples * s.sampleSize
            s.memorySize = float(rng.choice([16., 64., 4096.]))
            s.numChunks = 1
        rs.memorySize = ms.memorySize
        for n in names[:-1]:
            assert getattr(rs, n)() == getattr(ms, n)(), (n, trial)
        data = rng.randn(n_samples, 3).astype(np.float32)
        for align in (None, cfg.batch_size * int(np.ceil(n_samples / float(cfg.batch_size)))):
            if align is None and rs.getNumSamplesPerMacroBatch() < n_samples:
                continue
            assert np.array_equal(rs.alignData(data, alignSize=align), ms.alignData(data, alignSize=align)), trial


@live
@pytest.mark.parametrize('kind,cfg', [
    ('PoseRegNet', dict(type=0, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30)),
    ('ScaleNet', dict(type=1, nChan=1, wIn=128, hIn=128, batchSize=2, resizeFactor=2, numJoints=1, nDims=3)),
    ('ResNet', dict(type=1, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=14, nDims=3))])
def test_live_checkpoints_interoperate_with_reference_netbase(kind, cfg, tmp_path):
    """a pickle written by the reference's NetBase.save loads into the product net, and a pickle
    written by the product loads through the reference's NetBase.load (net/netbase.py:405-477) - same keys
    ('{layerNum}-values'), same order inside a layer (params then params_nontrained), same array layouts."""
    ref_file = os.path.join(str(tmp_path), 'from_reference.pkl')
    mine_file = os.path.join(str(tmp_path), 'from_product.pkl')
    ref = RH.reference_checkpoint_roundtrip(kind, cfg, save_path=ref_file, seed=777)       # reference weights, seed 777
    net = _product_net(kind, cfg)                                                        # product weights, seed 23455
    some = [p for l in net.layers for p in l.params][0]
    assert not np.array_equal(some.get_value(), ref['values'][some.name])
    net.load(ref_file)
    for l in net.layers:
        for p in list(l.params) + list(l.params_nontrained):
            assert np.array_equal(p.get_value(), ref['values'][p.name]), p.name
    # the description the reference compares on load (:446-452); NumPy 2 prints numpy.prod's result as np.int64(..)
    import re
    assert str(net) == re.sub(r'np\.int64\((\d+)\)', r'\1', ref['network'])
    # the other direction: perturb, save with the product, load with the reference's code
    rng = np.random.RandomState(5)
    for l in net.layers:
        for p in list(l.params) + list(l.params_nontrained):
            p.set_value((p.get_value() + rng.randn(*p.get_value().shape) * 0.01).astype(np.float32))
    net.save(mine_file)
    back = RH.reference_checkpoint_roundtrip(kind, cfg, load_path=mine_file, seed=1)
    for l in net.layers:
        for p in list(l.params) + list(l.params_nontrained):
            assert np.array_equal(p.get_value(), back['values'][p.name]), p.name


@live
def test_live_reference_layer_init_values():
    """net/layer.py getInitVals (all methods / modes, orthogonal included) and orthogonalize, reference vs product."""
    from unittest import mock
    from net.layer import Layer
    theano = mock.MagicMock()
    theano.config.floatX = 'float32'
    fake = {'theano': theano, 'theano.tensor': theano.tensor}
    with RH._reference_net_modules(fake) as mods:
        RL = mods['net.layer'].Layer
        for seed, (shape, mode) in enumerate([((8, 1, 5, 5), 'conv'), ((16, 8, 3, 3), 'conv'), ((200, 30), 'fc')]):
            for method, act in (('He', None), ('Xavier', None), ('sigmoid', None), ('tanh', None), (None, 'ReLU'),
                                (None, 'None')):
                for orth in (False, True) if mode == 'conv' else (False,):
                    a = RL(np.random.RandomState(seed)).getInitVals(shape, mode, act_fn=act, method=method, orthogonal=orth)
                    b = Layer(np.random.RandomState(seed)).getInitVals(shape, mode, act_fn=act, method=method, orthogonal=orth)
                    assert a.dtype == b.dtype and np.array_equal(a, b), (shape, mode, method, act, orth)


@live
@pytest.mark.parametrize('name', ['NYU', 'ICVL'])
def test_live_reference_dataset_stack(name):
    """The reference's own Dataset.imgStackDepthOnly (data/dataset.py:75-111) on a synthetic NamedImgSequence yields
    exactly the crops / labels data.synthetic.generate hands to the tests and the bench (its normalisation restates
    dataset.py:99-103); the product's data.dataset.Dataset produces the same stack on the device (GPU test)."""
    from data import synthetic
    ref = RH.reference_modules()
    seq = synthetic.generate_sequence(name, 6, seed=41)
    RSeq = ref['basetypes'].NamedImgSequence(seq.name, [ref['basetypes'].DepthFrame(*f) for f in seq.data], seq.config)
    img, lab = ref['dataset'].Dataset([RSeq]).imgStackDepthOnly('train')
    ds = synthetic.generate(name, 6, seed=41)
    assert img.dtype == np.float32 and np.array_equal(img, ds['x'])
    assert np.array_equal(lab, ds['gt3D'])
    assert np.array_equal(np.stack([f.com for f in seq.data]), ds['com3D'])
    assert np.array_equal(np.stack([f.T for f in seq.data]), ds['M'])
    assert np.array_equal(np.stack([f.gt3Dcrop for f in seq.data]), ds['gt3Dcrop'])


@live
def test_live_reference_augment_poses_driver():
    """PoseRegNetTrainer.augment_poses (trainer/poseregnettrainer.py:221-264) executed from the reference - with the
    reference's augmentCrop, HandDetector, importer and a PCA projection - against oracle.augment_poses on the same
    random stream: crops bit-exact for com / rot / none, embedded labels to float32 rounding."""
    import types
    from data import synthetic
    ref = RH.reference_modules()
    name, n = 'NYU', 40
    ds = synthetic.generate(name, n, seed=77)
    comp, mean = synthetic.random_orthonormal_pca(30, ds['gt3Dcrop'].shape[1] * 3, seed=2)

    class Proj(object):                                   # sklearn's PCA.transform
        def transform(self, X):
            return np.dot(X - mean, comp.T)
    rdi = ref['importers'].NYUImporter('/nonexistent/')
    rhd = ref['handdetector'].HandDetector(np.zeros((128, 128), f32) + 1., abs(rdi.fx), abs(rdi.fy), importer=rdi)
    cam = OA.Camera(**CAMS[name])
    modes = ['com', 'rot', 'none']
    tr = types.SimpleNamespace(train_data_xDB=ds['x'], train_data_comDB=ds['com3D'], train_data_cubeDB=ds['cube'],
                               train_data_MDB=ds['M'], train_gt3DcropDB=ds['gt3Dcrop'], rng=np.random.RandomState(3),
                               getNumMacroBatches=lambda: 1)
    tr.augmentCrop = types.MethodType(RH.reference_function('trainer/nettrainer.py', 'augmentCrop', {'numpy': np}), tr)
    augment_poses = RH.reference_function('trainer/poseregnettrainer.py', 'augment_poses', {'numpy': np})

    class OracleProjectionImporter(object):     # 'di': joint3DToImg with NumPy-1.x rounding (float32 input: the
                                                # reference-as-run-here is 1-2 ulp off, see oracle/ref_harness.py)
        joint3DToImg = staticmethod(cam.joint3DToImg)
        jointImgTo3D = staticmethod(rdi.jointImgTo3D)
        joints3DToImg = staticmethod(rdi.joints3DToImg)
        jointsImgTo3D = staticmethod(rdi.jointsImgTo3D)
    rhd.importer = rdi
    params = {'fun': 'augment_poses', 'args': {'normZeroOne': False, 'di': OracleProjectionImporter, 'aug_modes': modes,
                                               'hd': rhd, 'proj': Proj()}}
    new_data = {'train_data_x': np.zeros((n, 1, 128, 128), f32), 'train_data_y': np.zeros((n, 30), f32)}
    augment_poses(tr, params, 0, False, list(range(n)), list(range(n)), new_data)
    rng = np.random.RandomState(3)
    draws = [OA.draw_aug_params(rng, len(modes)) for _ in range(n)]
    ox, oy = OA.augment_poses(ds['x'], ds['com3D'], ds['cube'], ds['M'], ds['gt3Dcrop'], list(range(n)), draws, modes,
                              cam, OA.Hand(cam, use_cv2=True), pca_mean=mean, pca_components=comp)
    assert np.array_equal(new_data['train_data_x'], ox)
    assert np.abs(new_data['train_data_y'] - oy).max() <= 64 * ULP
    assert {modes[d[0]] for d in draws} == set(modes)


@live
def test_live_reference_importer_constants():
    """Camera models and per-dataset constants of the product's data/importers.py against the reference's importers
    (data/importers.py:186-210, 536-570, 880-920): intrinsics, joint counts, crop joints, default cubes, projections."""
    from data import importers as P
    ref = RH.reference_modules()['importers']
    rng = np.random.RandomState(0)
    for cls in ('ICVLImporter', 'MSRA15Importer', 'NYUImporter'):
        r, p = getattr(ref, cls)('/nonexistent/'), getattr(P, cls)(None)
        for k in ('fx', 'fy', 'ux', 'uy', 'numJoints', 'crop_joint_idx', 'depth_map_size', 'default_cubes', 'sides'):
            assert getattr(r, k) == getattr(p, k), (cls, k)
        if hasattr(r, 'restrictedJointsEval'):
            assert list(r.restrictedJointsEval) == list(p.restrictedJointsEval)
        assert np.array_equal(r.getCameraProjection(), p.getCameraProjection())
        pts = np.stack([rng.uniform(-200, 200, 50), rng.uniform(-200, 200, 50), rng.uniform(300, 1000, 50)], axis=1)
        for q in pts:                                      # float64 points: identical in both NumPy generations
            assert np.array_equal(r.joint3DToImg(q), p.joint3DToImg(q))
            uvd = np.asarray(p.joint3DToImg(q), np.float64)
            assert np.array_equal(r.jointImgTo3D(uvd), p.jointImgTo3D(uvd))
        assert np.array_equal(r.joints3DToImg(pts), p.joints3DToImg(pts))
        assert np.array_equal(r.jointsImgTo3D(pts), p.jointsImgTo3D(pts))


@live
@pytest.mark.parametrize('kind,cfg,train', [
    ('ResNet', dict(type=4, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30), True),      # dropout + 30-D bottleneck
    ('ResNet', dict(type=2, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30), False),
    ('PoseRegNet', dict(type=11, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30), True),
    ('ScaleNet', dict(type=1, nChan=1, wIn=128, hIn=128, batchSize=2, resizeFactor=2, numJoints=1, nDims=3), True)])
def test_live_reference_arithmetic_other_types(kind, cfg, train):
    _check_net_eval(kind, cfg, 555, train, MK.net_eval_case(kind, cfg, 555, train))


@live
def test_live_reference_adam_over_several_steps():
    """Optimizer.ADAM (trainer/optimizer.py:58-90) evaluated from the reference source for 5 consecutive steps (state
    carried between the rebuilt graphs) against oracle.Adam - the timestep-dependent parts (bias corrections
    1 - beta^t, beta1 * gamma^(t-1), t += 1) are what a single step cannot show.  The reference evaluation is float64,
    the oracle folds constants in float32 like a floatX graph: 1e-5 of the accumulated update."""
    import torch
    from oracle import nets as ON
    rng = np.random.RandomState(12)
    shapes = [(7, 5), (11,)]
    p0 = [rng.randn(*s).astype(f32) for s in shapes]
    grads = [[(rng.randn(*s) * (10. ** rng.uniform(-3, 1))).astype(f32) for s in shapes] for _ in range(5)]
    lrs = [1e-3, 1e-3, 3e-4, 3e-4, 1e-4]
    hist = RH.run_reference_adam(p0, grads, lrs)
    ps = [torch.tensor(p.copy()) for p in p0]
    adam = ON.Adam(ps)
    for step in range(5):
        adam.step([torch.from_numpy(g) for g in grads[step]], lrs[step])
        for p, want, start in zip(ps, hist[step], p0):
            moved = np.abs(want - start).max()
            # + a few float32 ulps of the parameter itself (the oracle stores float32, the eager run float64)
            assert np.abs(p.numpy() - want).max() <= 1e-5 * moved + 4 * ULP * np.abs(want).max(), step
    assert float(adam.t) == 6.0


@live
def test_live_reference_layer_params_classes():
    """Every *LayerParams class on the path: derived dimensions, memory requirement, activation name and output range
    after construction and after reassigning inputDim / filters / stride / pool size, reference vs product."""
    from unittest import mock
    import importlib
    theano = mock.MagicMock()
    theano.config.floatX = 'float32'
    fake = {'theano': theano, 'theano.tensor': theano.tensor}

    def snapshot(p):
        out = {}
        for k in ('inputDim', 'outputDim', 'activation_str', 'filter_shape', 'image_shape', 'stride', 'border_mode',
                  'nFilters', 'filterDim', 'poolsize', 'poolType', 'hasBias'):
            if hasattr(p, k):
                v = getattr(p, k)
                out[k] = tuple(int(x) for x in v) if isinstance(v, (tuple, list)) else v
        out['mem'] = int(p.getMemoryRequirement()) if hasattr(p, 'getMemoryRequirement') else None
        out['range'] = [float(x) for x in p.getOutputRange()]
        return out

    def relu_like(mods):
        return mods['util.theano_helpers'].ReLU

    with RH._reference_net_modules(fake) as mods:
        ref_cases = []
        R = mods
        act = relu_like(mods)
        cp = R['net.convpoollayer'].ConvPoolLayerParams(inputDim=(4, 1, 128, 128), nFilters=8, filterDim=(5, 5),
                                                        poolsize=(4, 4), activation=act)
        ref_cases.append(('cp', snapshot(cp)))
        cp.inputDim = (4, 1, 64, 64)
        cp.poolsize = (2, 2)
        ref_cases.append(('cp2', snapshot(cp)))
        cv = R['net.convlayer'].ConvLayerParams(inputDim=(4, 32, 64, 64), nFilters=16, filterDim=(1, 1), stride=(2, 2),
                                                border_mode='same', activation=None, init_method='He')
        ref_cases.append(('cv', snapshot(cv)))
        cv.filterDim = (3, 3)
        cv.stride = (1, 1)
        cv.nFilters = 64
        ref_cases.append(('cv2', snapshot(cv)))
        cv.border_mode = 'valid'
        ref_cases.append(('cv3', snapshot(cv)))
        hl = R['net.hiddenlayer'].HiddenLayerParams(inputDim=(4, 16384), outputDim=(4, 1024), activation=act)
        ref_cases.append(('hl', snapshot(hl)))
        hn = R['net.hiddenlayer'].HiddenLayerParams(inputDim=(4, 30), outputDim=(4, 42), activation=None)
        ref_cases.append(('hn', snapshot(hn)))
        bn = R['net.batchnormlayer'].BatchNormLayerParams(inputDim=(4, 64, 8, 8))
        ref_cases.append(('bn', snapshot(bn)))
        nl = R['net.nonlinearitylayer'].NonlinearityLayerParams(inputDim=(4, 64, 8, 8), activation=act)
        ref_cases.append(('nl', snapshot(nl)))
        dr = R['net.dropoutlayer'].DropoutLayerParams(inputDim=(4, 1024), outputDim=(4, 1024))
        ref_cases.append(('dr', snapshot(dr)))
    from util.theano_helpers import ReLU
    from net.convpoollayer import ConvPoolLayerParams
    from net.convlayer import ConvLayerParams
    from net.hiddenlayer import HiddenLayerParams
    from net.batchnormlayer import BatchNormLayerParams
    from net.nonlinearitylayer import NonlinearityLayerParams
    from net.dropoutlayer import DropoutLayerParams
    mine = []
    cp = ConvPoolLayerParams(inputDim=(4, 1, 128, 128), nFilters=8, filterDim=(5, 5), poolsize=(4, 4), activation=ReLU)
    mine.append(('cp', snapshot(cp)))
    cp.inputDim = (4, 1, 64, 64)
    cp.poolsize = (2, 2)
    mine.append(('cp2', snapshot(cp)))
    cv = ConvLayerParams(inputDim=(4, 32, 64, 64), nFilters=16, filterDim=(1, 1), stride=(2, 2), border_mode='same',
                         activation=None, init_method='He')
    mine.append(('cv', snapshot(cv)))
    cv.filterDim = (3, 3)
    cv.stride = (1, 1)
    cv.nFilters = 64
    mine.append(('cv2', snapshot(cv)))
    cv.border_mode = 'valid'
    mine.append(('cv3', snapshot(cv)))
    mine.append(('hl', snapshot(HiddenLayerParams(inputDim=(4, 16384), outputDim=(4, 1024), activation=ReLU))))
    mine.append(('hn', snapshot(HiddenLayerParams(inputDim=(4, 30), outputDim=(4, 42), activation=None))))
    mine.append(('bn', snapshot(BatchNormLayerParams(inputDim=(4, 64, 8, 8)))))
    mine.append(('nl', snapshot(NonlinearityLayerParams(inputDim=(4, 64, 8, 8), activation=ReLU))))
    mine.append(('dr', snapshot(DropoutLayerParams(inputDim=(4, 1024), outputDim=(4, 1024)))))
    for (tag, r), (_, m) in zip(ref_cases, mine):
        assert r == m, (tag, r, m)


@live
def test_live_reference_helpers_and_entry_script_imports():
    """(1) util/helpers.py against the reference's helpers (shuffle order with the same RandomState, chunks, cartesian,
    gaussian_kernel).  (2) Every import statement of the reference's entry scripts (main_*_posereg_embedding*.py,
    test_realtimepipeline.py) resolves against the product package: the scripts' own data preparation / training /
    evaluation code finds the modules, classes and functions it names.  Stubbed: matplotlib (absent in this image,
    plots are out of scope), cPickle (= pickle), util.cameradevice (camera capture, out of scope)."""
    import ast
    import sys
    import types
    import util.helpers as PH
    RHm = RH.load_module('ref_helpers_tmp', 'util/helpers.py')
    sys.modules.pop('ref_helpers_tmp', None)
    a = [np.arange(40).reshape(20, 2).copy(), np.arange(20).copy()]
    b = [x.copy() for x in a]
    RHm.shuffle_many_inplace(a, np.random.RandomState(4))
    PH.shuffle_many_inplace(b, np.random.RandomState(4))
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and not np.array_equal(a[1], np.arange(20))
    assert list(RHm.chunks(list(range(11)), 4)) == list(PH.chunks(list(range(11)), 4))
    assert np.array_equal(RHm.cartesian(([1, 2, 3], [4, 5], [6, 7])), PH.cartesian(([1, 2, 3], [4, 5], [6, 7])))
    for k in (3, 5, 8):
        np.testing.assert_allclose(RHm.gaussian_kernel(k), PH.gaussian_kernel(k), rtol=1e-6)
    stubs = {'matplotlib': types.ModuleType('matplotlib'), 'matplotlib.pyplot': types.ModuleType('matplotlib.pyplot'),
             'cPickle': __import__('pickle'), 'util.cameradevice': types.ModuleType('util.cameradevice')}
    stubs['matplotlib'].use = lambda *a, **k: None
    stubs['matplotlib'].pyplot = stubs['matplotlib.pyplot']
    stubs['util.cameradevice'].CreativeCameraDevice = stubs['util.cameradevice'].FileDevice = object
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        for script in ('main_nyu_posereg_embedding.py', 'main_icvl_posereg_embedding.py',
                       'main_msra15_posereg_embedding_crossval.py', 'test_realtimepipeline.py'):
            tree = ast.parse(RH.py3_source(os.path.join(RH.REF_SRC, script)))
            imports = [n for n in ast.walk(tree) if isinstance(n, (ast.Import, ast.ImportFrom))]
            assert len(imports) >= 8
            for node in imports:
                code = compile(ast.Module(body=[node], type_ignores=[]), script, 'exec')
                try:
                    exec(code, {})
                except Exception as e:      # name the statement that does not resolve
                    raise AssertionError("%s: `%s` does not resolve against the product package: %r"
                                         % (script, ast.unparse(node), e))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@live
def test_live_reference_pipeline_detect_and_estimate_pose():
    """RealtimeHandposePipeline.detect (tracking branch) and estimatePose (util/realtimehandposepipeline.py:296-368)
    executed from the reference, for the left and the (mirrored) right hand, against oracle.cascade_frame."""
    from data import synthetic
    ref = RH.reference_modules()
    rdi = ref['importers'].NYUImporter('/nonexistent/')
    cam = OA.Camera(**CAMS['NYU'])
    fr = synthetic.generate_frames('NYU', 6, seed=71, edge_fraction=0.5)
    lastcom = fr['lastcom'].astype(f32).astype(f64)
    refine_fn, pose_fn = _tiny_fns(77)
    for right in (False, True):
        for i in range(6):
            net, pnet = MK.RecordingNet(refine_fn), MK._PoseNetStub(pose_fn)
            pipe = MK.reference_pipeline(ref, rdi, 588., 587., fr['cube'], net, pnet, right_hand=right)
            pipe.lastcom = lastcom[i].copy()
            crop, M, com3D = pipe.detect(fr['frames'][i])
            jj = pipe.estimatePose(crop, com3D)
            pose = jj * fr['cube'][2] / 2. + com3D                      # processVideo, :197-198
            want = OC.cascade_frame(fr['frames'][i], lastcom[i], fr['cube'], cam, 588., 587., refine_fn, pose_fn,
                                    right_hand=right)
            # same refined CoM in (the reference's): crop, mirrored net input and pose agree
            crop_o, M_o, c3_o = OC.pipeline_detect(fr['frames'][i], pipe.lastcom, fr['cube'], cam, 588., 587.)
            assert np.array_equal(crop, crop_o) and np.array_equal(M, M_o)
            assert np.array_equal(pnet.seen[0, 0], crop_o[:, ::-1] if right else crop_o)
            jo = OC.estimate_pose(crop_o, pose_fn, right)
            assert np.array_equal(jj, jo)
            np.testing.assert_allclose(pose, jo * f32(fr['cube'][2] / 2.) + c3_o, rtol=1e-6, atol=1e-4)
            np.testing.assert_allclose(want['com'], pipe.lastcom, rtol=4 * ULP, atol=2e-4)


@live
@pytest.mark.parametrize('kind,cfg', [
    ('PoseRegNet', dict(type=0, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30)),
    ('ResNet', dict(type=0, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=30))])
def test_live_reference_netbase_bookkeeping(kind, cfg):
    """NetBase (net/netbase.py:141-203, 318-403): deterministic switch, hasDropout, params / weights views and their
    filters - the same sequence of calls on the reference's net and on the product's."""
    ref = RH.reference_netbase_behaviour(kind, cfg)
    mine = RH.observe_netbase(_product_net(kind, cfg))
    assert ref == mine, {k: (ref[k], mine[k]) for k in ref if ref[k] != mine[k]}
