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
def test_live_reference_network_constructor():
    desc = RH.describe_reference_net('ResNet', type=0, nChan=1, wIn=128, hIn=128, batchSize=3, numJoints=1, nDims=30)
    _check_net_against(desc)
    desc = RH.describe_reference_net('PoseRegNet', type=0, nChan=1, wIn=128, hIn=128, batchSize=3, numJoints=1, nDims=30)
    _check_net_against(desc)
