"""GPU parity of HandDetector.sampleRandomPoses (reference src/util/handdetector.py:805-909, called by
main_nyu_posereg_embedding.py:87-88 to fit the PCA prior on 1e6 sampled poses): ``dpp_sample_poses`` against the CPU
oracle (itself pinned to the reference's own code, tests/test_reference_pins.py) and against the committed reference
fixture.  Float work: the kernel repeats the reference's operations one by one in the same precisions, so the
expected difference is 0; the asserted bar is 4 float32 ulps of the normalised poses (north_star: 1e-4 relative)."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32 = np.float32
ULP = 2.0 ** -23
MODES = ['com', 'rot', 'sc', 'none', 'rot+com', 'rot+com+sc']


@pytest.mark.parametrize('name', ['NYU', 'ICVL', 'MSRA15'])
def test_sample_random_poses_matches_oracle(name):
    from data import synthetic
    from util.handdetector import HandDetector
    from oracle import augment as OA
    cam = OA.Camera(**{'NYU': OA.NYU_CAM, 'ICVL': OA.ICVL_CAM, 'MSRA15': OA.MSRA_CAM}[name])
    ds = synthetic.generate(name, 12, seed=52)
    di = ds['importer']
    n = 400
    got = HandDetector.sampleRandomPoses(di, np.random.RandomState(8), ds['gt3Dcrop'], ds['com3D'], ds['cube'], n, MODES,
                                         retall=True)
    want = OA.sample_random_poses(cam, np.random.RandomState(8), ds['gt3Dcrop'], ds['com3D'], ds['cube'], n, MODES,
                                  retall=True)
    assert got[0].shape == (n, ds['gt3Dcrop'].shape[1], 3) and got[0].dtype == np.float32
    print(name, "poses identical:", float((got[0] == want[0]).mean()), "max diff", float(np.abs(got[0] - want[0]).max()))
    assert np.abs(got[0] - want[0]).max() <= 4 * ULP
    np.testing.assert_allclose(got[1], want[1], rtol=2 * ULP)
    np.testing.assert_allclose(got[2], want[2], rtol=2 * ULP)
    assert np.array_equal(got[3], want[3])
    # plain return, a single mode, and the stream position after the call (five array draws, always)
    r1, r2 = np.random.RandomState(9), np.random.RandomState(9)
    a = HandDetector.sampleRandomPoses(di, r1, ds['gt3Dcrop'], ds['com3D'], ds['cube'], 50, ['rot'])
    b = OA.sample_random_poses(cam, r2, ds['gt3Dcrop'], ds['com3D'], ds['cube'], 50, ['rot'])
    assert np.abs(a - b).max() <= 4 * ULP and r1.randint(1 << 30) == r2.randint(1 << 30)
    # aug_modes == ['none'] returns the normalised BASE poses (handdetector.py:843-847)
    a = HandDetector.sampleRandomPoses(di, np.random.RandomState(1), ds['gt3Dcrop'], ds['com3D'], ds['cube'], 50, ['none'])
    assert np.array_equal(a, OA.sample_random_poses(cam, np.random.RandomState(1), ds['gt3Dcrop'], ds['com3D'], ds['cube'],
                                                    50, ['none']))
    with pytest.raises(NotImplementedError):        # the orderings the reference itself cannot sample (:893)
        HandDetector.sampleRandomPoses(di, np.random.RandomState(1), ds['gt3Dcrop'], ds['com3D'], ds['cube'], 50,
                                       ['sc+rot+com'])
    with pytest.raises(NotImplementedError):
        HandDetector.sampleRandomPoses(di, np.random.RandomState(1), ds['gt3Dcrop'], ds['com3D'], ds['cube'], 50, MODES,
                                       rot3D=True)


@pytest.mark.parametrize('name', ['NYU', 'ICVL', 'MSRA15'])
def test_sample_random_poses_matches_reference_fixture(name):
    """tests/golden/reference_pins.npz holds the output of the reference's own sampleRandomPoses (executed under
    NumPy 2, see tests/test_reference_pins.py): float32-scalar promotion differs from NumPy 1.x in the last bits."""
    from data import synthetic
    from util.handdetector import HandDetector
    G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_pins.npz'))
    g = lambda k: G['poses_%s_%s' % (name, k)]
    di = synthetic.make_importer(name)
    r = g('out_poses')
    got = HandDetector.sampleRandomPoses(di, np.random.RandomState(int(g('rng_seed'))), g('base_poses'), g('base_com'),
                                         g('base_cube'), r.shape[0], MODES, retall=True)
    assert np.abs(got[0] - r).max() <= 32 * ULP
    assert np.array_equal(got[1], g('out_com')) and np.array_equal(got[3], g('out_rot'))
