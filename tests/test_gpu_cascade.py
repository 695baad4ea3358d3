"""GPU parity of the inference cascade (SURVEY 8f row f1 / BASELINE config 5; reference
src/util/realtimehandposepipeline.py:296-370, src/util/handdetector.py:204-296,336-351,382-490,511-533,634-676)
and of the pose error metrics (row f4; src/util/handpose_evaluation.py:92-181).

 * ``dpp_recrop_fwd`` against the oracle's crops: BIT-EXACT (integer index work + float32 compares / one subtract
   and one divide), windows leaving the frame, the canvas filler and mirroring included;
 * the whole cascade (ScaleNet -> re-crop -> ResNet type 1) through ``RealtimeHandposePipeline.processBatch`` against
   the oracle cascade with the oracle nets: refined CoM and normalised pose within 1e-4 relative (the tolerance
   north_star states for floating point), crops bit-exact given the same refined CoM;
 * the single-frame reference-signature surface (detect / estimatePose) against the batched one."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
f32, f64 = np.float32, np.float64
FX, FY = 588., 587.


def _rel(a, b):
    return float(np.abs(np.asarray(a, f64) - np.asarray(b, f64)).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize('name', ['NYU', 'ICVL', 'MSRA15'])
def test_recrop_kernel_bit_exact(name):
    from data import synthetic
    from dpp_b200 import cascade as PC
    from oracle import cascade as OC, augment as OA
    cam = {'NYU': OA.NYU_CAM, 'ICVL': OA.ICVL_CAM, 'MSRA15': OA.MSRA_CAM}[name]
    n = 24
    fr = synthetic.generate_frames(name, n, seed=9, edge_fraction=0.5, nd=0. if name != 'ICVL' else 32001.)
    di, ocam, cube, frames = fr['importer'], OA.Camera(**cam), fr['cube'], fr['frames']
    fx, fy = (FX, FY) if name == 'NYU' else (di.fx, di.fy)
    dev = torch.from_numpy(frames).cuda()
    for coms in (fr['lastcom'], fr['lastcom'].astype(f32)):
        rec = PC.refine_records(coms, cube, fx, fy, frames.shape[1:])
        x0 = torch.full((n, 128, 128), -7., device='cuda')
        x1 = torch.full((n, 64, 64), -7., device='cuda')
        x2 = torch.full((n, 32, 32), -7., device='cuda')
        PC.run_crop_records(dev, rec, x0, x1, x2)
        x0, x1, x2 = x0.cpu().numpy(), x1.cpu().numpy(), x2.cpu().numpy()
        for i in range(n):
            b = OC.com_to_bounds(coms[i], cube, fx, fy)
            t = OC.refine_inputs(OC.resize_nn_cv2(OC.get_crop(frames[i], *b), (128, 128)), cube, coms[i])
            assert np.array_equal(t[0][0, 0], x0[i]), (name, i)
            assert np.array_equal(t[1][0, 0], x1[i]) and np.array_equal(t[2][0, 0], x2[i]), (name, i)
        nd = np.array([PC.nd_value(f) for f in frames], f32)
        for mirror in (False, True):
            rec, M, com3D = PC.pose_records(coms, cube, fx, fy, di, frames.shape[1:], nd, mirror=mirror)
            out = torch.full((n, 128, 128), -7., device='cuda')
            PC.run_crop_records(dev, rec, out)
            out = out.cpu().numpy()
            for i in range(n):
                crop, Mo, c3 = OC.pipeline_detect(frames[i], coms[i], cube, ocam, fx, fy, use_cv2=True)
                assert np.array_equal(crop[:, ::-1] if mirror else crop, out[i]), (name, i, mirror)
                assert np.array_equal(Mo, M[i]) and np.array_equal(c3, com3D[i])
    # records may address any frame in any order, and an empty batch is a no-op
    rec, _, _ = PC.pose_records(fr['lastcom'][::-1], cube, fx, fy, di, frames.shape[1:], 0., src_index=np.arange(n)[::-1])
    out = torch.empty((n, 128, 128), device='cuda')
    PC.run_crop_records(dev, rec, out)
    crop, _, _ = OC.pipeline_detect(frames[n - 1], fr['lastcom'][n - 1], cube, ocam, fx, fy, ndvalue=0., use_cv2=True)
    assert np.array_equal(out[0].cpu().numpy(), crop)
    PC.run_crop_records(dev, rec[:0], torch.empty((0, 128, 128), device='cuda'))
    with pytest.raises(ValueError):
        bad = rec.copy()
        bad['src_index'][0] = n
        PC.run_crop_records(dev, bad, out)


def _build_nets(B, J):
    from net.resnet import ResNet, ResNetParams
    from net.scalenet import ScaleNet, ScaleNetParams
    from oracle import nets as ON
    pose = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=1, nChan=1, wIn=128, hIn=128, batchSize=B,
                                                                      numJoints=J, nDims=3))
    ref = ScaleNet(np.random.RandomState(23455), cfgParams=ScaleNetParams(type=1, nChan=1, wIn=128, hIn=128, batchSize=B,
                                                                         resizeFactor=2, numJoints=1, nDims=3))
    opose = ON.build_resnet(np.random.RandomState(23455), type=1, batchSize=B, numJoints=J, nDims=3)
    oref = ON.build_scalenet(np.random.RandomState(23455), type=1, batchSize=B, numJoints=1, nDims=3)
    return pose, ref, opose, oref


def test_cascade_matches_oracle():
    from data import synthetic
    from util.realtimehandposepipeline import RealtimeHandposePipeline
    from oracle import cascade as OC, augment as OA
    B, J, n = 4, 14, 6
    fr = synthetic.generate_frames('NYU', n, seed=33, edge_fraction=0.34)
    di, ocam, cube, frames = fr['importer'], OA.Camera(**OA.NYU_CAM), fr['cube'], fr['frames']
    pose, ref, opose, oref = _build_nets(B, J)
    config = {'fx': FX, 'fy': FY, 'cube': cube}
    rtp = RealtimeHandposePipeline(pose, config, di, verbose=False, comrefNet=ref)
    rtp.initNets()

    def refine_fn(xs):            # one frame through the oracle ScaleNet (deterministic; batch padded by repetition)
        with torch.no_grad():
            o, _ = oref.forward([torch.from_numpy(np.repeat(x, B, axis=0)) for x in xs], deterministic=True)
        return o.numpy()[:1]

    def pose_fn(x):
        with torch.no_grad():
            o, _ = opose.forward(torch.from_numpy(np.repeat(x, B, axis=0)), deterministic=True)
        return o.numpy()[:1]

    for hand in (RealtimeHandposePipeline.HAND_LEFT, RealtimeHandposePipeline.HAND_RIGHT):
        rtp.hand = hand
        right = hand == RealtimeHandposePipeline.HAND_RIGHT
        got = [rtp.processBatch(frames[lo:lo + B], fr['lastcom'][lo:lo + B]) for lo in (0, B)]   # 4 + 2: padded batch
        for key in ('pose', 'pose_norm', 'com', 'com3D', 'M'):
            got[0][key] = np.concatenate([got[0][key], got[1][key]])
        got = got[0]
        assert got['pose'].shape == (n, J, 3)
        for i in range(n):
            want = OC.cascade_frame(frames[i], fr['lastcom'][i], cube, ocam, FX, FY, refine_fn, pose_fn,
                                    right_hand=right, use_cv2=True)
            # refined CoM: (u, v) a few hundred px, d several hundred mm -> 1e-4 relative of each coordinate
            assert np.all(np.abs(got['com'][i] - want['com']) <= 1e-4 * np.abs(want['com'])), (i, got['com'][i], want['com'])
            # same CoM in -> bit-identical crop (what the device crop kernel produced for ITS refined CoM)
            crop, M, c3 = OC.pipeline_detect(frames[i], got['com'][i], cube, ocam, FX, FY, use_cv2=True)
            assert np.array_equal(M, got['M'][i]) and np.array_equal(c3, got['com3D'][i])
            jj = OC.estimate_pose(crop, pose_fn, right)
            r = _rel(got['pose_norm'][i], jj)
            print("frame", i, "hand", hand, "pose rel err", r)
            assert r < 1e-4
            pose_mm = jj * f32(cube[2] / 2.) + c3
            assert np.abs(got['pose'][i] - pose_mm).max() < 1e-4 * np.abs(pose_mm).max()

    # the crop the device fed to the pose net (last batch, last call) is the oracle's, bit for bit
    res = rtp._cascade.run(frames[:B], fr['lastcom'][:B], right_hand=False, return_crops=True)
    for i in range(B):
        crop, _, _ = OC.pipeline_detect(frames[i], res['com'][i], cube, ocam, FX, FY, use_cv2=True)
        assert np.array_equal(crop, res['crop'][i])

    # single-frame reference surface == batched surface
    rtp.hand = RealtimeHandposePipeline.HAND_LEFT
    rtp.lastcom = fr['lastcom'][1]
    one = rtp.processFrame(frames[1])
    assert _rel(one, res['pose'][1]) < 1e-4
    # device-resident frames need the sensor's undefined-depth value
    with pytest.raises(ValueError):
        rtp.processBatch(torch.from_numpy(frames[:B]).cuda(), fr['lastcom'][:B])
    res2 = rtp.processBatch(torch.from_numpy(frames[:B]).cuda(), fr['lastcom'][:B], ndvalue=0.)
    assert _rel(res2['pose'], res['pose']) < 1e-5 and np.array_equal(res2['M'], res['M'])


def test_joint_errors_match_reference_formulas():
    from dpp_b200.cascade import joint_errors
    rng = np.random.RandomState(3)
    for n, J in ((7, 14), (3, 21), (1, 1), (5, 40)):
        gt = (rng.randn(n, J, 3) * 40).astype(f32)
        pr = (gt + rng.randn(n, J, 3) * 9).astype(f32)
        if J > 2:
            gt[0, 1] = np.nan                   # joints without annotation are NaN (nanmean / nanmax)
        e = np.sqrt(np.square(gt - pr).sum(axis=2))
        err, fmean, fmax = [t.cpu().numpy() for t in joint_errors(pr, gt)]
        np.testing.assert_allclose(err, e, rtol=1e-6, atol=0, equal_nan=True)
        np.testing.assert_allclose(fmean, np.nanmean(e, axis=1), rtol=1e-5)
        np.testing.assert_allclose(fmax, np.nanmax(e, axis=1), rtol=1e-6)
        # handpose_evaluation.py:98, :127: getMeanError / getMaxError
        assert np.isclose(fmean.mean(), np.nanmean(np.nanmean(e, axis=1)), rtol=1e-5)
        assert np.isclose(fmax.max(), np.nanmax(e), rtol=1e-6)
    gt = np.full((2, 3, 3), np.nan, f32)
    _, fmean, fmax = [t.cpu().numpy() for t in joint_errors(np.zeros((2, 3, 3), f32), gt)]
    assert np.isnan(fmean).all() and np.isnan(fmax).all()


def test_handpose_evaluation_metrics():
    """util/handpose_evaluation.py surface (reference src/util/handpose_evaluation.py:92-228) against the reference's
    NumPy one-liners; float32 on the device vs float64-accumulating NumPy -> 1e-5 relative."""
    import scipy.stats  # noqa: F401
    from util.handpose_evaluation import HandposeEvaluation
    rng = np.random.RandomState(11)
    n, J = 40, 14
    gt = (rng.randn(n, J, 3) * 40).astype(f32)
    pr = (gt + rng.randn(n, J, 3) * 9).astype(f32)
    gt[3, 2] = np.nan
    gt[7] = np.nan                                          # a frame without annotation
    ev = HandposeEvaluation(list(gt), list(pr))
    e = np.sqrt(np.square(gt - pr).sum(axis=2))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        tol = dict(rtol=1e-5)
        assert np.isclose(ev.getMeanError(), np.nanmean(np.nanmean(e, axis=1)), **tol)
        assert np.isclose(ev.getStdError(), np.nanmean(np.nanstd(e, axis=1)), **tol)
        np.testing.assert_allclose(ev.getMeanErrorOverSeq(), np.nanmean(e, axis=1), **tol)
        assert np.isclose(ev.getMedianError(), np.nanmedian(e), **tol)
        assert np.isclose(ev.getMaxError(), np.nanmax(e), **tol)
        np.testing.assert_allclose(ev.getMaxErrorOverSeq(), np.nanmax(e, axis=1), **tol)
        for j in (0, 2, 13):
            assert np.isclose(ev.getJointMeanError(j), np.nanmean(e[:, j]), **tol)
            assert np.isclose(ev.getJointStdError(j), np.nanstd(e[:, j]), **tol)
            assert np.isclose(ev.getJointMaxError(j), np.nanmax(e[:, j]), **tol)
            np.testing.assert_allclose(ev.getJointErrorOverSeq(j), e[:, j], rtol=1e-6)
            assert np.array_equal(ev.getJointDiffOverSeq(j), (gt - pr)[:, j], equal_nan=True)
        for dist in (10., 20., 40.):
            assert ev.getNumFramesWithinMaxDist(dist) == (np.nanmax(e, axis=1) <= dist).sum()
            assert ev.getNumFramesWithinMeanDist(dist) == (np.nanmean(e, axis=1) <= dist).sum()
            assert ev.getNumFramesWithinMedianDist(dist) == (np.median(e, axis=1) <= dist).sum()
            assert ev.getJointNumFramesWithinMaxDist(dist, 5) == (e[:, 5] <= dist).sum()
    with pytest.raises(ValueError):
        HandposeEvaluation(list(gt), list(pr[:-1]))


@pytest.mark.parametrize('name', ['NYU', 'MSRA15'])
def test_dataset_stack_on_device(name):
    """data.dataset.Dataset.imgStackDepthOnly (reference src/data/dataset.py:75-111) through dpp_recrop_fwd: the
    stack equals the host restatement in data.synthetic.generate - itself pinned to the reference's Dataset code,
    tests/test_reference_pins.py - bit for bit."""
    from data import synthetic
    from data import dataset as D
    seq = synthetic.generate_sequence(name, 9, seed=43)
    ds = synthetic.generate(name, 9, seed=43)
    cls = {'NYU': D.NYUDataset, 'MSRA15': D.MSRA15Dataset}[name]
    dset = cls([seq])
    img, lab = dset.imgStackDepthOnly('train')
    assert img.shape == (9, 1, 128, 128) and img.dtype == np.float32
    assert np.array_equal(img, ds['x']) and np.array_equal(lab, ds['gt3D'])
    assert dset.imgStackDepthOnly('train')[0] is img              # localCache
    assert dset.imgStackDepthOnly('nope') == [] and dset.imgSeq('train') is seq
    with pytest.raises(NotImplementedError):
        D.Dataset([seq], localCache=False).imgStackDepthOnly('train', normZeroOne=True)
