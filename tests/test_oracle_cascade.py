"""CPU checks of the inference cascade's oracle and host logic:
 * the NN-resize index model against cv2 4.13.0 itself (the pin of oracle/cascade.py),
 * the vectorised host geometry of dpp_b200/cascade.py against the per-sample oracle functions,
 * the ``dpp_crop_rec`` records (interpreted by tests/recrop_model.py exactly as k_recrop does) against the oracle's
   crops - bit-exact, windows leaving the frame and mirroring included,
 * the committed golden vector tests/golden/cascade_nyu.npz."""
import os
import numpy as np
import pytest

from oracle import cascade as OC
from oracle import augment as OA
import recrop_model

cv2 = pytest.importorskip('cv2')
f32, f64 = np.float32, np.float64
FX, FY = 588., 587.            # test_realtimepipeline.py:66 (the detector's focal lengths, not the importer's)


def test_resize_nn_model_matches_cv2():
    rng = np.random.RandomState(0)
    for _ in range(1500):
        Ws, Hs, w, h = rng.randint(1, 480), rng.randint(1, 480), rng.randint(1, 200), rng.randint(1, 200)
        src = rng.rand(Hs, Ws).astype(f32)
        assert np.array_equal(OC.resize_nn(src, (w, h)), OC.resize_nn_cv2(src, (w, h))), (Ws, Hs, w, h)
    for Ws in range(100, 460, 7):                # the cascade's shapes: windows of 100..460 px onto 128 / 127 / 126
        src = rng.rand(Ws + 1, Ws).astype(f32)
        for w, h in ((128, 128), (127, 128), (128, 126)):
            assert np.array_equal(OC.resize_nn(src, (w, h)), OC.resize_nn_cv2(src, (w, h)))


@pytest.mark.parametrize('name,cam', [('NYU', OA.NYU_CAM), ('ICVL', OA.ICVL_CAM), ('MSRA15', OA.MSRA_CAM)])
def test_vectorised_geometry_matches_per_sample_oracle(name, cam):
    from data import synthetic
    from dpp_b200 import cascade as PC
    fr = synthetic.generate_frames(name, 24, seed=5, edge_fraction=0.5)
    di, ocam = fr['importer'], OA.Camera(**cam)
    fx, fy = (FX, FY) if name == 'NYU' else (di.fx, di.fy)
    for coms in (fr['lastcom'], fr['lastcom'].astype(f32)):
        xs, xe, ys, ye, zs, ze = PC.bounds_batch(coms, fr['cube'], fx, fy)
        p3 = PC.img_to_3d_batch(di, coms)
        for i in range(len(coms)):
            assert (xs[i], xe[i], ys[i], ye[i], zs[i], ze[i]) == OC.com_to_bounds(coms[i], fr['cube'], fx, fy)
            assert np.array_equal(p3[i], ocam.jointImgTo3D(coms[i]))
            assert np.array_equal(p3[i], di.jointImgTo3D(coms[i]))
        back = PC.to_img_batch(di, p3)
        for i in range(len(coms)):
            assert np.array_equal(back[i], ocam.joint3DToImg(p3[i]))
    z = np.zeros((1, 3), f32)
    assert np.array_equal(PC.to_img_batch(di, z)[0], ocam.joint3DToImg(z[0]))
    with pytest.raises(ValueError):
        PC.bounds_batch(np.zeros((1, 3)), fr['cube'], fx, fy)


@pytest.mark.parametrize('name,cam', [('NYU', OA.NYU_CAM), ('ICVL', OA.ICVL_CAM)])
def test_crop_records_reproduce_oracle_crops(name, cam):
    from data import synthetic
    from dpp_b200 import cascade as PC
    n = 16
    fr = synthetic.generate_frames(name, n, seed=9, edge_fraction=0.5, nd=0. if name == 'NYU' else 32001.)
    di, ocam, cube, frames = fr['importer'], OA.Camera(**cam), fr['cube'], fr['frames']
    fx, fy = (FX, FY) if name == 'NYU' else (di.fx, di.fy)
    assert (PC.bounds_batch(fr['lastcom'], cube, fx, fy)[0] < 0).any() or \
        (PC.bounds_batch(fr['lastcom'], cube, fx, fy)[1] > frames.shape[2]).any()      # padding is exercised
    for coms in (fr['lastcom'], fr['lastcom'].astype(f32)):
        # stage 1: the refinement net's three inputs
        rec = PC.refine_records(coms, cube, fx, fy, frames.shape[1:])
        x0, x1, x2 = recrop_model.run(frames, rec, centre_crops=True)
        for i in range(n):
            b = OC.com_to_bounds(coms[i], cube, fx, fy)
            for use_cv2 in (False, True):
                rz = (OC.resize_nn_cv2 if use_cv2 else OC.resize_nn)(OC.get_crop(frames[i], *b), (128, 128))
                t = OC.refine_inputs(rz, cube, coms[i])
                assert np.array_equal(t[0][0, 0], x0[i]) and np.array_equal(t[1][0, 0], x1[i])
                assert np.array_equal(t[2][0, 0], x2[i])
        # stage 2: the pose net's input, plain and mirrored
        nd = np.array([PC.nd_value(f) for f in frames], f32)
        assert all(nd[i] == OC.nd_value(frames[i]) for i in range(n))
        for mirror in (False, True):
            rec, M, com3D = PC.pose_records(coms, cube, fx, fy, di, frames.shape[1:], nd, mirror=mirror)
            out = recrop_model.run(frames, rec)
            for i in range(n):
                crop, Mo, c3 = OC.pipeline_detect(frames[i], coms[i], cube, ocam, fx, fy, use_cv2=True)
                assert np.array_equal(crop[:, ::-1] if mirror else crop, out[i])
                assert np.array_equal(Mo, M[i]) and np.array_equal(c3, com3D[i])
        assert (rec['rw'] != 128).any() or (rec['rh'] != 128).any()       # the canvas filler is exercised


def _tiny_fns(seed):
    rng = np.random.RandomState(seed)
    wr = (rng.randn(3, 128 * 128 + 64 * 64 + 32 * 32) * 1e-3).astype(f32)
    wp = (rng.randn(42, 128 * 128) * 1e-3).astype(f32)

    def refine_fn(xs):
        return (wr @ np.concatenate([x.reshape(-1) for x in xs]))[None].astype(f32)

    def pose_fn(x):
        return (wp @ x.reshape(-1))[None].astype(f32)
    return refine_fn, pose_fn


def test_cascade_golden_vector():
    """tests/golden/cascade_nyu.npz (made by tests/golden/make_golden.py with cv2's resize): the oracle's index
    model reproduces it, i.e. the restatement has not drifted from the cv2-backed run that was committed."""
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'cascade_nyu.npz'))
    ocam = OA.Camera(**OA.NYU_CAM)
    refine_fn, pose_fn = _tiny_fns(int(g['fn_seed']))
    for i in range(g['frames'].shape[0]):
        r = OC.cascade_frame(g['frames'][i], g['lastcom'][i], tuple(g['cube']), ocam, FX, FY, refine_fn, pose_fn,
                             right_hand=bool(i % 2), use_cv2=False)
        assert np.array_equal(r['com'], g['com'][i])
        assert np.array_equal(r['crop'], g['crop'][i])
        assert np.array_equal(r['com3D'], g['com3D'][i]) and np.array_equal(r['M'], g['M'][i])
        np.testing.assert_allclose(r['pose'], g['pose'][i], rtol=1e-6, atol=1e-4)


def test_joint_error_reference_formulas():
    """The reductions dpp_joint_errors feeds (handpose_evaluation.py:92-135), restated for the GPU test."""
    rng = np.random.RandomState(1)
    gt, pr = rng.randn(5, 14, 3).astype(f32) * 40, rng.randn(5, 14, 3).astype(f32) * 40
    e = np.sqrt(np.square(gt - pr).sum(axis=2))
    assert e.shape == (5, 14) and np.isclose(np.nanmean(np.nanmean(e, axis=1)), e.mean())


def test_crop_records_with_anisotropic_cubes_and_textured_frames():
    """cubes with different x / y extents (non-square windows: aspect-preserving resize + canvas filler on either
    axis), dense random depth with holes, frames addressed out of order, float64 and float32 CoMs."""
    from data import synthetic
    from dpp_b200 import cascade as PC
    rng = np.random.RandomState(5)
    di, ocam = synthetic.make_importer('NYU'), OA.Camera(**OA.NYU_CAM)
    Hf, Wf = 480, 640
    frames = np.rint(rng.uniform(300, 1200, (4, Hf, Wf))).astype(f32)
    frames[rng.rand(4, Hf, Wf) < 0.3] = 0
    for cube in ((300, 240, 300), (200, 260, 250)):
        c = np.stack([rng.uniform(40, Wf - 40, 12), rng.uniform(40, Hf - 40, 12), rng.uniform(350, 1500, 12)], axis=1)
        idx = rng.randint(0, 4, 12)
        for coms in (c, c.astype(f32)):
            rec, M, c3 = PC.pose_records(coms, cube, FX, FY, di, (Hf, Wf), 0., src_index=idx)
            out = recrop_model.run(frames, rec)
            x0 = recrop_model.run(frames, PC.refine_records(coms, cube, FX, FY, (Hf, Wf), src_index=idx))
            assert (rec['rw'] != rec['rh']).any()
            for i in range(12):
                crop, Mo, c3o = OC.pipeline_detect(frames[idx[i]], coms[i], cube, ocam, FX, FY, ndvalue=0., use_cv2=True)
                b = OC.com_to_bounds(coms[i], cube, FX, FY)
                t = OC.refine_inputs(OC.resize_nn_cv2(OC.get_crop(frames[idx[i]], *b), (128, 128)), cube, coms[i])
                assert np.array_equal(crop, out[i]) and np.array_equal(Mo, M[i]) and np.array_equal(c3o, c3[i])
                assert np.array_equal(t[0][0, 0], x0[i])
    with pytest.raises(ValueError):                   # a window entirely outside the frame is rejected, not guessed
        PC.pose_records(np.array([[-400., 100., 600.]]), (300, 300, 300), FX, FY, di, (Hf, Wf), 0.)
