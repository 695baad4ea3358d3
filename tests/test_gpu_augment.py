"""GPU parity of the fused augmentation kernel (dpp_augment_fwd) against the oracle restatement of
NetTrainer.augmentCrop + HandDetector.{moveCoM,rotateHand,scaleHand} (oracle/augment.py), on
seeded synthetic crops, with the random draws made on the host in the reference's order
(nettrainer.py:954-957) and fed to both.  Bar: pixels BIT-EXACT, labels bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CAMS = {'NYU': 'NYU_CAM', 'ICVL': 'ICVL_CAM', 'MSRA15': 'MSRA_CAM'}


def _oracle_and_records(name, n, aug_modes, seed):
    from oracle import augment as A
    from data import synthetic
    ds = synthetic.generate(name, n, seed=seed)
    cam = A.Camera(**getattr(A, CAMS[name]))
    ohd = A.Hand(cam, use_cv2=False)
    rng = np.random.RandomState(seed + 1)
    draws = [A.draw_aug_params(rng, len(aug_modes)) for _ in range(n)]
    ox, oy = A.augment_poses(ds['x'], ds['com3D'], ds['cube'], ds['M'], ds['gt3Dcrop'], list(range(n)), draws,
                             aug_modes, cam, ohd)
    hd, di = ds['hd'], ds['importer']
    recs, labels = [], []
    for i, (mode, off, rot, sc) in enumerate(draws):
        com = di.joint3DToImg(ds['com3D'][i])
        rec, lab, _, _, _ = hd.aug_record(i, aug_modes[mode], off, rot, sc, com, ds['cube'][i].copy(),
                                          ds['M'][i].copy(), ds['gt3Dcrop'][i].copy())
        recs.append(rec)
        labels.append(lab.reshape(-1))
    return ds, np.array(recs), np.stack(labels), ox, oy, draws


@pytest.mark.parametrize("name,aug_modes", [('NYU', ['com', 'rot', 'none']), ('ICVL', ['com', 'rot', 'none']),
                                            ('MSRA15', ['com', 'rot', 'sc', 'none'])])
def test_augment_bit_exact(name, aug_modes):
    from dpp_b200.augment import run_records
    n = 96
    ds, recs, labels, ox, oy, draws = _oracle_and_records(name, n, aug_modes, seed=7)
    out = run_records(ds['x'][:, 0], recs)
    modes = np.array([aug_modes[d[0]] for d in draws])
    for m in set(modes):
        sel = modes == m
        bad = int((out[sel] != ox[sel, 0]).sum())
        print(name, m, "samples", int(sel.sum()), "mismatching pixels", bad)
    assert np.array_equal(out, ox[:, 0])
    assert np.array_equal(labels, oy)
    assert out.min() >= -1.0 - 1e-6 and out.max() <= 1.0 + 1e-6


def test_augment_identity_and_ragged_batch():
    """'none' mode reproduces the stored crop's normalisation; n=1 and n=0 launch cleanly."""
    from dpp_b200.augment import run_records
    ds, recs, labels, ox, oy, draws = _oracle_and_records('NYU', 5, ['none'], seed=3)
    out = run_records(ds['x'][:, 0], recs[:1])
    assert np.array_equal(out[0], ox[0, 0])
    out0 = run_records(ds['x'][:, 0], recs[:0])
    assert out0.shape[0] == 0


def test_handdetector_methods_raw_warp():
    """HandDetector.rotateHand / moveCoM / scaleHand with the reference signatures (raw warp)."""
    from oracle import augment as A
    from data import synthetic
    ds = synthetic.generate('NYU', 4, seed=11)
    cam = A.Camera(**A.NYU_CAM)
    ohd = A.Hand(cam)
    hd, di = ds['hd'], ds['importer']
    i = 1
    com = di.joint3DToImg(ds['com3D'][i])
    cube, M, gt = ds['cube'][i], ds['M'][i], ds['gt3Dcrop'][i]
    dpt = (ds['x'][i, 0] * np.float32(cube[2] / 2.) + com[2]).astype(np.float32)
    a, ja, _ = hd.rotateHand(dpt.copy(), cube, com, 33.3, gt)
    b, jb, _ = ohd.rotateHand(dpt.copy(), cube, com, 33.3, gt)
    assert np.array_equal(a, b) and np.array_equal(ja, jb)
    off = np.array([4., -3., 6.])
    a, ja, ca, Ma = hd.moveCoM(dpt.copy(), cube, com, off, gt, M)
    b, jb, cb, Mb = ohd.moveCoM(dpt.copy(), cube, com, off, gt, M)
    assert np.array_equal(a, b) and np.array_equal(ja, jb) and np.array_equal(ca, cb) and np.array_equal(Ma, Mb)
    a, ja, cua, Ma = hd.scaleHand(dpt.copy(), cube, com, 1.03, gt, M)
    b, jb, cub, Mb = ohd.scaleHand(dpt.copy(), cube, com, 1.03, gt, M)
    assert np.array_equal(a, b) and np.allclose(cua, cub) and np.array_equal(Ma, Mb)
