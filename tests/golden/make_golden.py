"""Generates the committed golden vectors from the CPU oracle (run in the build container:
`python tests/golden/make_golden.py`).  The reference has no tests/fixtures of its own and cannot
run here (Theano), so these oracle outputs - produced from the reference-following restatement in
oracle/ with cv2 4.13.0 as the executable ground truth for the warps - are the pins."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'deep-prior-pp_b200'))


def augment_case(name, cam_name, aug_modes, n, seed):
    from oracle import augment as A
    from data import synthetic
    ds = synthetic.generate(name, n, seed=seed)
    cam = A.Camera(**getattr(A, cam_name))
    rng = np.random.RandomState(seed + 1)
    draws = [A.draw_aug_params(rng, len(aug_modes)) for _ in range(n)]
    ox, oy = A.augment_poses(ds['x'], ds['com3D'], ds['cube'], ds['M'], ds['gt3Dcrop'], list(range(n)), draws,
                             aug_modes, cam, A.Hand(cam, use_cv2=False))
    cx, cy = A.augment_poses(ds['x'], ds['com3D'], ds['cube'], ds['M'], ds['gt3Dcrop'], list(range(n)), draws,
                             aug_modes, cam, A.Hand(cam, use_cv2=True))
    return dict(x=ds['x'], com3D=ds['com3D'], cube=ds['cube'], M=ds['M'], gt3Dcrop=ds['gt3Dcrop'],
                mode=np.array([d[0] for d in draws]), off=np.array([d[1] for d in draws]),
                rot=np.array([d[2] for d in draws]), sc=np.array([d[3] for d in draws]),
                out_x=ox, out_y=oy, cv2_x=cx, cv2_y=cy)


def resnet_case():
    import torch
    from oracle import nets as O
    B, D = 2, 30
    rng = np.random.RandomState(4242)
    x = rng.uniform(-1, 1, (B, 1, 128, 128)).astype(np.float32)
    y = rng.randn(B, D).astype(np.float32)
    net = O.build_resnet(np.random.RandomState(23455), type=0, batchSize=B, numJoints=1, nDims=D)
    adam = O.Adam(net.params)
    with torch.no_grad():
        out_det, _ = net.forward(torch.from_numpy(x), deterministic=True)
    cost, out, grads = O.train_step(net, adam, torch.from_numpy(x), torch.from_numpy(y), 1e-3, 1, D)
    gn = np.array([float(g.norm()) for g in grads], np.float64)
    with torch.no_grad():
        out2, _ = net.forward(torch.from_numpy(x), deterministic=False)
    return dict(x=x, y=y, out_det=out_det.numpy(), out_train=out.numpy(), cost=np.float64(cost), grad_norms=gn,
                fc2_grad=grads[-2].numpy(), out_after_step=out2.numpy())


def scalenet_case():
    """ScaleNet type 1 (scalenet.py:49-193), deterministic forward on a synthetic crop and its centre crops."""
    import torch
    from oracle import nets as O
    from data import synthetic
    B = 2
    ds = synthetic.generate('NYU', B, seed=31)
    x0 = ds['x'].astype(np.float32)
    x1 = np.ascontiguousarray(x0[:, :, 32:96, 32:96])
    x2 = np.ascontiguousarray(x0[:, :, 48:80, 48:80])
    net = O.build_scalenet(np.random.RandomState(23455), type=1, batchSize=B, numJoints=1, nDims=3)
    with torch.no_grad():
        out, _ = net.forward([torch.from_numpy(x0), torch.from_numpy(x1), torch.from_numpy(x2)], deterministic=True)
    return dict(x0=x0, out_det=out.numpy())


def cascade_case():
    """Inference cascade (track -> cropArea3D -> normalise -> pose; oracle/cascade.py) on synthetic NYU frames with
    cv2.resize as the executable ground truth and two tiny linear maps standing in for the nets (the nets have their
    own vectors); odd frames are treated as right hands (mirrored)."""
    from oracle import cascade as OC, augment as OA
    from data import synthetic
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from test_oracle_cascade import _tiny_fns, FX, FY
    n, fn_seed = 6, 77
    fr = synthetic.generate_frames('NYU', n, seed=21, edge_fraction=0.5)
    # the golden file stores the windows only (a 640x480 frame is 1.2 MB): keep frames small by cropping the
    # frame to uint16-representable integer mm and letting npz compress it
    refine_fn, pose_fn = _tiny_fns(fn_seed)
    ocam = OA.Camera(**OA.NYU_CAM)
    res = [OC.cascade_frame(fr['frames'][i], fr['lastcom'][i], fr['cube'], ocam, FX, FY, refine_fn, pose_fn,
                            right_hand=bool(i % 2), use_cv2=True) for i in range(n)]
    return dict(frames=fr['frames'], lastcom=fr['lastcom'], cube=np.array(fr['cube']), fn_seed=np.int64(fn_seed),
                com=np.stack([r['com'] for r in res]), crop=np.stack([r['crop'] for r in res]),
                com3D=np.stack([r['com3D'] for r in res]), M=np.stack([r['M'] for r in res]),
                pose=np.stack([r['pose'] for r in res]))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'cascade':      # add the cascade vector without touching the others
        np.savez_compressed(os.path.join(HERE, 'cascade_nyu.npz'), **cascade_case())
        sys.exit(0)
    np.savez_compressed(os.path.join(HERE, 'augment_nyu.npz'), **augment_case('NYU', 'NYU_CAM', ['com', 'rot', 'none'], 8, 11))
    np.savez_compressed(os.path.join(HERE, 'augment_msra.npz'),
                        **augment_case('MSRA15', 'MSRA_CAM', ['com', 'rot', 'sc', 'none'], 8, 13))
    np.savez_compressed(os.path.join(HERE, 'resnet_b2.npz'), **resnet_case())
    np.savez_compressed(os.path.join(HERE, 'scalenet_b2.npz'), **scalenet_case())
    np.savez_compressed(os.path.join(HERE, 'cascade_nyu.npz'), **cascade_case())
    print("golden vectors written to", HERE)
