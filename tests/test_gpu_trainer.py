"""GPU test of the reference-facing trainer surface end to end (the call sequence of
main_nyu_posereg_embedding.py:96-167): PoseRegNetTrainerParams attributes, setData / addStaticData /
addManagedData / compileFunctions / train, on-device augmentation per epoch from the ORIGINAL crops, validation
observers, save / load and computeOutput with the PCA prior layer appended."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(tmp_path, n_train=80, n_val=32, B=16):
    from sklearn.decomposition import PCA
    from data import synthetic
    from net.resnet import ResNet, ResNetParams
    from trainer.poseregnettrainer import PoseRegNetTrainer, PoseRegNetTrainerParams
    from util.handdetector import HandDetector
    rng = np.random.RandomState(23455)
    ds = synthetic.generate('NYU', n_train, seed=23455)
    dv = synthetic.generate('NYU', n_val, seed=777)
    di = ds['importer']
    pca = PCA(n_components=30)
    g = ds['gt3D']
    pca.fit(np.concatenate([g, g + rng.randn(*g.shape).astype('float32') * 0.03]).reshape(-1, g.shape[1] * 3))
    ty = pca.transform(g.reshape(g.shape[0], -1))
    vy = pca.transform(dv['gt3D'].reshape(dv['gt3D'].shape[0], -1))
    net = ResNet(rng, cfgParams=ResNetParams(type=0, nChan=1, wIn=128, hIn=128, batchSize=B, numJoints=1, nDims=30))
    cfg = PoseRegNetTrainerParams()
    cfg.batch_size = B
    cfg.learning_rate = 0.001
    cfg.weightreg_factor = 0.0
    cfg.force_macrobatch_reload = True
    cfg.para_augment = True
    cfg.para_num_proc = 1
    cfg.validation_frequency = 5
    cfg.augment_fun_params = {'fun': 'augment_poses', 'args': {
        'normZeroOne': False, 'di': di, 'aug_modes': ['com', 'rot', 'none'],
        'hd': HandDetector(ds['x'][0, 0].copy(), abs(di.fx), abs(di.fy), importer=di), 'proj': pca}}
    tr = PoseRegNetTrainer(net, cfg, rng, str(tmp_path))
    tr.setData(ds['x'], ty, dv['x'], vy)
    tr.addStaticData({'val_data_y3D': dv['gt3D']})
    tr.addStaticData({'pca_data': pca.components_, 'mean_data': pca.mean_})
    tr.addManagedData({'train_data_cube': ds['cube'], 'train_data_com': ds['com3D'], 'train_data_M': ds['M'],
                       'train_gt3Dcrop': ds['gt3Dcrop']})
    tr.compileFunctions(compileDebugFcts=False)
    tr.verbose = False
    return net, tr, ds, dv, pca


def test_trainer_runs_epochs_with_device_augmentation(tmp_path):
    from net.resnet import ResNet, ResNetParams
    from net.hiddenlayer import HiddenLayer, HiddenLayerParams
    net, tr, ds, dv, pca = _setup(tmp_path)
    B = tr.cfgParams.batch_size
    x_orig = ds['x'].copy()
    epochs = 3
    train_costs, wvals, val_errs = tr.train(n_epochs=epochs)
    nb = tr.getNumFullMiniBatches()
    assert nb == 80 // B and len(train_costs) == epochs * nb
    assert np.all(np.isfinite(train_costs))
    # training makes progress instead of diverging (15 ADAM steps at lr/10, lr/3, lr*e^-0.12: nettrainer.py:54)
    print("train cost first / last epoch: %.4f / %.4f" % (np.mean(train_costs[:nb]), np.mean(train_costs[-nb:])))
    assert np.mean(train_costs[-nb:]) < 1.25 * np.mean(train_costs[:nb])
    # validation observers ran before training and every validation_frequency minibatches (nettrainer.py:796-870)
    assert len(val_errs) >= 2 and all(len(v) == 1 + (epochs * nb) // 5 for v in val_errs)
    # the stored training crops are never modified: augmentation always starts from the originals (a6)
    assert np.array_equal(tr.train_data_xDB[:80], x_orig)
    # snapshot written by the loop, reference pickle schema (netbase.py:405-477)
    assert os.path.exists(os.path.join(str(tmp_path), 'net_last.pkl'))
    # save -> load into a fresh net -> the same deterministic outputs (a18)
    path = os.path.join(str(tmp_path), 'net.pkl')
    net.save(path)
    out = net.computeOutput(dv['x'][:B])
    cfg2 = ResNetParams(type=0, nChan=1, wIn=128, hIn=128, batchSize=B, numJoints=1, nDims=30)
    cfg2.loadFile = path
    net2 = ResNet(np.random.RandomState(1), cfgParams=cfg2)
    out2 = net2.computeOutput(dv['x'][:B])
    # (not bitwise: the split-K FC GEMM accumulates with floating-point atomics, whose order varies between runs)
    assert np.abs(out - out2).max() < 1e-5 * np.abs(out).max()
    # PCA prior layer appended the way the main script does it (:144-158): joints = emb * components + mean
    cfgp = HiddenLayerParams(inputDim=(B, 30), outputDim=(B, 42), activation=None)
    pl = HiddenLayer(np.random.RandomState(2), net.layers[-1].output, cfgp, layerNum=len(net.layers))
    pl.W.set_value(pca.components_.astype('float32'))
    pl.b.set_value(pca.mean_.astype('float32'))
    net.layers.append(pl)
    net.output = pl.output
    net.cfgParams.numJoints, net.cfgParams.nDims, net.cfgParams.outputDim = 14, 3, pl.cfgParams.outputDim
    j = net.computeOutput(dv['x'][:B])
    assert j.shape == (B, 42)
    ref = out.astype(np.float64) @ pca.components_ + pca.mean_
    assert np.abs(j - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
