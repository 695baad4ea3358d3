"""GPU parity of PoseRegNet (what the reference's main_*_posereg_embedding.py scripts build,
net/poseregnet.py:60-99) against the oracle: deterministic forward through computeOutput with the
PCA prior layer appended (BASELINE config 1), and a training step with injected dropout masks."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / (np.abs(b).max() + 1e-30))


def test_config1_forward_one_crop_with_pca_prior():
    from net.poseregnet import PoseRegNet, PoseRegNetParams
    from net.hiddenlayer import HiddenLayer, HiddenLayerParams
    from oracle import nets as O
    rng = np.random.RandomState(23455)
    net = PoseRegNet(rng, cfgParams=PoseRegNetParams(type=0, nChan=1, wIn=128, hIn=128, batchSize=1, numJoints=1, nDims=30))
    orng = np.random.RandomState(23455)
    onet = O.build_poseregnet(orng, type=0, batchSize=1, numJoints=1, nDims=30)
    pr = np.random.RandomState(5)
    comp, mean = pr.randn(30, 42).astype('float32'), pr.randn(42).astype('float32')
    cfg = HiddenLayerParams(inputDim=(1, 30), outputDim=(1, 42), activation=None)
    pcalayer = HiddenLayer(rng, net.layers[-1].output, cfg, layerNum=len(net.layers))
    pcalayer.W.set_value(comp)
    pcalayer.b.set_value(mean)
    net.layers.append(pcalayer)
    net.output = pcalayer.output
    net.cfgParams.numJoints, net.cfgParams.nDims, net.cfgParams.outputDim = 14, 3, pcalayer.cfgParams.outputDim
    O.append_pca_layer(onet, comp, mean)
    x = pr.uniform(-1, 1, (1, 1, 128, 128)).astype(np.float32)
    net.setDeterministic()
    out = net.computeOutput(x)
    with torch.no_grad():
        oout, _ = onet.forward(torch.from_numpy(x), deterministic=True)
    assert out.shape == (1, 42)
    assert _rel(out, oout.numpy()) < 1e-4


def test_train_step_with_injected_dropout_masks():
    from net.poseregnet import PoseRegNet, PoseRegNetParams
    from oracle import nets as O
    from dpp_b200.engine import Engine
    B, D = 8, 30
    net = PoseRegNet(np.random.RandomState(7), cfgParams=PoseRegNetParams(type=0, batchSize=B, numJoints=1, nDims=D))
    onet = O.build_poseregnet(np.random.RandomState(7), type=0, batchSize=B, numJoints=1, nDims=D)
    eng = Engine(net)
    net._eng = eng
    pr = np.random.RandomState(3)
    x = pr.uniform(-1, 1, (B, 1, 128, 128)).astype(np.float32)
    y = pr.randn(B, D).astype(np.float32)
    masks = [(pr.rand(B, 1024) < 0.7).astype(np.float32) for _ in range(2)]
    eng.set_dropout_masks(masks)
    adam = O.Adam(onet.params)
    for step in range(2):
        eng.set_input_nchw(x)
        eng.y_in.copy_(torch.from_numpy(y))
        cost = float(eng.train_step(1e-3, use_graph=False).cpu()[0])
        ocost, _, ograds = O.train_step(onet, adam, torch.from_numpy(x), torch.from_numpy(y), 1e-3, 1, D,
                                        masks=[torch.from_numpy(m) for m in masks])
        assert abs(cost - ocost) < 1e-4 * abs(ocost), (cost, ocost)
        g = eng.gradients()
        for p, og in zip(net.params, ograds):
            og = og.numpy()
            assert np.abs(g[id(p)] - og).max() <= 2e-3 * (np.abs(og).max() + 1e-12), p.name
    for p, op_ in zip(net.params, onet.params):
        assert np.abs(p.get_value() - op_.detach().numpy()).max() < 2.5e-3
