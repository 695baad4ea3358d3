"""GPU parity of ScaleNet (reference src/net/scalenet.py:49-193, the CoM-refinement net of the
inference cascade, util/handdetector.py:634-676): three-input forward through the reference-surface class +
libdpp_b200.so against the CPU oracle, deterministic (dropout = 0.7 scale) as ``refineCoM`` runs it."""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / (np.abs(b).max() + 1e-30))


def _inputs(n, seed):
    from data import synthetic
    x0 = synthetic.generate('NYU', n, seed=seed)['x'].astype(np.float32)
    return [x0, np.ascontiguousarray(x0[:, :, 32:96, 32:96]), np.ascontiguousarray(x0[:, :, 48:80, 48:80])]


def test_scalenet_forward_matches_oracle_and_golden(tmp_path):
    from net.scalenet import ScaleNet, ScaleNetParams
    from oracle import nets as O
    B = 4
    cfg = ScaleNetParams(type=1, nChan=1, wIn=128, hIn=128, batchSize=B, numJoints=1, nDims=3)
    assert cfg.numInputs == 3 and [d[2] for d in cfg.inputDim] == [128, 64, 32]
    assert cfg.layers[9].inputDim == (B, 968 + 968 + 512)
    net = ScaleNet(np.random.RandomState(23455), cfgParams=cfg)
    onet = O.build_scalenet(np.random.RandomState(23455), type=1, batchSize=B, numJoints=1, nDims=3)
    xs = _inputs(6, seed=31)                      # 6 samples at batch 4: exercises the last-batch padding
    net.setDeterministic()
    out = net.computeOutput(xs)
    assert out.shape == (6, 3)
    outs = []
    for lo in (0, 2):                             # oracle on samples 0-3 and 2-5
        with torch.no_grad():
            o, _ = onet.forward([torch.from_numpy(x[lo:lo + 4]) for x in xs], deterministic=True)
        outs.append(o.numpy())
    ref = np.concatenate([outs[0], outs[1][2:]])
    r = _rel(out, ref)
    print("ScaleNet forward rel err", r)
    assert r < 1e-4
    # committed golden vector (tests/golden/make_golden.py): same net seed, first two samples
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'scalenet_b2.npz'))
    assert np.array_equal(g['x0'], xs[0][:2])
    assert _rel(out[:2], g['out_det']) < 1e-4
    # pickles keep the reference schema and layouts: save -> load into a fresh net -> same outputs
    path = os.path.join(str(tmp_path), 'scalenet.pkl')
    net.save(path)
    cfg2 = ScaleNetParams(type=1, nChan=1, wIn=128, hIn=128, batchSize=B, numJoints=1, nDims=3)
    cfg2.loadFile = path
    net2 = ScaleNet(np.random.RandomState(7), cfgParams=cfg2)
    net2.setDeterministic()
    assert _rel(net2.computeOutput(xs), out) < 1e-5
    w = net.layers[9].W.get_value()               # FC0 rows come back in the reference (c,h,w)-per-tower order
    assert np.array_equal(w, onet.layers[9].params[0].detach().numpy().astype(np.float32))
