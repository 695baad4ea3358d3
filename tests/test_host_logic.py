"""CPU tests of the host side: reference-surface classes, layer bookkeeping, checkpoints,
trainer batch arithmetic, the C-ABI library's exported symbols (no compute without a GPU)."""
import ctypes
import os
import re
import numpy as np
import pytest

from net.resnet import ResNet, ResNetParams
from net.poseregnet import PoseRegNet, PoseRegNetParams
from net.hiddenlayer import HiddenLayer, HiddenLayerParams
from net.batchnormlayer import BatchNormLayer
from net.convlayer import ConvLayer, ConvLayerParams
from net.convpoollayer import ConvPoolLayerParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_resnet_layer_numbering_and_param_order():
    net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=2, numJoints=1, nDims=30))
    assert len(net.layers) == 189
    assert [l.layerNum for l in net.layers] == list(range(189))
    assert sum(isinstance(l, BatchNormLayer) for l in net.layers) == 61
    assert sum(isinstance(l, ConvLayer) for l in net.layers) == 63
    # projection block = 10 layers, identity = 9; stage-1 block 0 starts at layer 1
    assert isinstance(net.layers[10], ConvLayer) and net.layers[10].cfgParams.stride == (2, 2)   # shortcut conv
    assert net.layers[10].cfgParams.inputDim == (2, 32, 64, 64)
    assert net.layers[-1].cfgParams.outputDim == (2, 30)
    n = sum(int(np.prod(p.get_value().shape)) for p in net.params)
    assert n == 18713150
    names = [p.name for p in net.params[:6]]
    assert names == ['convW0', 'convB0', 'beta1', 'gamma1', 'convW3', 'convB3']
    assert ResNet(np.random.RandomState(1), cfgParams=ResNetParams(type=1, batchSize=2, numJoints=14, nDims=3)).layers[-1] \
        .cfgParams.outputDim == (2, 42)


def test_stage4_ignores_stride_like_the_reference():
    net = ResNet(np.random.RandomState(0), cfgParams=ResNetParams(type=0, batchSize=1, numJoints=1, nDims=30))
    convs = [l for l in net.layers if isinstance(l, ConvLayer)]
    assert convs[-1].cfgParams.outputDim == (1, 256, 8, 8)      # stage 4 stays at 8x8 (resnet.py:353)


def test_init_matches_oracle_draw_for_draw():
    from oracle import nets as O
    net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=4, batchSize=2, numJoints=14, nDims=3))
    onet = O.build_resnet(np.random.RandomState(23455), type=4, batchSize=2, numJoints=14, nDims=3)
    assert len(net.params) == len(onet.params)
    for a, b in zip(net.params, onet.params):
        assert np.array_equal(a.get_value(), b.detach().numpy())
    p = PoseRegNet(np.random.RandomState(5), cfgParams=PoseRegNetParams(type=11, batchSize=2, numJoints=1, nDims=30))
    op = O.build_poseregnet(np.random.RandomState(5), type=11, batchSize=2, numJoints=1, nDims=30)
    for a, b in zip(p.params, op.params):
        assert np.array_equal(a.get_value(), b.detach().numpy())


def test_layer_dims():
    p = ConvPoolLayerParams(inputDim=(4, 1, 128, 128), nFilters=8, filterDim=(5, 5), poolsize=(4, 4))
    assert p.outputDim == (4, 8, 31, 31)
    c = ConvLayerParams(inputDim=(4, 32, 64, 64), nFilters=16, filterDim=(1, 1), stride=(2, 2), border_mode='same')
    assert c.outputDim == (4, 16, 32, 32) and c.border_mode == 'half'
    pn = PoseRegNetParams(type=0, batchSize=4, numJoints=1, nDims=30)
    assert [l.outputDim for l in pn.layers[:4]] == [(4, 8, 31, 31), (4, 8, 13, 13), (4, 8, 11, 11), (4, 1024)]


def test_save_load_roundtrip_and_pca_layer_append(tmp_path):
    rng = np.random.RandomState(3)
    net = PoseRegNet(rng, cfgParams=PoseRegNetParams(type=0, batchSize=2, numJoints=1, nDims=30))
    f = str(tmp_path / 'net.pkl')
    net.save(f)
    net2 = PoseRegNet(np.random.RandomState(99), cfgParams=PoseRegNetParams(type=0, batchSize=2, numJoints=1, nDims=30))
    assert not np.array_equal(net.layers[0].W.get_value(), net2.layers[0].W.get_value())
    net2.load(f)
    for a, b in zip(net.params, net2.params):
        assert np.array_equal(a.get_value(), b.get_value())
    # main_nyu_posereg_embedding.py:148-158
    pca_w, pca_b = rng.randn(30, 42).astype('float32'), rng.randn(42).astype('float32')
    cfg = HiddenLayerParams(inputDim=(2, 30), outputDim=(2, 42), activation=None)
    pcalayer = HiddenLayer(rng, net.layers[-1].output, cfg, layerNum=len(net.layers))
    pcalayer.W.set_value(pca_w)
    pcalayer.b.set_value(pca_b)
    net.layers.append(pcalayer)
    net.output = pcalayer.output
    net.cfgParams.numJoints, net.cfgParams.nDims, net.cfgParams.outputDim = 14, 3, pcalayer.cfgParams.outputDim
    assert len(net.params) == 14 and np.array_equal(net.params[-2].get_value(), pca_w)
    net.save(str(tmp_path / 'net_prior.pkl.gz'))


def test_deterministic_switch_and_dropout():
    net = PoseRegNet(np.random.RandomState(3), cfgParams=PoseRegNetParams(type=0, batchSize=2, numJoints=1, nDims=30))
    assert net.hasDropout() and not net.isDeterministic()
    net.setDeterministic()
    assert net.isDeterministic()
    net.unsetDeterministic()
    assert not net.isDeterministic()
    r = ResNet(np.random.RandomState(0), cfgParams=ResNetParams(type=0, batchSize=1, numJoints=1, nDims=30))
    assert not r.hasDropout()


def test_trainer_batch_arithmetic_and_padding():
    from trainer.poseregnettrainer import PoseRegNetTrainer, PoseRegNetTrainerParams
    net = PoseRegNet(np.random.RandomState(3), cfgParams=PoseRegNetParams(type=0, batchSize=8, numJoints=1, nDims=30))
    cfg = PoseRegNetTrainerParams()
    cfg.batch_size = 8
    cfg.weightreg_factor = 0.0
    tr = PoseRegNetTrainer(net, cfg, np.random.RandomState(1), '/tmp')
    N = 21
    x = np.arange(N * 4, dtype='float32').reshape(N, 1, 2, 2)
    y = np.arange(N * 30, dtype='float32').reshape(N, 30)
    tr.setData(x, y, x[:16], y[:16])
    assert tr.getNumMiniBatches() == 3 and tr.getNumMacroBatches() == 1 and tr.getNumSamplesPerMacroBatch() == 24
    assert tr.train_data_xDB.shape[0] == 24
    rng = np.random.RandomState(N)                        # nettrainer.py:401-407
    for i in range(3):
        j = rng.randint(0, N)
        assert np.array_equal(tr.train_data_xDB[N + i], x[j])
    # x and y padded with the same draws
    rng = np.random.RandomState(N)
    for i in range(3):
        assert np.array_equal(tr.train_data_yDB[N + i], y[rng.randint(0, N)])
    lr = cfg.lr_of_ep
    cfg.learning_rate = 1e-3
    assert lr(1) == np.float32(1e-4) and lr(2) == np.float32(1e-3 / 3.) and abs(lr(10) - 1e-3 * np.exp(-0.4)) < 1e-9


def test_c_abi_exports_every_declared_symbol():
    from dpp_b200.lib import library_path, EXPORTED_SYMBOLS
    hdr = open(os.path.join(ROOT, 'include', 'dpp_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(dpp_[a-z0-9_]+)\s*\(', hdr)))
    assert declared, "no declarations found"
    dll = ctypes.CDLL(library_path())
    for name in declared:
        assert hasattr(dll, name), name
    assert set(EXPORTED_SYMBOLS) == set(declared)
    dll.dpp_abi_version.restype = ctypes.c_int
    assert dll.dpp_abi_version() == 2


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dpp_b200 import DppError
    net = PoseRegNet(np.random.RandomState(3), cfgParams=PoseRegNetParams(type=0, batchSize=2, numJoints=1, nDims=30))
    with pytest.raises(DppError):
        net.computeOutput(np.zeros((2, 1, 128, 128), 'float32'))


def test_aug_record_layout_matches_c_struct():
    from dpp_b200.lib import AUG_REC_DTYPE, AugRec
    dt = np.dtype(AUG_REC_DTYPE)
    assert dt.itemsize == ctypes.sizeof(AugRec) == 112
    for name in dt.names:
        assert dt.fields[name][1] == getattr(AugRec, name).offset


def test_host_augmentation_geometry_matches_oracle():
    """labels, matrices and thresholds of the dpp_aug_rec records vs the oracle (pixels are the
    GPU test's job)"""
    from oracle import augment as A
    from data import synthetic
    ds = synthetic.generate('MSRA15', 40, seed=5)
    cam = A.Camera(**A.MSRA_CAM)
    ohd = A.Hand(cam)
    hd, di = ds['hd'], ds['importer']
    modes = ['com', 'rot', 'sc', 'none']
    rng = np.random.RandomState(2)
    for i in range(40):
        mode, off, rot, sc = A.draw_aug_params(rng, 4)
        com = di.joint3DToImg(ds['com3D'][i])
        assert np.array_equal(com, cam.joint3DToImg(ds['com3D'][i]))
        rec, lab, cube2, com2, M2 = hd.aug_record(i, modes[mode], off, rot, sc, com, ds['cube'][i].copy(),
                                                  ds['M'][i].copy(), ds['gt3Dcrop'][i].copy())
        _, olab, ocube, ocom, oM = A.augment_crop(ds['x'][i, 0], ds['gt3Dcrop'][i].copy(), com, ds['cube'][i].copy(),
                                                   ds['M'][i].copy(), modes[mode], off, rot, sc, ohd)
        assert np.array_equal(lab, olab)
        assert np.allclose(cube2, ocube, rtol=0, atol=0) and np.array_equal(com2, ocom) and np.array_equal(M2, oM)


def test_aug_records_batch_is_bit_identical_to_per_sample_records():
    """HandDetector.aug_records_batch (the vectorised host path the trainer and bench.py use) against aug_record."""
    from data import synthetic
    for name in ('NYU', 'ICVL', 'MSRA15'):
        n = 240
        ds = synthetic.generate(name, 32, seed=5)
        hd, di = ds['hd'], ds['importer']
        rng = np.random.RandomState(3)
        modes = ['com', 'rot', 'sc', 'none']
        idxs = rng.randint(0, 32, n)
        md = rng.randint(0, 4, n)
        off, rot, sc = rng.randn(n, 3) * 5., rng.uniform(-180, 180, n), np.abs(1. + rng.randn(n) * 0.02)
        off[5], rot[6], sc[7] = 0., 0., 1.               # the reference's early-outs (allclose checks)
        md[5], md[6], md[7] = 0, 1, 2
        com = np.stack([di.joint3DToImg(ds['com3D'][i]) for i in idxs])
        assert np.array_equal(hd._toimg(ds['com3D'][idxs]), com)
        assert np.array_equal(hd._to3d(com), np.stack([di.jointImgTo3D(c) for c in com]))
        per = [hd.aug_record(i, modes[md[k]], off[k], rot[k], sc[k], com[k], ds['cube'][i].copy(), ds['M'][i].copy(),
                             ds['gt3Dcrop'][i].copy()) for k, i in enumerate(idxs)]
        rec, lab = hd.aug_records_batch(idxs, [modes[m] for m in md], off, rot, sc, com, ds['cube'][idxs], ds['M'][idxs],
                                        ds['gt3Dcrop'][idxs])
        assert rec.tobytes() == np.array([p[0] for p in per]).tobytes(), name
        assert np.array_equal(lab, np.stack([p[1] for p in per])), name
        assert set(rec['mode']) == {0, 1, 2}
    with pytest.raises(NotImplementedError):
        hd.aug_records_batch(idxs[:1], ['comb'], off[:1], rot[:1], sc[:1], com[:1], ds['cube'][idxs[:1]], ds['M'][idxs[:1]],
                             ds['gt3Dcrop'][idxs[:1]])


def test_trainer_record_worker_matches_per_sample_path():
    """trainer/nettrainer.py::_records_chunk (vectorised) against the per-sample records + sklearn-style projection."""
    from data import synthetic
    from trainer.nettrainer import _records_chunk
    ds = synthetic.generate('NYU', 16, seed=9)
    hd, di = ds['hd'], ds['importer']
    comp, mean = synthetic.random_orthonormal_pca(30, 42, seed=1)

    class Proj(object):
        def transform(self, X):
            return np.dot(X - mean, comp.T)
    rng = np.random.RandomState(1)
    modes = ['com', 'rot', 'none']
    idxs = list(range(16))
    draws = [(rng.randint(0, 3), rng.randn(3) * 5., rng.uniform(-180, 180), abs(1. + rng.randn() * 0.02)) for _ in idxs]
    recs, labels = _records_chunk(((hd, di, modes, ds['com3D'], ds['cube'], ds['M'], ds['gt3Dcrop'], Proj()), idxs, draws))
    for k, i in enumerate(idxs):
        r, lab, _, _, _ = hd.aug_record(i, modes[draws[k][0]], draws[k][1], draws[k][2], draws[k][3],
                                        di.joint3DToImg(ds['com3D'][i]), ds['cube'][i].copy(), ds['M'][i].copy(),
                                        ds['gt3Dcrop'][i].copy())
        assert recs[k].tobytes() == r.tobytes()
        np.testing.assert_allclose(labels[k], Proj().transform(lab.reshape(1, -1))[0], rtol=1e-6, atol=1e-6)
    assert labels.dtype == np.float32 and labels.shape == (16, 30)
    empty = _records_chunk(((hd, di, modes, ds['com3D'], ds['cube'], ds['M'], ds['gt3Dcrop'], None), [], []))
    assert len(empty[0]) == 0
    # a worker that replays the epoch's draw sequence from the generator state builds the same records for its rows
    # and reports the state the sequence leaves behind (NetTrainer._request_augmentation / _swap_augmentation)
    rng2 = np.random.RandomState(1)
    rows = [3, 4, 5, 9, 15]
    state = (hd, di, modes, ds['com3D'][rows], ds['cube'][rows], ds['M'][rows], ds['gt3Dcrop'][rows], Proj())
    out = _records_chunk((state, list(range(len(rows))), ('replay', rng2.get_state(), 16, (3, 5., 0.02, 180.), rows), rows))
    assert len(out) == 3
    for k, i in enumerate(rows):
        assert out[0][k].tobytes() == recs[i].tobytes() and (out[1][k] == labels[i]).all()
    a, b = out[2], rng.get_state()
    assert a[0] == b[0] and (a[1] == b[1]).all() and a[2:] == b[2:]


def test_importers_load_synthetic_sequences_with_reference_signature(monkeypatch, capsys):
    """data/importers.py ``loadSequence`` (reference importers.py:233, :597, :943 signatures) without dataset files."""
    from data.importers import NYUImporter, ICVLImporter, MSRA15Importer
    from data.basetypes import NamedImgSequence, DepthFrame
    rng = np.random.RandomState(23455)
    di = NYUImporter('../data/NYU/', refineNet=None)
    # a dataset directory is named but its readers are not part of the package: refuse unless the caller opts in
    monkeypatch.delenv('DPP_SYNTHETIC', raising=False)
    with pytest.raises(NotImplementedError, match="DPP_SYNTHETIC=1"):
        di.loadSequence('train', Nmax=6)
    monkeypatch.setenv('DPP_SYNTHETIC', '1')
    s1 = di.loadSequence('train', Nmax=6, shuffle=True, rng=rng, docom=False)
    assert "SYNTHETIC" in capsys.readouterr().out               # ... and says so on every call
    s2 = di.loadSequence('test_1', Nmax=4, docom=False)
    assert isinstance(s1, NamedImgSequence) and s1.name == 'train' and len(s1.data) == 6 and len(s2.data) == 4
    assert s1.config == {'cube': (300, 300, 300)} and isinstance(s1.data[0], DepthFrame)
    f = s1.data[0]
    assert f.dpt.shape == (128, 128) and f.dpt.dtype == np.float32 and f.gt3Dcrop.shape == (14, 3)
    assert f.T.shape == (3, 3) and f.com.shape == (3,) and f.gtorig.shape == (14, 3) and f.gtcrop.shape == (14, 3)
    assert np.allclose(f.gt3Dorig, f.gt3Dcrop + f.com)
    again = NYUImporter(None).loadSequence('train', Nmax=6)           # deterministic per (dataset, sequence)
    assert sorted(float(x.com[2]) for x in again.data) == sorted(float(x.com[2]) for x in s1.data)
    assert [float(x.com[2]) for x in again.data] != [float(x.com[2]) for x in s1.data]        # s1 was shuffled
    assert len(ICVLImporter(None).loadSequence('train', ['0'], Nmax=3, cube=(200, 200, 200)).data) == 3
    assert ICVLImporter(None).loadSequence('test_seq_1', Nmax=2).data[0].gt3Dcrop.shape == (16, 3)
    m = MSRA15Importer(None).loadSequence('P3', Nmax=2)
    assert m.config['cube'] == (180, 180, 180) and m.data[0].gt3Dcrop.shape == (21, 3)
    with pytest.raises(TypeError):
        ICVLImporter(None).loadSequence('train', subSeq='0')
    with pytest.raises(KeyError):
        NYUImporter(None).loadSequence('no_such_sequence')


def test_aug_records_batch_edge_cases_match_per_sample_path():
    """offsets / angles / scales at the reference's ``allclose`` early-outs (1e-9 offsets, 0 / 360 / 1e-7 degrees, scale
    1 +- 1e-10), large offsets, both rotation directions: the vectorised path must stay bit-identical."""
    from data import synthetic
    modes = ['com', 'rot', 'sc', 'none']
    for name in ('NYU', 'MSRA15'):
        ds = synthetic.generate(name, 32, seed=99)
        hd = ds['hd']
        rng = np.random.RandomState(17)
        for rep in range(3):
            n = 300
            idxs, md = rng.randint(0, 32, n), rng.randint(0, 4, n)
            off = rng.randn(n, 3) * rng.choice([1e-9, 0.5, 5., 40.], n)[:, None]
            rot = rng.uniform(-180, 180, n) * rng.choice([1e-9, 1e-3, 1., 1.], n)
            rot[rng.rand(n) < 0.05] = rng.choice([0., 360., -360., 180., 90., 1e-7])
            sc = np.abs(1. + rng.randn(n) * rng.choice([1e-10, 0.02, 0.2], n))
            sc[rng.rand(n) < 0.05] = 1.0
            com = hd._toimg(ds['com3D'][idxs])
            per = [hd.aug_record(i, modes[md[k]], off[k], rot[k], sc[k], com[k], ds['cube'][i].copy(), ds['M'][i].copy(),
                                 ds['gt3Dcrop'][i].copy()) for k, i in enumerate(idxs)]
            rec, lab = hd.aug_records_batch(idxs, [modes[m] for m in md], off, rot, sc, com, ds['cube'][idxs],
                                            ds['M'][idxs], ds['gt3Dcrop'][idxs])
            assert rec.tobytes() == np.array([p[0] for p in per]).tobytes(), (name, rep)
            assert np.array_equal(lab, np.stack([p[1] for p in per])), (name, rep)


def test_engine_lowering_is_pure_host_logic():
    """dpp_b200.engine.Engine._lower turns the recorded layer graph into the op list the C ABI executes; it needs no
    device.  Checks the op inventory of every net on the path and the convolution descriptors against SURVEY 8a's
    shape table (64 convolutions incl. the stem, 20 of them 3x3, 6 with stride 2 - stage 4 keeps 256 channels, so its
    blocks are identity blocks and ignore the stride, resnet.py:349-414; 60 BatchNorms in front of convolutions)."""
    from dpp_b200.engine import Engine
    from net.scalenet import ScaleNet, ScaleNetParams

    def lower(net):
        eng = Engine.__new__(Engine)
        eng.net, eng.output_sym, eng.B, eng.precision = net, net.output, int(net.cfgParams.batch_size), 1
        eng._lower()
        return eng

    def kinds(eng):
        out = {}
        for op in eng.ops:
            out[op['kind']] = out.get(op['kind'], 0) + 1
        return out
    rng = np.random.RandomState(23455)
    eng = lower(ResNet(rng, cfgParams=ResNetParams(type=0, batchSize=4, numJoints=1, nDims=30)))
    assert kinds(eng) == {'convpool': 1, 'conv': 63, 'bn_apply': 1, 'fc': 3} and eng.t_out.shape == (4, 30)
    descs = [eng._conv_desc(op) for op in eng.ops if op['kind'] == 'conv']
    assert sum(1 for d in descs if d.k == 3) == 20 and sum(1 for d in descs if d.stride == 2) == 6
    assert all(d.pad == d.k // 2 and d.N == 4 for d in descs)
    macs = sum(d.Ho * d.Wo * d.Cout * d.k * d.k * d.Cin for d in descs) + 128 * 128 * 32 * 25
    assert abs(macs / 1e6 - 107.0) < 0.1                          # 107.0 M conv MACs per frame
    assert len(set(id(op['in_bn']) for op in eng.ops if op['kind'] == 'conv' and op['in_bn'] is not None)) == 60
    eng = lower(ResNet(rng, cfgParams=ResNetParams(type=1, batchSize=2, numJoints=14, nDims=3)))
    assert kinds(eng)['fc'] == 4 and eng.t_out.shape == (2, 42)
    eng = lower(PoseRegNet(rng, cfgParams=PoseRegNetParams(type=0, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1,
                                                           nDims=30)))
    assert kinds(eng) == {'convpool': 3, 'fc': 3}
    eng = lower(ScaleNet(rng, cfgParams=ScaleNetParams(type=1, nChan=1, wIn=128, hIn=128, batchSize=2, numJoints=1, nDims=3)))
    assert kinds(eng) == {'convpool': 9, 'concat': 1, 'fc': 3} and len(eng.t_ins) == 3
    assert [t.shape for t in eng.t_ins] == [(2, 1, 128, 128), (2, 1, 64, 64), (2, 1, 32, 32)]


def test_parameter_layout_conversions_between_reference_and_device():
    """engine.Slot: reference layouts (what get_value / set_value / the pickles carry) <-> device layouts documented in
    include/dpp_b200.h.  Conv weights: W_kc[(r*kw + s)*Cin + c][o] = W[o][c][kh-1-r][kw-1-s] (true convolution,
    pre-flipped); FC weights behind a flattened NCHW tensor: rows (c,h,w) -> (h,w,c) because activations are NHWC on the
    device; ScaleNet's concatenated towers: the same permutation per tower segment.  Round trips are exact."""
    from dpp_b200.engine import Slot
    from net.sym import shared
    rng = np.random.RandomState(3)
    w = rng.randn(6, 4, 5, 3).astype(np.float32)                       # (O, I, kh, kw)
    s = Slot(shared(w, kind='convW'), 'w', 0)
    d = s.to_device_layout(w).reshape(5 * 3 * 4, 6)
    for (r, t, c, o) in [(0, 0, 0, 0), (4, 2, 3, 5), (1, 2, 0, 3), (3, 0, 2, 1)]:
        assert d[(r * 3 + t) * 4 + c, o] == w[o, c, 5 - 1 - r, 3 - 1 - t]
    assert np.array_equal(s.from_device_layout(s.to_device_layout(w)), w)
    C, H, Wd, n_out = 3, 2, 4, 5
    f = rng.randn(C * H * Wd, n_out).astype(np.float32)                # rows in the reference's (c,h,w) flatten order
    s = Slot(shared(f, kind='fcW'), 'w', 0)
    s.chw = (C, H, Wd)
    d = s.to_device_layout(f).reshape(H * Wd * C, n_out)
    for (c, h, x) in [(0, 0, 0), (2, 1, 3), (1, 0, 2)]:
        assert np.array_equal(d[(h * Wd + x) * C + c], f[(c * H + h) * Wd + x])
    assert np.array_equal(s.from_device_layout(s.to_device_layout(f)), f)
    # two towers of (2,1,2) and (1,2,2) features concatenated (ScaleNet, scalenet.py:174-178)
    g = rng.randn(4 + 4, 3).astype(np.float32)
    s = Slot(shared(g, kind='fcW'), 'w', 0)
    s.segments = [(0, (2, 1, 2)), (4, (1, 2, 2))]
    d = s.to_device_layout(g).reshape(8, 3)
    assert np.array_equal(d[(0 * 2 + 1) * 2 + 1], g[(1 * 1 + 0) * 2 + 1])          # tower 0: (c=1,h=0,w=1)
    assert np.array_equal(d[4 + (1 * 2 + 0) * 1 + 0], g[4 + (0 * 2 + 1) * 2 + 0])  # tower 1: (c=0,h=1,w=0)
    assert np.array_equal(s.from_device_layout(s.to_device_layout(g)), g)
    b = rng.randn(7).astype(np.float32)
    s = Slot(shared(b), 'w', 0)
    assert np.array_equal(s.from_device_layout(s.to_device_layout(b)), b)


def test_gradient_arena_zero_ranges_skip_the_assigned_hiddenlayer_weights():
    """Engine._g_zero_ranges: the HiddenLayer weight gradients are ASSIGNED by dpp_fc_bwd_ex (DPP_FC_DW_ASSIGN), so the
    per-step clear of the gradient arena covers exactly everything else - conv weights, biases, BatchNorm parameters,
    HiddenLayer biases, slot padding - in sorted, disjoint ranges."""
    import torch
    from dpp_b200.engine import Engine
    net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=4, numJoints=1, nDims=30))
    eng = Engine.__new__(Engine)
    eng.net, eng.output_sym, eng.B, eng.precision = net, net.output, 4, 1
    eng.torch, eng.dev = torch, torch.device('cpu')
    eng._lower()
    eng._alloc_params()
    ranges = eng._g_zero_ranges()
    assert ranges == sorted(ranges) and all(lo < hi for lo, hi in ranges)
    assert all(ranges[i][1] <= ranges[i + 1][0] for i in range(len(ranges) - 1))
    cleared = np.zeros(eng.n_w, bool)
    for lo, hi in ranges:
        cleared[lo:hi] = True
    fcw = [eng.slots[id(o['layer'].W)] for o in eng.ops if o['kind'] == 'fc']
    assert len(fcw) == 3 and max(s.size for s in fcw) == 16384 * 1024
    for s in fcw:
        assert not cleared[s.offset:s.offset + s.size].any()          # assigned, never cleared
        cleared[s.offset:s.offset + s.size] = True
    assert cleared.all()                                               # ... and nothing else is left out
    skipped = sum(s.size for s in fcw) / float(eng.n_w)
    assert 0.9 < skipped < 0.99                                        # the HiddenLayers hold ~95 % of the parameters
