"""Training entry script - the reference's src/main_nyu_posereg_embedding.py after the mechanical
Python-3 pass (print(), range, pickle), with the SAME imports, class names, constructor arguments
and attribute pokes (:96-128, :144-167).  Only the data section differs: the NYU files are absent
(data/PUT_DATA_HERE), so ``data.synthetic`` supplies crops with the same arrays the importer /
dataset classes produce (:60-65).  Plots and the Tompson-baseline comparison (matplotlib, external
.mat file) are dropped.

  DPP_TRAIN=4096 DPP_VAL=512 DPP_EPOCHS=2 DPP_NET=resnet python main_nyu_posereg_embedding.py
"""
import os
import pickle
import numpy
from sklearn.decomposition import PCA

from data import synthetic
from net.hiddenlayer import HiddenLayer, HiddenLayerParams
from net.poseregnet import PoseRegNetParams, PoseRegNet
from net.resnet import ResNetParams, ResNet
from trainer.poseregnettrainer import PoseRegNetTrainerParams, PoseRegNetTrainer
from util.handdetector import HandDetector

if __name__ == '__main__':
    eval_prefix = 'NYU_EMB_t0nF8mp421fD553h1024_PCA30_AUGMENT'
    if not os.path.exists('./eval/' + eval_prefix + '/'):
        os.makedirs('./eval/' + eval_prefix + '/')

    rng = numpy.random.RandomState(23455)

    print("create data")
    aug_modes = ['com', 'rot', 'none']  # 'sc',

    n_train = int(os.environ.get('DPP_TRAIN', '4096'))
    n_val = int(os.environ.get('DPP_VAL', '512'))
    ds = synthetic.generate('NYU', n_train, seed=23455)
    dv = synthetic.generate('NYU', n_val, seed=777)
    di = ds['importer']
    train_data, train_gt3D = ds['x'], ds['gt3D']
    train_data_cube, train_data_com, train_data_M, train_gt3Dcrop = ds['cube'], ds['com3D'], ds['M'], ds['gt3Dcrop']
    val_data, val_gt3D = dv['x'], dv['gt3D']
    test_data1, test_gt3D1 = dv['x'], dv['gt3D']

    imgSizeW = train_data.shape[3]
    imgSizeH = train_data.shape[2]
    nChannels = train_data.shape[1]

    ####################################
    # convert data to embedding (main_nyu_posereg_embedding.py:85-92: PCA on 1e6 sampled poses)
    pca = PCA(n_components=30)
    pca.fit(HandDetector.sampleRandomPoses(di, rng, train_gt3Dcrop, train_data_com, train_data_cube,
                                           float(os.environ.get('DPP_POSES', '1e6')),
                                           aug_modes).reshape((-1, train_gt3D.shape[1] * 3)))
    train_gt3D_embed = pca.transform(train_gt3D.reshape((train_gt3D.shape[0], train_gt3D.shape[1] * 3)))
    val_gt3D_embed = pca.transform(val_gt3D.reshape((val_gt3D.shape[0], val_gt3D.shape[1] * 3)))

    ############################################################################
    print("create network")
    batchSize = 128
    if os.environ.get('DPP_NET', 'poseregnet') == 'resnet':
        poseNetParams = ResNetParams(type=0, nChan=nChannels, wIn=imgSizeW, hIn=imgSizeH, batchSize=batchSize,
                                     numJoints=1, nDims=train_gt3D_embed.shape[1])
        poseNet = ResNet(rng, cfgParams=poseNetParams)
    else:
        poseNetParams = PoseRegNetParams(type=0, nChan=nChannels, wIn=imgSizeW, hIn=imgSizeH, batchSize=batchSize,
                                         numJoints=1, nDims=train_gt3D_embed.shape[1])
        poseNet = PoseRegNet(rng, cfgParams=poseNetParams)

    poseNetTrainerParams = PoseRegNetTrainerParams()
    poseNetTrainerParams.batch_size = batchSize
    poseNetTrainerParams.learning_rate = 0.001
    poseNetTrainerParams.weightreg_factor = 0.0
    poseNetTrainerParams.force_macrobatch_reload = True
    poseNetTrainerParams.para_augment = True
    poseNetTrainerParams.augment_fun_params = {'fun': 'augment_poses', 'args': {'normZeroOne': False,
                                                                                'di': di,
                                                                                'aug_modes': aug_modes,
                                                                                'hd': HandDetector(train_data[0, 0].copy(), abs(di.fx), abs(di.fy), importer=di),
                                                                                'proj': pca}}

    print("setup trainer")
    poseNetTrainer = PoseRegNetTrainer(poseNet, poseNetTrainerParams, rng, './eval/' + eval_prefix)
    poseNetTrainer.setData(train_data, train_gt3D_embed, val_data, val_gt3D_embed)
    poseNetTrainer.addStaticData({'val_data_y3D': val_gt3D})
    poseNetTrainer.addStaticData({'pca_data': pca.components_, 'mean_data': pca.mean_})
    poseNetTrainer.addManagedData({'train_data_cube': train_data_cube,
                                   'train_data_com': train_data_com,
                                   'train_data_M': train_data_M,
                                   'train_gt3Dcrop': train_gt3Dcrop})
    poseNetTrainer.compileFunctions(compileDebugFcts=False)

    ###################################################################
    # TRAIN
    poseNetTrainer.verbose = bool(int(os.environ.get('DPP_VERBOSE', '0')))
    poseNetTrainerParams.validation_frequency = int(os.environ.get('DPP_VALFREQ', '1000'))
    train_res = poseNetTrainer.train(n_epochs=int(os.environ.get('DPP_EPOCHS', '100')))
    train_costs = train_res[0]
    val_errs = train_res[2]

    # save results
    poseNet.save("./eval/{}/net_{}.pkl".format(eval_prefix, eval_prefix))

    # add prior to network
    cfg = HiddenLayerParams(inputDim=(batchSize, train_gt3D_embed.shape[1]),
                            outputDim=(batchSize, numpy.prod(train_gt3D.shape[1:])), activation=None)
    pcalayer = HiddenLayer(rng, poseNet.layers[-1].output, cfg, layerNum=len(poseNet.layers))
    pcalayer.W.set_value(pca.components_)
    pcalayer.b.set_value(pca.mean_)
    poseNet.layers.append(pcalayer)
    poseNet.output = pcalayer.output
    poseNet.cfgParams.numJoints = train_gt3D.shape[1]
    poseNet.cfgParams.nDims = train_gt3D.shape[2]
    poseNet.cfgParams.outputDim = pcalayer.cfgParams.outputDim
    poseNet.save("./eval/{}/network_prior.pkl".format(eval_prefix))

    ###################################################################
    print("Testing ...")
    jts_embed = poseNet.computeOutput(test_data1)
    joints = []
    gt3D = []
    for i in range(test_data1.shape[0]):
        joints.append(jts_embed[i].reshape((-1, 3)) * (dv['cube'][i][2] / 2.) + dv['com3D'][i])
        gt3D.append(dv['gt3Dcrop'][i] + dv['com3D'][i])
    joints = numpy.array(joints)
    gt3D = numpy.array(gt3D)
    # handpose_evaluation.py:97: mean over frames and joints of the euclidean joint error
    err = numpy.sqrt(numpy.square(gt3D - joints).sum(axis=2))
    print("Train samples: {}, test samples: {}".format(train_data.shape[0], len(gt3D)))
    print("Mean error: {}mm, max error: {}mm".format(numpy.nanmean(err), numpy.nanmax(err)))
    print("first/last train cost: {} / {}".format(train_costs[0], train_costs[-1]))
    pickle.dump(joints, open("./eval/{}/result_{}.pkl".format(eval_prefix, eval_prefix), "wb"), protocol=2)
