"""Writes tests/golden/bench_first_step.json: the float64 oracle's cost of the batch bench.py's cost check runs
through the TIMED code path (CUDA graph, batch 128, 3xTF32 kernels) before it starts measuring.

Recipe (shared with bench.py through bench.first_batch): synthetic NYU set seed 23455 (2048 crops), 30-D orthonormal
PCA stand-in seed 1, batch = 128 indices + augmentation draws from RandomState(1234), augmentation by the oracle
(oracle/augment.py, NumPy index model of cv2), ResNet type 0 from RandomState(23455) in train mode, cost =
mean_b sum_d (out - y)^2 (trainer/poseregnettrainer.py:92-99).  Run in the build container:  python tests/golden/make_bench_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'deep-prior-pp_b200')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from oracle import augment as OA, nets as ON  # noqa: E402


def main():
    ds, comp, mean = bench.make_workload(seed=23455)
    idxs, recs, y, draws = bench.first_batch(ds, comp, mean, bench.B)
    cam = OA.Camera(**OA.NYU_CAM)
    x, yo = OA.augment_poses(ds['x'], ds['com3D'], ds['cube'], ds['M'], ds['gt3Dcrop'], list(idxs), draws,
                             bench.AUG_MODES, cam, OA.Hand(cam), pca_mean=mean, pca_components=comp)
    assert np.allclose(yo, y, atol=1e-5), "host label path and oracle label path disagree"
    onet = ON.build_resnet(np.random.RandomState(23455), type=0, batchSize=bench.B, numJoints=1, nDims=bench.E)
    with torch.no_grad():
        out, _ = onet.forward(torch.from_numpy(x), deterministic=False)
        cost = float(ON.cost_fn(onet, out, torch.from_numpy(y), bench.B, 1, bench.E, 0.0))
    json.dump({"cost": cost, "batch": bench.B, "n_resident": bench.N_RESIDENT, "recipe": "see tests/golden/make_bench_golden.py",
               "x_sha1": __import__('hashlib').sha1(np.ascontiguousarray(x).tobytes()).hexdigest()},
              open(os.path.join(HERE, 'bench_first_step.json'), 'w'), indent=1)
    print("oracle cost of bench.py's check batch:", cost)


if __name__ == '__main__':
    main()
