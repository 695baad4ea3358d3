"""Deterministic synthetic hand-crop generator.  The datasets the reference trains
on (NYU / ICVL / MSRA15) are not available; this produces inputs with the value ranges and the
per-sample side arrays the entry scripts build (main_nyu_posereg_embedding.py:60-65):
normalised crops in [-1,1] as Dataset.imgStackDepthOnly makes them (data/dataset.py:97-103),
``cube``, ``com3D``, ``M`` (float32 crop transform) and ``gt3Dcrop`` (mm)."""
import numpy as np

f32 = np.float32
f64 = np.float64

DATASETS = {
    # name: (importer class name, cube, joints)
    'NYU': ('NYUImporter', (300, 300, 300), 14),
    'ICVL': ('ICVLImporter', (250, 250, 250), 16),
    'MSRA15': ('MSRA15Importer', (200, 200, 200), 21),
}


def make_importer(name):
    from data import importers
    return getattr(importers, DATASETS[name][0])(None)


def generate(name, n, seed=23455, cube=None, size=128):
    """Returns dict(x (n,1,S,S) f32 in [-1,1], gt3D (n,J,3) f32 = gt3Dcrop/(cube_z/2), cube (n,3),
    com3D (n,3), M (n,3,3), gt3Dcrop (n,J,3)) all float32."""
    from util.handdetector import HandDetector
    rng = np.random.RandomState(seed)
    di = make_importer(name)
    cube = tuple(cube if cube is not None else DATASETS[name][1])
    J = DATASETS[name][2]
    hd = HandDetector(np.zeros((size, size), f32) + 1., abs(di.fx), abs(di.fy), importer=di)
    x = np.zeros((n, 1, size, size), f32)
    cubes = np.asarray([cube] * n, f32)
    com3D = np.zeros((n, 3), f32)
    Ms = np.zeros((n, 3, 3), f32)
    gt3Dcrop = np.zeros((n, J, 3), f32)
    yy, xx = np.mgrid[0:size, 0:size]
    half = cube[2] / 2.
    for i in range(n):
        c3 = np.array([rng.uniform(-150, 150), rng.uniform(-150, 150), rng.uniform(400, 900)], f32)
        com = di.joint3DToImg(c3)
        com3D[i] = c3
        Ms[i] = hd.comToTransform(com, cubes[i], (size, size)).astype(f32)
        d = np.zeros((size, size), f32)
        for _ in range(6):
            cx, cy = rng.uniform(16, size - 16, 2)
            ax, ay = rng.uniform(4, 22, 2)
            depth = np.rint(f64(com[2]) + rng.uniform(-0.4, 0.4) * half)
            m = ((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 <= 1.0
            m &= (xx >= 16) & (xx < size - 16) & (yy >= 16) & (yy < size - 16)
            d[m] = depth
        # dataset.py:99-103 (normZeroOne False)
        imgD = d.copy()
        imgD[imgD == 0] = f32(f64(com[2]) + half)
        imgD -= com[2]
        imgD /= f32(half)
        x[i, 0] = imgD
        gt3Dcrop[i] = np.clip(rng.randn(J, 3) * 35., -half, half).astype(f32)
    gt3D = gt3Dcrop / f32(half)
    return dict(x=x, gt3D=gt3D.astype(f32), cube=cubes, com3D=com3D, M=Ms, gt3Dcrop=gt3Dcrop, importer=di, hd=hd)


def random_orthonormal_pca(n_components, dim, seed=0):
    """Stand-in for sklearn PCA fitted on sampled poses: orthonormal rows + a mean vector."""
    rng = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rng.randn(dim, dim))
    return q[:n_components].astype(f64), (rng.randn(dim) * 0.05).astype(f64)
