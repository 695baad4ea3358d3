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


def _samples(name, n, seed, cube, size):
    """the synthetic recipe: yields per sample (com3D f32, com f32 image coords, M f64, crop in mm, gt3Dcrop f32)"""
    from util.handdetector import HandDetector
    rng = np.random.RandomState(seed)
    di = make_importer(name)
    hd = HandDetector(np.zeros((size, size), f32) + 1., abs(di.fx), abs(di.fy), importer=di)
    J = DATASETS[name][2]
    yy, xx = np.mgrid[0:size, 0:size]
    half = cube[2] / 2.
    cube32 = np.asarray(cube, f32)
    for i in range(n):
        c3 = np.array([rng.uniform(-150, 150), rng.uniform(-150, 150), rng.uniform(400, 900)], f32)
        com = di.joint3DToImg(c3)
        M = hd.comToTransform(com, cube32, (size, size))
        d = np.zeros((size, size), f32)
        for _ in range(6):
            cx, cy = rng.uniform(16, size - 16, 2)
            ax, ay = rng.uniform(4, 22, 2)
            depth = np.rint(f64(com[2]) + rng.uniform(-0.4, 0.4) * half)
            m = ((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 <= 1.0
            m &= (xx >= 16) & (xx < size - 16) & (yy >= 16) & (yy < size - 16)
            d[m] = depth
        gt = np.clip(rng.randn(J, 3) * 35., -half, half).astype(f32)
        yield c3, com, M, d, gt
    return


def generate(name, n, seed=23455, cube=None, size=128):
    """Returns dict(x (n,1,S,S) f32 in [-1,1], gt3D (n,J,3) f32 = gt3Dcrop/(cube_z/2), cube (n,3),
    com3D (n,3), M (n,3,3), gt3Dcrop (n,J,3)) all float32."""
    from util.handdetector import HandDetector
    di = make_importer(name)
    cube = tuple(cube if cube is not None else DATASETS[name][1])
    J = DATASETS[name][2]
    hd = HandDetector(np.zeros((size, size), f32) + 1., abs(di.fx), abs(di.fy), importer=di)
    x = np.zeros((n, 1, size, size), f32)
    cubes = np.asarray([cube] * n, f32)
    com3D = np.zeros((n, 3), f32)
    Ms = np.zeros((n, 3, 3), f32)
    gt3Dcrop = np.zeros((n, J, 3), f32)
    half = cube[2] / 2.
    for i, (c3, com, M, d, gt) in enumerate(_samples(name, n, seed, cube, size)):
        com3D[i] = c3
        Ms[i] = M.astype(f32)
        # dataset.py:99-103 (normZeroOne False)
        imgD = d.copy()
        imgD[imgD == 0] = f32(f64(com[2]) + half)
        imgD -= com[2]
        imgD /= f32(half)
        x[i, 0] = imgD
        gt3Dcrop[i] = gt
    gt3D = gt3Dcrop / f32(half)
    return dict(x=x, gt3D=gt3D.astype(f32), cube=cubes, com3D=com3D, M=Ms, gt3Dcrop=gt3Dcrop, importer=di, hd=hd)


def generate_sequence(name, n, seed=23455, cube=None, size=128, seq_name='train'):
    """The same samples as ``generate`` packed the way the importers' ``loadSequence`` returns them (reference
    src/data/importers.py:1068-1075 for NYU): a ``NamedImgSequence`` of ``DepthFrame``s with the crop in mm
    (background 0), the crop transform ``T``, ``gt3Dcrop`` and the 3-D crop centre in ``com`` - the input of
    ``data.dataset.Dataset.imgStackDepthOnly`` and of the entry scripts' side arrays."""
    from data.basetypes import DepthFrame, NamedImgSequence
    from data.transformations import transformPoints2D
    cube = tuple(cube if cube is not None else DATASETS[name][1])
    di = make_importer(name)
    frames = []
    for c3, com, M, d, gt in _samples(name, n, seed, cube, size):
        gt3Dorig = (gt + c3).astype(f32)
        gtorig = di.joints3DToImg(gt3Dorig)                        # joints in the original image (u, v, d)
        gtcrop = transformPoints2D(gtorig, M)                      # ... and in the crop
        frames.append(DepthFrame(d, gtorig, gtcrop, M.astype(f32), gt3Dorig, gt, c3, '', '', 'left', {}))
    return NamedImgSequence(seq_name, frames, {'cube': cube})


def random_orthonormal_pca(n_components, dim, seed=0):
    """Stand-in for sklearn PCA fitted on sampled poses: orthonormal rows + a mean vector."""
    rng = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rng.randn(dim, dim))
    return q[:n_components].astype(f64), (rng.randn(dim) * 0.05).astype(f64)


def generate_frames(name, n, seed=23455, cube=None, edge_fraction=0.25, nd=0.):
    """Synthetic full depth frames for the inference cascade (BASELINE config 5; the reference's FileDevice feeds
    NYU test frames, src/test_realtimepipeline.py:66-73): ``importer.depth_map_size`` frames in integer mm with
    undefined depth ``nd``, a hand-like union of ellipses inside the cube around a random CoM, background clutter
    behind the cube, and - for ``edge_fraction`` of the frames - a CoM so close to the border that the crop window
    leaves the frame (getCrop's zero padding).  Returns dict(frames (n,H,W) f32, com3D (n,3) f32 true CoMs,
    lastcom (n,3) f64 the perturbed CoMs a detector / previous frame would hand over, cube, importer)."""
    rng = np.random.RandomState(seed)
    di = make_importer(name)
    cube = tuple(cube if cube is not None else DATASETS[name][1])
    Wf, Hf = di.depth_map_size
    frames = np.full((n, Hf, Wf), nd, f32)
    com3D = np.zeros((n, 3), f32)
    lastcom = np.zeros((n, 3), f64)
    yy, xx = np.mgrid[0:Hf, 0:Wf]
    for i in range(n):
        z = rng.uniform(450, 900)
        r_px = 0.5 * cube[0] / z * di.fx                     # half the window edge in pixels
        if rng.rand() < edge_fraction:
            u = rng.choice([rng.uniform(0.3 * r_px, 0.9 * r_px), Wf - rng.uniform(0.3 * r_px, 0.9 * r_px)])
            v = rng.uniform(r_px, Hf - r_px)
            if rng.rand() < 0.5:
                v = rng.choice([rng.uniform(0.3 * r_px, 0.9 * r_px), Hf - rng.uniform(0.3 * r_px, 0.9 * r_px)])
        else:
            u = rng.uniform(1.1 * r_px, Wf - 1.1 * r_px)
            v = rng.uniform(1.1 * r_px, Hf - 1.1 * r_px)
        com = np.array([u, v, z], f64)
        d = frames[i]
        # clutter behind the hand (cut by the z-threshold) and a near occluder in one corner of some frames
        m = (yy > v + 0.5 * r_px) & (np.abs(xx - u) < 0.4 * r_px)
        d[m] = np.rint(z + cube[2] * rng.uniform(0.6, 1.5))
        if rng.rand() < 0.3:
            m = (xx < u - 0.7 * r_px) & (yy < v - 0.7 * r_px)
            d[m] = np.rint(z - cube[2] * rng.uniform(0.55, 0.8))
        for _ in range(6):
            cx, cy = u + rng.uniform(-0.6, 0.6, 2) * r_px
            ax, ay = rng.uniform(0.06, 0.35, 2) * r_px
            depth = np.rint(z + rng.uniform(-0.4, 0.4) * cube[2] / 2.)
            m = ((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 <= 1.0
            d[m] = depth
        com3D[i] = di.jointImgTo3D(com)
        lastcom[i] = com + np.array([rng.uniform(-6, 6), rng.uniform(-6, 6), rng.uniform(-15, 15)])
    return dict(frames=frames, com3D=com3D, lastcom=lastcom, cube=cube, importer=di, nd=f32(nd))
