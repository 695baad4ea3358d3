"""CPU oracle: restatement of the reference's depth-crop augmentation (NumPy + cv2).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED against the reference's own code
executed in the build container (oracle/ref_harness.py -> tests/golden/reference_pins.npz,
tests/test_reference_pins.py: warp indices bit-exact, floats within a few float32 ulps), against
cv2 4.13.0 itself (``*_cv2`` variants) and the committed golden vectors.

Reference code restated (all under /root/reference/src):
  trainer/nettrainer.py:919-997      NetTrainer.augmentCrop
  trainer/poseregnettrainer.py:221-264  PoseRegNetTrainer.augment_poses
  util/handdetector.py:204-226 comToBounds, :228-258 comToTransform, :678-710 moveCoM,
      :712-747 rotateHand, :750-780 scaleHand, :782-803 recropHand
  data/transformations.py:71-88 rotatePoint2D
  data/importers.py:80-119 (ICVL/base), :756-793 (MSRA15), :1187-1224 (NYU) projections
  util/handdetector.py:805-909 sampleRandomPoses (rot3D=False), data/transformations.py:91-103 rotatePoints2D

Dtype discipline: the reference ran on NumPy 1.x value-based casting.
Every scalar expression below is written with the dtype NumPy 1.x would have produced:
  * np.float32-scalar (op) python-float / np.float64  -> float64
  * np.float32-scalar (op) np.float32-scalar          -> float32
  * float32-array (op) float64-scalar                 -> float32 loop, scalar cast to f32
  * stores into np.float32 arrays round to float32.
py2 integer division is explicit (``//``).
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

f32 = np.float32
f64 = np.float64


# --------------------------------------------------------------------------------------
# camera projections
# --------------------------------------------------------------------------------------
class Camera(object):
    """Pin-hole projection. ``flip_y`` True for NYU (importers.py:1187-1224) and MSRA15
    (:756-793); False for ICVL / base class (:80-119)."""

    def __init__(self, fx, fy, ux, uy, flip_y):
        self.fx, self.fy, self.ux, self.uy, self.flip_y = float(fx), float(fy), float(ux), float(uy), bool(flip_y)

    def jointImgTo3D(self, s):
        ret = np.zeros((3,), f32)
        s0, s1, s2 = s[0], s[1], s[2]
        # (sample[0]-ux) -> f64 ; * sample[2] ; / fx
        ret[0] = (f64(s0) - self.ux) * f64(s2) / self.fx
        if self.flip_y:
            ret[1] = (self.uy - f64(s1)) * f64(s2) / self.fy
        else:
            ret[1] = (f64(s1) - self.uy) * f64(s2) / self.fy
        ret[2] = s2
        return ret

    def joint3DToImg(self, s):
        ret = np.zeros((3,), f32)
        if s[2] == 0.:
            ret[0] = self.ux
            ret[1] = self.uy
            return ret
        # sample[0]/sample[2]: stays in the sample's own dtype (f32/f32 -> f32), then f64
        if np.asarray(s).dtype == f32:
            q0 = f64(f32(s[0]) / f32(s[2]))
            q1 = f64(f32(s[1]) / f32(s[2]))
        else:
            q0 = f64(s[0]) / f64(s[2])
            q1 = f64(s[1]) / f64(s[2])
        ret[0] = q0 * self.fx + self.ux
        if self.flip_y:
            ret[1] = self.uy - q1 * self.fy
        else:
            ret[1] = q1 * self.fy + self.uy
        ret[2] = s[2]
        return ret

    def jointsImgTo3D(self, S):
        return np.stack([self.jointImgTo3D(S[i]) for i in range(S.shape[0])]).astype(f32)

    def joints3DToImg(self, S):
        return np.stack([self.joint3DToImg(S[i]) for i in range(S.shape[0])]).astype(f32)


NYU_CAM = dict(fx=588.03, fy=587.07, ux=320., uy=240., flip_y=True)        # importers.py:891
ICVL_CAM = dict(fx=241.42, fy=241.42, ux=160., uy=120., flip_y=False)      # importers.py:199
MSRA_CAM = dict(fx=241.42, fy=241.42, ux=160., uy=120., flip_y=True)       # importers.py:547


# --------------------------------------------------------------------------------------
# crop geometry, handdetector.py:204-258
# --------------------------------------------------------------------------------------
def com_to_bounds(com, size, fx, fy):
    """handdetector.py:204-226 (non-degenerate branch; caller guards com[2]~0).
    ``com`` is a float32 (u,v,d) array; ``size`` float32 array or list of float64."""
    c0, c1, c2 = f32(com[0]), f32(com[1]), f32(com[2])
    zstart = f64(c2) - f64(size[2]) / 2.
    zend = f64(c2) + f64(size[2]) / 2.
    # com[0]*com[2] is f32*f32 -> f32 ; then / fx -> f64
    p0 = f64(f32(c0 * c2))
    p1 = f64(f32(c1 * c2))
    xstart = int(np.floor((p0 / fx - f64(size[0]) / 2.) / f64(c2) * fx + 0.5))
    xend = int(np.floor((p0 / fx + f64(size[0]) / 2.) / f64(c2) * fx + 0.5))
    ystart = int(np.floor((p1 / fy - f64(size[1]) / 2.) / f64(c2) * fy + 0.5))
    yend = int(np.floor((p1 / fy + f64(size[1]) / 2.) / f64(c2) * fy + 0.5))
    return xstart, xend, ystart, yend, zstart, zend


def com_to_transform(com, size, fx, fy, dsize=(128, 128)):
    """handdetector.py:228-258, py2 integer division explicit, sz[1]/sz[0] swap kept."""
    xstart, xend, ystart, yend, _, _ = com_to_bounds(com, size, fx, fy)
    trans = np.eye(3)
    trans[0, 2] = -xstart
    trans[1, 2] = -ystart
    wb = (xend - xstart)
    hb = (yend - ystart)
    if wb > hb:
        scale = np.eye(3) * dsize[0] / float(wb)
        sz = (dsize[0], hb * dsize[0] // wb)
    else:
        scale = np.eye(3) * dsize[1] / float(hb)
        sz = (wb * dsize[1] // hb, dsize[1])
    scale[2, 2] = 1
    xs = int(np.floor(dsize[0] / 2. - sz[1] / 2.))
    ys = int(np.floor(dsize[1] / 2. - sz[0] / 2.))
    off = np.eye(3)
    off[0, 2] = xs
    off[1, 2] = ys
    return np.dot(off, np.dot(scale, trans))


# --------------------------------------------------------------------------------------
# nearest-neighbour warp index models of cv2 4.13.0, and the cv2 calls
# --------------------------------------------------------------------------------------
def affine_inverse(M):
    """cv2 warpAffine inverts the 2x3 forward matrix in fp64 like this (imgwarp.cpp)."""
    M = np.asarray(M, f64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1. / D if D != 0 else 0.
    A11 = M[1, 1] * D
    A22 = M[0, 0] * D
    i00, i01, i10, i11 = A11, -M[0, 1] * D, -M[1, 0] * D, A22
    b0 = -i00 * M[0, 2] - i01 * M[1, 2]
    b1 = -i10 * M[0, 2] - i11 * M[1, 2]
    return np.array([i00, i01, b0, i10, i11, b1], f64)


def warp_affine_nn_indices(inv6, W, H):
    """AB_BITS=10 fixed-point source indices; returns (Y, X, inside)."""
    i00, i01, b0, i10, i11, b1 = [f64(v) for v in inv6]
    x = np.arange(W, dtype=f64)
    y = np.arange(H, dtype=f64)
    adelta = np.rint(i00 * x * 1024.).astype(np.int64)
    bdelta = np.rint(i10 * x * 1024.).astype(np.int64)
    X0 = np.rint((i01 * y + b0) * 1024.).astype(np.int64) + 512
    Y0 = np.rint((i11 * y + b1) * 1024.).astype(np.int64) + 512
    X = (X0[:, None] + adelta[None, :]) >> 10
    Y = (Y0[:, None] + bdelta[None, :]) >> 10
    inside = (X >= 0) & (X < W) & (Y >= 0) & (Y < H)
    return Y, X, inside


def warp_affine_nn(img, M, border=0.):
    H, W = img.shape
    Y, X, inside = warp_affine_nn_indices(affine_inverse(M), W, H)
    out = np.full((H, W), border, img.dtype)
    out[inside] = img[Y[inside], X[inside]]
    return out


def warp_affine_nn_cv2(img, M, border=0.):
    return cv2.warpAffine(img, M, (img.shape[1], img.shape[0]), flags=cv2.INTER_NEAREST,
                          borderMode=cv2.BORDER_CONSTANT, borderValue=border)


def rotation_matrix_2d(center, angle_deg, scale=1.0):
    """cv2.getRotationMatrix2D restated (fp64, libm cos/sin, angle*(pi/180) constant-folded
    as OpenCV does); bit-identical to cv2 4.13.0 on 200k random angles (tests)."""
    import math
    a = float(angle_deg) * (math.pi / 180.)
    alpha = math.cos(a) * scale
    beta = math.sin(a) * scale
    cx, cy = float(center[0]), float(center[1])
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy],
                     [-beta, alpha, beta * cx + (1 - alpha) * cy]], f64)


def invert3x3_cv(S):
    """cv::invert for a 3x3 CV_64F matrix (closed-form cofactors, fp64) - what
    cv2.warpPerspective applies to the forward matrix.  NOT numpy.linalg.inv (LAPACK LU):
    the last-bit differences decide nearest-neighbour ties."""
    S = np.asarray(S, f64)
    d = S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0]) \
        + S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0])
    d = 1. / d
    t = np.zeros(9, f64)
    t[0] = (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d
    t[1] = (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d
    t[2] = (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d
    t[3] = (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d
    t[4] = (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d
    t[5] = (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d
    t[6] = (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d
    t[7] = (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d
    t[8] = (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d
    return t


def warp_perspective_nn_indices(Hinv, W, H, lanes=4):
    """cv2 4.13.0 INTER_NEAREST warpPerspective coordinate rule, recovered by black-box
    probing.  It reproduces cv2 exactly on 2988 of 3000 realistic crop matrices (whose
    ratios of small integers put whole rows/columns on exact .5 ties); on the remaining
    0.4 % one tie row/column resolves differently (a last-ulp effect inside cv2's SIMD
    code that the probing could not attribute).  tests/test_oracle_warp.py bounds this.
      * per-row constants are RUNNING fp64 sums down the rows: r(0)=m2, r(y+1)=r(y)+m1
      * along a row, `lanes`(=4) fp64 accumulators: a_j = j*m0 + r(y), then += 4*m0 per step
        (same for the y numerator and for w)
      * sx = nx / w ; border unless 0 <= sx <= W-1 and 0 <= sy <= H-1 (continuous test)
      * X = floor(sx + 0.5), Y = floor(sy + 0.5)
    The rounding of every partial sum matters: the crop matrices are ratios of small
    integers, so exact .5 ties cover whole rows/columns."""
    m = np.asarray(Hinv, f64).reshape(9)
    assert W % lanes == 0

    def line(a, b, c):
        r = np.zeros((H, 1), f64)
        v = f64(c)
        for i in range(H):
            r[i, 0] = v
            v = v + b
        acc = np.arange(lanes, dtype=f64)[None, :] * a + r
        out = np.zeros((H, W), f64)
        for k in range(W // lanes):
            out[:, k * lanes:(k + 1) * lanes] = acc
            acc = acc + lanes * a
        return out

    nx = line(m[0], m[1], m[2])
    ny = line(m[3], m[4], m[5])
    w = line(m[6], m[7], m[8])
    with np.errstate(divide='ignore', invalid='ignore'):
        fx = nx / w
        fy = ny / w
    inside = (fx >= 0) & (fx <= W - 1) & (fy >= 0) & (fy <= H - 1)
    X = np.where(inside, np.floor(fx + 0.5), 0).astype(np.int64)
    Y = np.where(inside, np.floor(fy + 0.5), 0).astype(np.int64)
    return Y, X, inside


def warp_perspective_nn(img, Hm, border=0.):
    Hh, W = img.shape
    Y, X, inside = warp_perspective_nn_indices(invert3x3_cv(Hm), W, Hh)
    out = np.full((Hh, W), border, img.dtype)
    out[inside] = img[Y[inside], X[inside]]
    return out


def warp_perspective_nn_cv2(img, Hm, border=0.):
    return cv2.warpPerspective(img, np.asarray(Hm, f64), (img.shape[1], img.shape[0]),
                               flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT,
                               borderValue=float(border))


# --------------------------------------------------------------------------------------
# hand manipulation, handdetector.py:678-803
# --------------------------------------------------------------------------------------
class Hand(object):
    """The slice of HandDetector the augmentation touches. ``use_cv2`` selects the real
    cv2 warps (ground truth) or the NumPy index models (what the CUDA kernel mirrors)."""

    def __init__(self, cam, use_cv2=False):
        self.cam = cam
        self.fx, self.fy = abs(cam.fx), abs(cam.fy)       # main_nyu...py:110
        self.use_cv2 = use_cv2

    def recropHand(self, crop, M, Mnew, target_size, background_value=0., nv_val=0.,
                   thresh_z=True, com=None, size=(250, 250, 250)):
        """handdetector.py:782-803."""
        Hm = np.dot(M, Mnew)
        if self.use_cv2:
            warped = warp_perspective_nn_cv2(crop, Hm, background_value)
        else:
            warped = warp_perspective_nn(crop, Hm, background_value)
        # numpy.isclose(warped, nv_val): |a-b| <= atol + rtol*|b| on the f32 array
        close = np.abs(warped - f32(nv_val)) <= f32(1e-8 + 1e-5 * abs(nv_val))
        warped[close] = background_value
        if thresh_z:
            _, _, _, _, zstart, zend = com_to_bounds(com, size, self.fx, self.fy)
            zs, ze = f32(zstart), f32(zend)          # f32 array vs f64 scalar: f32 loop
            msk1 = np.logical_and(warped < zs, warped != 0)
            msk2 = np.logical_and(warped > ze, warped != 0)
            warped[msk1] = zs
            warped[msk2] = 0.
        return warped

    def moveCoM(self, dpt, cube, com, off, joints3D, M):
        """handdetector.py:678-710."""
        if np.allclose(off, 0.):
            return dpt, joints3D, com, M
        new_com = self.cam.joint3DToImg(self.cam.jointImgTo3D(com).astype(f64) + np.asarray(off, f64))
        if not (np.allclose(com[2], 0.) or np.allclose(new_com[2], 0.)):
            Mnew = com_to_transform(new_com, cube, self.fx, self.fy, dpt.shape)
            new_dpt = self.recropHand(dpt, Mnew, np.linalg.inv(M), dpt.shape, 0., 32000., True,
                                      new_com, cube)
        else:
            Mnew = M
            new_dpt = dpt
        new_joints3D = (joints3D + self.cam.jointImgTo3D(com)) - self.cam.jointImgTo3D(new_com)
        return new_dpt, new_joints3D.astype(f32), new_com, Mnew

    def rotateHand(self, dpt, cube, com, rot, joints3D):
        """handdetector.py:712-747 + transformations.py:71-88."""
        if np.allclose(rot, 0.):
            return dpt, joints3D, rot
        rot = np.mod(rot, 360)
        if self.use_cv2:
            M = cv2.getRotationMatrix2D((dpt.shape[1] // 2, dpt.shape[0] // 2), -rot, 1)
            new_dpt = warp_affine_nn_cv2(dpt, M, 0.)
        else:
            M = rotation_matrix_2d((dpt.shape[1] // 2, dpt.shape[0] // 2), -rot, 1)
            new_dpt = warp_affine_nn(dpt, M, 0.)
        com3D = self.cam.jointImgTo3D(com)
        joint_2D = self.cam.joints3DToImg((joints3D + com3D).astype(f32))
        data_2D = np.zeros_like(joint_2D)
        alpha = f64(rot) * np.pi / 180.
        ca, sa = np.cos(alpha), np.sin(alpha)
        for k in range(data_2D.shape[0]):
            pp = joint_2D[k].copy()
            pp[0:2] -= com[0:2]                         # f32
            pr = np.zeros_like(pp)
            pr[0] = f64(pp[0]) * ca - f64(pp[1]) * sa   # f32*f64 -> f64, store f32
            pr[1] = f64(pp[0]) * sa + f64(pp[1]) * ca
            pr[2] = pp[2]
            pr[0:2] += com[0:2]
            data_2D[k] = pr
        new_joints3D = (self.cam.jointsImgTo3D(data_2D) - com3D).astype(f32)
        return new_dpt, new_joints3D, rot

    def scaleHand(self, dpt, cube, com, sc, joints3D, M):
        """handdetector.py:750-780."""
        if np.allclose(sc, 1.):
            return dpt, joints3D, cube, M
        new_cube = [f64(s) * f64(sc) for s in cube]
        if not np.allclose(com[2], 0.):
            Mnew = com_to_transform(com, new_cube, self.fx, self.fy, dpt.shape)
            new_dpt = self.recropHand(dpt, Mnew, np.linalg.inv(M), dpt.shape, 0., 32000., True,
                                      com, cube)
        else:
            Mnew = M
            new_dpt = dpt
        return new_dpt, joints3D, new_cube, Mnew


# --------------------------------------------------------------------------------------
# augmentCrop, nettrainer.py:919-997  (normZeroOne=False branch; the mains use it)
# --------------------------------------------------------------------------------------
def draw_aug_params(rng, n_modes, sigma_com=5., sigma_sc=0.02, rot_range=180.):
    """nettrainer.py:954-957: all four always drawn, in this order."""
    mode = rng.randint(0, n_modes)
    off = rng.randn(3) * sigma_com
    rot = rng.uniform(-rot_range, rot_range)
    sc = abs(1. + rng.randn() * sigma_sc)
    return mode, off, rot, sc


def augment_crop(img, gt3Dcrop, com, cube, M, mode_name, off, rot, sc, hd):
    """nettrainer.py:919-997 with the random draws passed in explicitly.
    img: (128,128) f32 normalised crop; gt3Dcrop (J,3) f32 mm; com (3,) f32 (u,v,d);
    cube (3,) f32; M (3,3) f32.  Returns (imgD f32, curLabel (J,3) f32, cube, com, M)."""
    img = np.asarray(img, f32)
    half = f32(f64(cube[2]) / 2.)                 # f32-array * f64-scalar -> f32 loop
    img = img * half + f32(com[2])
    premax = img.max()
    if mode_name == 'com':
        imgD, new_joints3D, com, M = hd.moveCoM(img.astype(f32), cube, com, off, gt3Dcrop, M)
        curLabel = new_joints3D / f32(f64(cube[2]) / 2.)
    elif mode_name == 'rot':
        imgD, new_joints3D, rot = hd.rotateHand(img.astype(f32), cube, com, rot, gt3Dcrop)
        curLabel = new_joints3D / f32(f64(cube[2]) / 2.)
    elif mode_name == 'sc':
        imgD, new_joints3D, cube, M = hd.scaleHand(img.astype(f32), cube, com, sc, gt3Dcrop, M)
        curLabel = new_joints3D / f32(f64(cube[2]) / 2.)
    elif mode_name == 'none':
        imgD = img
        curLabel = gt3Dcrop / f32(f64(cube[2]) / 2.)
    else:
        raise NotImplementedError()
    imgD = np.array(imgD, f32, copy=True)
    hi = f32(f64(com[2]) + f64(cube[2]) / 2.)
    lo = f32(f64(com[2]) - f64(cube[2]) / 2.)
    imgD[imgD == premax] = hi
    imgD[imgD == 0] = hi
    imgD[imgD >= hi] = hi
    imgD[imgD <= lo] = lo
    imgD -= f32(com[2])
    imgD /= f32(f64(cube[2]) / 2.)
    return imgD, np.asarray(curLabel, f32), np.asarray(cube), com, M


def pca_transform(label, mean, components):
    """sklearn PCA.transform: (X - mean_) . components_^T in fp64 (poseregnettrainer.py:262)."""
    X = np.asarray(label, f64).reshape(1, -1) - np.asarray(mean, f64)
    return np.dot(X, np.asarray(components, f64).T)[0]


def augment_poses(xDB, comDB, cubeDB, MDB, gt3DcropDB, idxs, draws, aug_modes, cam, hd,
                  pca_mean=None, pca_components=None):
    """poseregnettrainer.py:221-264 for the single-macro-batch case; ``draws`` is a list of
    (mode, off, rot, sc) per sample.  Returns (x (n,1,H,W) f32, y (n,E or J*3) f32)."""
    xs, ys = [], []
    for i, (mode, off, rot, sc) in zip(idxs, draws):
        img = xDB[i, 0].copy()
        com = cam.joint3DToImg(comDB[i])
        cube = cubeDB[i].copy()
        M = MDB[i].copy()
        gt = gt3DcropDB[i].copy()
        imgD, lab, _, _, _ = augment_crop(img, gt, com, cube, M, aug_modes[mode], off, rot, sc, hd)
        xs.append(imgD[None])
        if pca_mean is not None:
            ys.append(pca_transform(lab, pca_mean, pca_components).astype(f32))
        else:
            ys.append(lab.reshape(-1).astype(f32))
    return np.stack(xs).astype(f32), np.stack(ys).astype(f32)


# --------------------------------------------------------------------------------------
# sampleRandomPoses, handdetector.py:805-909  (rot3D=False; the mains never set it,
# main_nyu_posereg_embedding.py:87-88, and rot3D needs the absent transforms3d package)
# --------------------------------------------------------------------------------------
POSE_MODES = ['none', 'rot', 'sc', 'com', 'rot+com', 'com+rot', 'rot+com+sc', 'rot+sc+com', 'sc+rot+com',
              'sc+com+rot', 'com+sc+rot', 'com+rot+sc']


def draw_pose_params(rng, n_modes, n_base, num_poses, sigma_com=5., sigma_sc=0.02, rot_range=180.):
    """handdetector.py:837-841: five array draws, in this order."""
    n = int(num_poses)
    modes = rng.randint(0, n_modes, n)
    ridxs = rng.randint(0, n_base, n)
    off = rng.randn(n, 3) * sigma_com
    sc = np.fabs(rng.randn(n) * sigma_sc + 1.)
    rot = rng.uniform(-rot_range, rot_range, size=(n, 3))
    return modes, ridxs, off, sc, rot


def _rotate_points_2d(cam, pts3d, center2d, angle):
    """importer.joints3DToImg -> transformations.py:91-103 rotatePoints2D -> importer.jointsImgTo3D."""
    joint_2D = cam.joints3DToImg(pts3d.astype(f32))
    alpha = f64(angle) * np.pi / 180.
    ca, sa = np.cos(alpha), np.sin(alpha)
    data_2D = np.zeros_like(joint_2D)
    for k in range(joint_2D.shape[0]):
        pp = joint_2D[k].copy()
        pp[0:2] -= center2d[0:2]                   # f32
        pr = np.zeros_like(pp)
        pr[0] = f64(pp[0]) * ca - f64(pp[1]) * sa  # f32 scalar * f64 scalar -> f64, store f32
        pr[1] = f64(pp[0]) * sa + f64(pp[1]) * ca
        pr[2] = pp[2]
        pr[0:2] += center2d[0:2]
        data_2D[k] = pr
    return cam.jointsImgTo3D(data_2D)


def sample_random_poses(cam, rng, base_poses, base_com, base_cube, num_poses, aug_modes, retall=False,
                        sigma_com=None, sigma_sc=None, rot_range=None):
    """handdetector.py:805-909 with rot3D=False.  base_* are float32 arrays (the mains pass float32)."""
    sigma_com = 5. if sigma_com is None else sigma_com
    sigma_sc = 0.02 if sigma_sc is None else sigma_sc
    rot_range = 180. if rot_range is None else rot_range
    assert all(m in POSE_MODES for m in aug_modes)
    n = int(num_poses)
    new_poses = np.zeros((n, base_poses.shape[1], base_poses.shape[2]), dtype=base_poses.dtype)
    new_com = np.zeros((n, 3), dtype=base_poses.dtype)
    new_cube = np.zeros((n, 3), dtype=base_poses.dtype)
    modes, ridxs, off, sc, rot = draw_pose_params(rng, len(aug_modes), base_poses.shape[0], n, sigma_com, sigma_sc,
                                                  rot_range)
    if aug_modes == ['none']:
        half = (base_cube[:, 2] / f32(2.))                         # f32 array / python float stays f32
        if retall is True:
            return base_poses / half[:, None, None], base_com, base_cube
        return base_poses / half[:, None, None]
    for i in range(n):
        name = aug_modes[modes[i]]
        cube = base_cube[ridxs[i]]
        com3D = base_com[ridxs[i]]
        pose = base_poses[ridxs[i]]

        def half():                                   # new_cube[i][2]/2.: f32 scalar / python float -> f64 -> the
            return f32(f64(new_cube[i][2]) / 2.)      # f32 array is divided by it in f32
        if name == 'com':
            new_com[i] = com3D.astype(f64) + off[i]
            new_cube[i] = cube
            new_poses[i] = (pose + com3D - new_com[i]) / half()
        elif name == 'rot':
            new_com[i] = com3D
            new_cube[i] = cube
            data3D = _rotate_points_2d(cam, pose + new_com[i], cam.joint3DToImg(com3D), rot[i, 0])
            new_poses[i] = (data3D - new_com[i]) / half()
        elif name == 'sc':
            new_com[i] = com3D
            new_cube[i] = cube * f32(sc[i])           # f32 array * f64 scalar: the scalar is cast to f32
            new_poses[i] = pose / half()
        elif name == 'none':
            new_com[i] = com3D
            new_cube[i] = cube
            new_poses[i] = pose / half()
        elif name in ('rot+com', 'com+rot'):
            new_com[i] = com3D.astype(f64) + off[i]
            new_cube[i] = cube
            p = (pose + com3D - new_com[i])
            data3D = _rotate_points_2d(cam, p + com3D, cam.joint3DToImg(new_com[i]), rot[i, 0])
            new_poses[i] = (data3D - com3D) / half()
        elif name in ('rot+com+sc', 'rot+sc+com'):
            # the other four orderings compare the LIST aug_modes with a string in the reference (:893) and
            # therefore fall through to NotImplementedError
            new_com[i] = com3D.astype(f64) + off[i]
            new_cube[i] = cube
            p = (pose + com3D - new_com[i])
            p = p * f32(sc[i])
            data3D = _rotate_points_2d(cam, p + com3D, cam.joint3DToImg(new_com[i]), rot[i, 0])
            new_poses[i] = (data3D - com3D) / half()
        else:
            raise NotImplementedError()
    if retall is True:
        return new_poses, new_com, new_cube, rot
    return new_poses
