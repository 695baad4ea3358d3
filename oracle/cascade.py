"""CPU oracle: restatement of the reference's inference cascade (CoM refinement -> re-crop -> pose regression).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED against the reference's own HandDetector.track /
refineCoM / cropArea3D executed in the build container (oracle/ref_harness.py -> tests/golden/reference_pins.npz,
tests/test_reference_pins.py: the refinement net's inputs and the pose net's crop bit-exact); the nearest-neighbour
resize rule is also pinned against cv2 4.13.0 itself (``resize_nn_cv2`` vs ``resize_nn``,
tests/test_oracle_cascade.py) and the whole cascade by the committed vector tests/golden/cascade_nyu.npz.

Reference code restated (all under /root/reference/src):
  util/handdetector.py:204-226   comToBounds        (dtype-aware: ``com`` is float64 out of ``detect``,
                                                     float32 out of ``joint3DToImg``)
  util/handdetector.py:260-296   getCrop            (slice + zero padding + z-threshold)
  util/handdetector.py:336-351   resizeCrop         (cv2.resize INTER_NEAREST, resizeMethod = RESIZE_CV2_NN :69)
  util/handdetector.py:382-490   cropArea3D         (docom=False branch as the pipeline calls it)
  util/handdetector.py:511-533   track              (doHandSize=False: the CoM-refinement step)
  util/handdetector.py:634-676   refineCoM          (normalise, centre crops 64 / 32, ScaleNet, * cube_z/2)
  util/realtimehandposepipeline.py:296-333  detect  (crop + normalisation; ``crop.clip(..)`` there discards its
                                                     result, so the final crop is NOT clamped - kept as is)
  util/realtimehandposepipeline.py:335-368  estimatePose (mirror for the right hand)
  util/realtimehandposepipeline.py:197-198  pose = estimatePose(..) * cube_z/2. + com3D

Dtype discipline: NumPy 1.x value-based casting made explicit, as in oracle/augment.py.
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

f32 = np.float32
f64 = np.float64


# --------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------
def com_to_bounds(com, size, fx, fy):
    """handdetector.py:204-226, non-degenerate branch.  ``com[0]*com[2]`` is computed in com's own dtype
    (float32 product for a float32 com, float64 for the float64 com of ``detect``); everything after the first
    python-float operand is float64."""
    com = np.asarray(com)
    if com.dtype == f32:
        p0, p1 = f64(f32(com[0] * com[2])), f64(f32(com[1] * com[2]))
    else:
        p0, p1 = f64(com[0]) * f64(com[2]), f64(com[1]) * f64(com[2])
    c2 = f64(com[2])
    zstart = c2 - size[2] / 2.
    zend = c2 + size[2] / 2.
    xstart = int(np.floor((p0 / fx - size[0] / 2.) / c2 * fx + 0.5))
    xend = int(np.floor((p0 / fx + size[0] / 2.) / c2 * fx + 0.5))
    ystart = int(np.floor((p1 / fy - size[1] / 2.) / c2 * fy + 0.5))
    yend = int(np.floor((p1 / fy + size[1] / 2.) / c2 * fy + 0.5))
    return xstart, xend, ystart, yend, zstart, zend


def get_crop(dpt, xstart, xend, ystart, yend, zstart, zend, thresh_z=True, background=0):
    """handdetector.py:260-296 (2-D branch), literally."""
    cropped = dpt[max(ystart, 0):min(yend, dpt.shape[0]), max(xstart, 0):min(xend, dpt.shape[1])].copy()
    cropped = np.pad(cropped, ((abs(ystart) - max(ystart, 0), abs(yend) - min(yend, dpt.shape[0])),
                               (abs(xstart) - max(xstart, 0), abs(xend) - min(xend, dpt.shape[1]))),
                     mode='constant', constant_values=background)
    if thresh_z is True:
        # float32 array vs float64 scalar: the comparison and the store happen in float32
        msk1 = np.logical_and(cropped < f32(zstart), cropped != 0)
        msk2 = np.logical_and(cropped > f32(zend), cropped != 0)
        cropped[msk1] = f32(zstart)
        cropped[msk2] = 0.
    return cropped


def resize_nn_indices(n_dst, n_src):
    """cv2 4.13.0 resize(INTER_NEAREST) source index of every destination index along one axis:
    ``min(floor(x * (1. / (n_dst / n_src))), n_src - 1)`` in fp64 (resize.cpp, resizeNN)."""
    inv = 1. / (f64(n_dst) / f64(n_src))
    return np.minimum(np.floor(np.arange(n_dst, dtype=f64) * inv).astype(np.int64), n_src - 1)


def resize_nn(crop, sz):
    """Index model of ``cv2.resize(crop, sz, interpolation=cv2.INTER_NEAREST)``; ``sz`` = (width, height)."""
    sx = resize_nn_indices(int(sz[0]), crop.shape[1])
    sy = resize_nn_indices(int(sz[1]), crop.shape[0])
    return crop[sy][:, sx]


def resize_nn_cv2(crop, sz):
    return cv2.resize(crop, (int(sz[0]), int(sz[1])), interpolation=cv2.INTER_NEAREST)


def nd_value(dpt):
    """handdetector.py:122-130 getNDValue (with the detector's minDepth/maxDepth, :57-58): the mode of the
    under-range (or over-range) pixels; scipy.stats.mode returns the smallest of equally frequent values."""
    max_depth = min(1500, dpt.max())
    min_depth = max(10, dpt.min())
    lo = dpt[dpt < min_depth]
    hi = dpt[dpt > max_depth]
    sel = lo if lo.shape[0] > hi.shape[0] else hi
    vals, counts = np.unique(sel, return_counts=True)
    return vals[np.argmax(counts)]


# --------------------------------------------------------------------------------------
# refineCoM / track
# --------------------------------------------------------------------------------------
def refine_inputs(cropped128, size, com):
    """handdetector.py:640-664: normalised crop and its 64x64 / 32x32 centre crops, NCHW with N=C=1."""
    imgD = np.asarray(cropped128.copy(), 'float32')
    hi = f32(f64(com[2]) + size[2] / 2.)
    lo = f32(f64(com[2]) - size[2] / 2.)
    imgD[imgD == 0] = hi
    imgD[imgD >= hi] = hi
    imgD[imgD <= lo] = lo
    imgD -= f32(com[2])
    imgD /= f32(size[2] / 2.)
    t = np.zeros((1, 1) + cropped128.shape, f32)
    t[0, 0] = imgD
    H, W = t.shape[2], t.shape[3]
    d = (H // 2, W // 2)
    xs, ys = int(H / 2 - d[0] / 2), int(W / 2 - d[1] / 2)
    t2 = t[:, :, ys:ys + d[1], xs:xs + d[0]]
    d = (H // 4, W // 4)
    xs, ys = int(H / 2 - d[0] / 2), int(W / 2 - d[1] / 2)
    t4 = t[:, :, ys:ys + d[1], xs:xs + d[0]]
    return [t, np.ascontiguousarray(t2), np.ascontiguousarray(t4)]


def track(dpt, com, size, cam, fx, fy, refine_fn, dsize=(128, 128), use_cv2=False):
    """handdetector.py:511-533 with doHandSize=False.  ``refine_fn([x0, x1, x2]) -> (1, 3)`` float32 is the
    deterministic ScaleNet forward.  Returns the refined com (float32 image coordinates)."""
    rs = resize_nn_cv2 if use_cv2 else resize_nn
    xstart, xend, ystart, yend, zstart, zend = com_to_bounds(com, size, fx, fy)
    cropped = get_crop(dpt, xstart, xend, ystart, yend, zstart, zend)
    rz = rs(cropped, dsize)
    jts = np.asarray(refine_fn(refine_inputs(rz, size, com)), f32)
    off3d = jts[0] * f32(size[2] / 2.)                       # float32 array * python float -> float32
    newCom3D = (off3d + cam.jointImgTo3D(com)).astype(f32)
    new_com = cam.joint3DToImg(newCom3D)
    if np.allclose(new_com, 0.):
        new_com[2] = cropped[cropped.shape[0] // 2, cropped.shape[1] // 2]
    return new_com


# --------------------------------------------------------------------------------------
# cropArea3D (docom=False) + the pipeline's normalisation
# --------------------------------------------------------------------------------------
def crop_area_3d(dpt, com, size, fx, fy, dsize=(128, 128), ndvalue=None, use_cv2=False):
    """handdetector.py:382-490 with docom=False.  Returns (crop float32 (dsize[1], dsize[0]), M 3x3, com)."""
    rs = resize_nn_cv2 if use_cv2 else resize_nn
    xstart, xend, ystart, yend, zstart, zend = com_to_bounds(com, size, fx, fy)
    cropped = get_crop(dpt, xstart, xend, ystart, yend, zstart, zend)
    wb = (xend - xstart)
    hb = (yend - ystart)
    if wb > hb:
        sz = (dsize[0], hb * dsize[0] // wb)
    else:
        sz = (wb * dsize[1] // hb, dsize[1])
    trans = np.eye(3)
    trans[0, 2] = -xstart
    trans[1, 2] = -ystart
    if cropped.shape[0] > cropped.shape[1]:
        scale = np.eye(3) * sz[1] / float(cropped.shape[0])
    else:
        scale = np.eye(3) * sz[0] / float(cropped.shape[1])
    scale[2, 2] = 1
    rz = rs(cropped, sz)
    nd = nd_value(dpt) if ndvalue is None else ndvalue
    ret = np.ones((dsize[1], dsize[0]), f32) * f32(nd)
    xs = int(np.floor(dsize[0] / 2. - rz.shape[1] / 2.))
    xe = int(xs + rz.shape[1])
    ys = int(np.floor(dsize[1] / 2. - rz.shape[0] / 2.))
    ye = int(ys + rz.shape[0])
    ret[ys:ye, xs:xe] = rz
    off = np.eye(3)
    off[0, 2] = xs
    off[1, 2] = ys
    return ret, np.dot(off, np.dot(scale, trans)), com


def pipeline_detect(dpt, loc, cube, cam, fx, fy, dsize=(128, 128), ndvalue=None, use_cv2=False):
    """realtimehandposepipeline.py:324-333: crop at ``loc`` and normalise with the crop centre's depth."""
    crop, M, com = crop_area_3d(dpt, loc, cube, fx, fy, dsize, ndvalue, use_cv2)
    com3D = cam.jointImgTo3D(com)
    sc = (cube[2] / 2.)
    crop[crop == 0] = f32(f64(com3D[2]) + sc)
    crop.clip(f64(com3D[2]) - sc, f64(com3D[2]) + sc)        # result discarded in the reference too
    crop -= com3D[2]
    crop /= f32(sc)
    return crop, M, com3D


def estimate_pose(crop, pose_fn, right_hand=False):
    """realtimehandposepipeline.py:335-368 without the invX/invY config switches."""
    inp = crop[None, None, :, ::-1].astype('float32') if right_hand else crop[None, None, :, :].astype('float32')
    jts = np.asarray(pose_fn(np.ascontiguousarray(inp)), f32)
    jj = jts[0].reshape((-1, 3)).copy()
    if right_hand:
        jj[:, 0] *= (-1.)
    return jj


def cascade_frame(dpt, lastcom, cube, cam, fx, fy, refine_fn, pose_fn, ndvalue=None, right_hand=False,
                  use_cv2=False):
    """One frame through detect (tracking branch) + estimatePose + the de-normalisation of
    processVideo (:197-198).  Returns dict(com, crop, M, com3D, pose (J,3) mm)."""
    loc = track(dpt, lastcom, cube, cam, fx, fy, refine_fn, use_cv2=use_cv2)
    crop, M, com3D = pipeline_detect(dpt, loc, cube, cam, fx, fy, ndvalue=ndvalue, use_cv2=use_cv2)
    jj = estimate_pose(crop, pose_fn, right_hand)
    pose = jj * f32(cube[2] / 2.) + com3D
    return dict(com=loc, crop=crop, M=M, com3D=com3D, pose=pose.astype(f32), pose_norm=jj)
