"""Batched, device-resident inference cascade: CoM refinement (ScaleNet) -> re-crop -> pose regression (ResNet /
PoseRegNet) -> joints in mm.

Reference, per frame (src/util/realtimehandposepipeline.py:296-370 ``detect`` + ``estimatePose`` and :197-198;
src/util/handdetector.py:511-533 ``track``, :634-676 ``refineCoM``, :382-490 ``cropArea3D``, :204-226
``comToBounds``): crop a cube around the last CoM, resize to 128x128, normalise, ScaleNet on the crop and its
64 / 32 centre crops -> 3D offset -> new CoM -> crop again (aspect-preserving resize pasted on a getNDValue canvas),
normalise with the new CoM's depth, pose net, ``pose * cube_z/2 + com3D``.

Here a whole batch of frames stays in HBM.  The host computes only the window geometry, vectorised over the batch
in the reference's dtype discipline (fp64 scalars, float32 stores; NumPy-1.x value-based casting made explicit),
and ships one 88-byte ``dpp_crop_rec`` per frame; ``dpp_recrop_fwd`` writes the networks' input tensors directly
(for one channel NHWC and NCHW coincide).  Per batch: 2 record uploads, 2 crop launches, 2 network forwards and
two small device->host reads (the (B,3) offsets, the (B,3J) poses).  No CPU pixel path exists."""
import ctypes as C
import numpy as np

from .lib import lib, DppError, CROP_REC_DTYPE, CROP_NORMALISE, CROP_CLAMP, CROP_MIRROR

f32 = np.float32
f64 = np.float64


# -----------------------------------------------------------------------------------------
# vectorised host geometry (reference dtype discipline)
# -----------------------------------------------------------------------------------------
def bounds_batch(coms, size, fx, fy):
    """handdetector.py:204-226 for n CoMs at once.  coms (n,3) float32 or float64 (u, v, d); returns int64
    xstart, xend, ystart, yend and float64 zstart, zend.  A CoM at depth ~0 is ill-defined (the reference falls
    back to the middle of the frame with a warning): rejected here."""
    c = np.asarray(coms)
    if c.ndim != 2 or c.shape[1] != 3:
        raise ValueError("coms must be (n, 3)")
    if np.isclose(c[:, 2], 0.).any():
        raise ValueError("CoM ill-defined (depth 0): detect the hand first")
    if c.dtype == f32:
        p0 = (c[:, 0] * c[:, 2]).astype(f64)          # float32 products, as com[0]*com[2] on a float32 com
        p1 = (c[:, 1] * c[:, 2]).astype(f64)
    else:
        c = c.astype(f64)
        p0 = c[:, 0] * c[:, 2]
        p1 = c[:, 1] * c[:, 2]
    c2 = c[:, 2].astype(f64)
    s0, s1, s2 = f64(size[0]) / 2., f64(size[1]) / 2., f64(size[2]) / 2.
    zstart = c2 - s2
    zend = c2 + s2
    xstart = np.floor((p0 / fx - s0) / c2 * fx + 0.5).astype(np.int64)
    xend = np.floor((p0 / fx + s0) / c2 * fx + 0.5).astype(np.int64)
    ystart = np.floor((p1 / fy - s1) / c2 * fy + 0.5).astype(np.int64)
    yend = np.floor((p1 / fy + s1) / c2 * fy + 0.5).astype(np.int64)
    return xstart, xend, ystart, yend, zstart, zend


def img_to_3d_batch(di, coms):
    """importer.jointImgTo3D (data/importers.py:80-98 / :756-770 / :1187-1201) on (n,3); float32 result."""
    c = np.asarray(coms).astype(f64)
    ret = np.zeros((c.shape[0], 3), f32)
    ret[:, 0] = (c[:, 0] - di.ux) * c[:, 2] / di.fx
    if di.flip_y:
        ret[:, 1] = (di.uy - c[:, 1]) * c[:, 2] / di.fy
    else:
        ret[:, 1] = (c[:, 1] - di.uy) * c[:, 2] / di.fy
    ret[:, 2] = np.asarray(coms)[:, 2]
    return ret


def to_img_batch(di, pts):
    """importer.joint3DToImg (data/importers.py:100-119 / :772-793 / :1203-1224) on float32 (n,3)."""
    p = np.asarray(pts)
    if p.dtype != f32:
        raise ValueError("to_img_batch expects float32 points (the reference's newCom3D is float32)")
    ret = np.zeros((p.shape[0], 3), f32)
    nz = p[:, 2] != 0.
    z = np.where(nz, p[:, 2], f32(1.))
    q0 = (p[:, 0] / z).astype(f64)                    # float32 / float32 first, then the python floats
    q1 = (p[:, 1] / z).astype(f64)
    u = q0 * di.fx + di.ux
    v = (di.uy - q1 * di.fy) if di.flip_y else (q1 * di.fy + di.uy)
    ret[:, 0] = np.where(nz, u, di.ux)
    ret[:, 1] = np.where(nz, v, di.uy)
    ret[:, 2] = np.where(nz, p[:, 2], f32(0.))
    return ret


def _check_windows(xstart, xend, ystart, yend, Hf, Wf):
    wb, hb = xend - xstart, yend - ystart
    if (wb <= 0).any() or (hb <= 0).any():
        raise ValueError("empty crop window")
    if (xstart >= Wf).any() or (xend <= 0).any() or (ystart >= Hf).any() or (yend <= 0).any():
        raise ValueError("crop window entirely outside the frame")
    return wb, hb


def refine_records(coms, size, fx, fy, frame_shape, dsize=(128, 128), src_index=None):
    """Records of the CoM-refinement crop (track, handdetector.py:520-527 + refineCoM :640-647): the window is
    stretched to the full dsize, normalised with the CURRENT com and clamped to the cube."""
    n = len(coms)
    xstart, xend, ystart, yend, zstart, zend = bounds_batch(coms, size, fx, fy)
    wb, hb = _check_windows(xstart, xend, ystart, yend, frame_shape[0], frame_shape[1])
    rec = np.zeros(n, dtype=CROP_REC_DTYPE)
    rec['src_index'] = np.arange(n) if src_index is None else src_index
    rec['xstart'], rec['ystart'], rec['wb'], rec['hb'] = xstart, ystart, wb, hb
    rec['rw'], rec['rh'] = dsize[0], dsize[1]
    rec['flags'] = CROP_NORMALISE | CROP_CLAMP
    rec['zstart'], rec['zend'] = zstart.astype(f32), zend.astype(f32)
    c2 = np.asarray(coms)[:, 2]
    rec['hi'] = (c2.astype(f64) + f64(size[2]) / 2.).astype(f32)
    rec['lo'] = (c2.astype(f64) - f64(size[2]) / 2.).astype(f32)
    rec['comz'] = c2.astype(f32)
    rec['half'] = f32(f64(size[2]) / 2.)
    rec['ifx'] = 1. / (f64(dsize[0]) / wb.astype(f64))
    rec['ify'] = 1. / (f64(dsize[1]) / hb.astype(f64))
    return rec


def pose_records(coms, size, fx, fy, di, frame_shape, ndvalue, dsize=(128, 128), mirror=False, src_index=None):
    """Records of the pose net's crop (cropArea3D with docom=False, handdetector.py:403-476, and the pipeline's
    normalisation, realtimehandposepipeline.py:327-332).  Returns (records, M (n,3,3) float64, com3D (n,3) f32)."""
    n = len(coms)
    xstart, xend, ystart, yend, zstart, zend = bounds_batch(coms, size, fx, fy)
    wb, hb = _check_windows(xstart, xend, ystart, yend, frame_shape[0], frame_shape[1])
    wide = wb > hb
    rw = np.where(wide, dsize[0], wb * dsize[1] // hb)
    rh = np.where(wide, hb * dsize[0] // wb, dsize[1])
    if (rw <= 0).any() or (rh <= 0).any():
        raise ValueError("degenerate crop aspect")
    px = np.floor(dsize[0] / 2. - rw / 2.).astype(np.int64)
    py = np.floor(dsize[1] / 2. - rh / 2.).astype(np.int64)
    com3D = img_to_3d_batch(di, coms)
    sc = f64(size[2]) / 2.
    rec = np.zeros(n, dtype=CROP_REC_DTYPE)
    rec['src_index'] = np.arange(n) if src_index is None else src_index
    rec['xstart'], rec['ystart'], rec['wb'], rec['hb'] = xstart, ystart, wb, hb
    rec['rw'], rec['rh'], rec['px'], rec['py'] = rw, rh, px, py
    rec['flags'] = CROP_NORMALISE | (CROP_MIRROR if mirror else 0)
    rec['zstart'], rec['zend'] = zstart.astype(f32), zend.astype(f32)
    rec['fill'] = np.asarray(ndvalue, f32)
    rec['hi'] = (com3D[:, 2].astype(f64) + sc).astype(f32)
    rec['lo'] = (com3D[:, 2].astype(f64) - sc).astype(f32)       # unused: the pipeline does not clamp
    rec['comz'] = com3D[:, 2]
    rec['half'] = f32(sc)
    rec['ifx'] = 1. / (rw.astype(f64) / wb.astype(f64))
    rec['ify'] = 1. / (rh.astype(f64) / hb.astype(f64))
    # M = off . scale . trans (handdetector.py:448-489)
    factor = np.where(hb > wb, rh / hb.astype(f64), rw / wb.astype(f64))
    M = np.zeros((n, 3, 3), f64)
    M[:, 0, 0] = factor
    M[:, 1, 1] = factor
    M[:, 2, 2] = 1.
    M[:, 0, 2] = factor * (-xstart) + px
    M[:, 1, 2] = factor * (-ystart) + py
    return rec, M, com3D


def nd_value(dpt):
    """HandDetector.getNDValue (handdetector.py:122-130) of a host frame: mode of the out-of-range pixels
    (scipy.stats.mode semantics: the smallest of equally frequent values)."""
    dpt = np.asarray(dpt)
    max_depth = min(1500, dpt.max())
    min_depth = max(10, dpt.min())
    lo = dpt[dpt < min_depth]
    hi = dpt[dpt > max_depth]
    sel = lo if lo.shape[0] > hi.shape[0] else hi
    if sel.shape[0] == 0:
        raise ValueError("frame has no undefined-depth pixels: pass ndvalue explicitly")
    first = sel[0]
    if (sel == first).all():             # the usual case - one marker value (0 or 32001): no sort needed
        return first
    vals, counts = np.unique(sel, return_counts=True)
    return vals[np.argmax(counts)]


# -----------------------------------------------------------------------------------------
# device side
# -----------------------------------------------------------------------------------------
def run_crop_records(frames_dev, recs_np, out0, out1=None, out2=None):
    """Launch dpp_recrop_fwd.  frames_dev: torch CUDA (F, Hf, Wf) f32; recs_np: CROP_REC_DTYPE array (n,);
    out0 (n, H, W[, 1]) and optional centre-crop outputs are torch CUDA tensors written in place."""
    import torch
    if not frames_dev.is_cuda:
        raise DppError("dpp_recrop_fwd needs device frames; there is no CPU fallback")
    recs_np = np.ascontiguousarray(recs_np, dtype=CROP_REC_DTYPE)
    n = int(recs_np.shape[0])
    F, Hf, Wf = [int(v) for v in frames_dev.shape]
    if n and (recs_np['src_index'].min() < 0 or recs_np['src_index'].max() >= F):
        raise ValueError("record src_index outside the frame batch")
    H, W = int(out0.shape[1]), int(out0.shape[2])
    for o, div in ((out0, 1), (out1, 2), (out2, 4)):
        if o is not None and (not o.is_contiguous() or o.dtype != torch.float32 or o.numel() != n * (H // div) * (W // div)):
            raise ValueError("output tensor has the wrong size / layout")
    if not frames_dev.is_contiguous() or frames_dev.dtype != torch.float32:
        raise ValueError("frames must be contiguous float32")
    if n == 0:
        return out0
    rec_dev = torch.from_numpy(recs_np.view(np.uint8).reshape(n, recs_np.dtype.itemsize)).to(frames_dev.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.dpp_recrop_fwd(C.c_void_p(frames_dev.data_ptr()), C.c_void_p(rec_dev.data_ptr()), C.c_void_p(out0.data_ptr()),
                       C.c_void_p(out1.data_ptr()) if out1 is not None else None,
                       C.c_void_p(out2.data_ptr()) if out2 is not None else None, n, Hf, Wf, H, W, st)
    return out0


def joint_errors(pred, gt):
    """Per-joint Euclidean errors and per-frame mean / max on the device (dpp_joint_errors; reference
    util/handpose_evaluation.py:92-181).  pred, gt: (n, J, 3) numpy or torch CUDA.  Returns torch CUDA tensors
    (err (n,J), frame_mean (n,), frame_max (n,))."""
    import torch
    if not torch.cuda.is_available():
        raise DppError("dpp_joint_errors needs a CUDA device; there is no CPU fallback")

    def dev(a):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a, f32)).cuda()
        return a.contiguous().float()
    p, g = dev(pred), dev(gt)
    if p.shape != g.shape or p.dim() != 3 or p.shape[2] != 3:
        raise ValueError("pred / gt must both be (n, J, 3)")
    n, J = int(p.shape[0]), int(p.shape[1])
    err = torch.empty((n, J), dtype=torch.float32, device=p.device)
    fmean = torch.empty((n,), dtype=torch.float32, device=p.device)
    fmax = torch.empty((n,), dtype=torch.float32, device=p.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.dpp_joint_errors(C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(err.data_ptr()),
                         C.c_void_p(fmean.data_ptr()), C.c_void_p(fmax.data_ptr()), n, J, st)
    return err, fmean, fmax


class Cascade(object):
    """CoM-refine + pose-regression cascade on a batch of frames (BASELINE config 5).

    poseNet / comrefNet: constructed reference-surface nets (``ResNet``/``PoseRegNet`` with the PCA prior layer,
    ``ScaleNet``); both must have been built with the same ``batchSize``.  fx, fy: the detector's focal lengths
    (the pipeline's ``config['fx']``, not necessarily the importer's); cube: ``config['cube']``."""

    def __init__(self, poseNet, comrefNet, di, fx, fy, cube, dsize=(128, 128)):
        import torch
        if not torch.cuda.is_available():
            raise DppError("dpp_b200.Cascade needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.poseNet, self.comrefNet, self.di = poseNet, comrefNet, di
        self.fx, self.fy, self.cube, self.dsize = fx, fy, tuple(cube), tuple(dsize)
        self.B = int(poseNet.cfgParams.batch_size)
        if comrefNet is not None:
            if int(comrefNet.cfgParams.batch_size) != self.B:
                raise ValueError("poseNet and comrefNet must share one batch size")
            if comrefNet.cfgParams.numInputs != 3:
                raise NotImplementedError("Number of inputs is {}".format(comrefNet.cfgParams.numInputs))
            comrefNet.setDeterministic()
            self.ref_eng = comrefNet._engine()
            if tuple(self.ref_eng.t_ins[0].shape[2:]) != (dsize[1], dsize[0]):
                raise ValueError("comrefNet input size differs from dsize")
        poseNet.setDeterministic()
        self.pose_eng = poseNet._engine()
        if tuple(self.pose_eng.t_ins[0].shape[1:]) != (1, dsize[1], dsize[0]):
            raise ValueError("poseNet input must be (B, 1, %d, %d)" % (dsize[1], dsize[0]))
        self.launches_per_batch = None

    def _pad(self, a, n):
        if n == self.B:
            return a
        return np.concatenate([a, np.repeat(a[-1:], self.B - n, axis=0)])

    def refine(self, frames_dev, coms):
        """track() with doHandSize=False for a batch: returns the refined CoMs, float32 (n,3) image coords."""
        n = len(coms)
        if n == 0 or n > self.B:
            raise ValueError("between 1 and batch_size frames per call")
        rec = refine_records(coms, self.cube, self.fx, self.fy, frames_dev.shape[1:], self.dsize)
        e = self.ref_eng
        run_crop_records(frames_dev, self._pad(rec, n), e.t_ins[0].buf, e.t_ins[1].buf, e.t_ins[2].buf)
        jts = e.forward_device(deterministic=True).cpu().numpy()[:n].reshape(n, -1)   # (n, 3) normalised offsets
        off3d = jts * f32(f64(self.cube[2]) / 2.)            # refineCoM: jts[0]*(size[2]/2.), float32
        new3d = (off3d + img_to_3d_batch(self.di, coms)).astype(f32)
        new_com = to_img_batch(self.di, new3d)
        if np.isclose(new_com, 0.).all(axis=1).any():
            # handdetector.py:526-527 replaces the depth by the crop's centre pixel; needs u = v = d = 0
            raise NotImplementedError("refined CoM collapsed to the origin")
        return new_com

    def crop(self, frames_dev, coms, ndvalue, right_hand=False):
        """cropArea3D + the pipeline's normalisation (+ estimatePose's mirroring) into the pose net's input buffer.
        Returns (M (n,3,3), com3D (n,3))."""
        n = len(coms)
        rec, M, com3D = pose_records(coms, self.cube, self.fx, self.fy, self.di, frames_dev.shape[1:], ndvalue,
                                     self.dsize, mirror=right_hand)
        run_crop_records(frames_dev, self._pad(rec, n), self.pose_eng.t_ins[0].buf)
        return M, com3D

    def run(self, frames, lastcoms, ndvalue=None, right_hand=False, return_crops=False):
        """frames: (n, Hf, Wf) float32 depth in mm, numpy (copied to the device here) or torch CUDA; lastcoms (n,3)
        image coordinates of the previous CoMs (float64 as ``detect`` returns them, or float32 from a previous
        call); ndvalue: the frames' undefined-depth value (scalar or (n,)), default ``getNDValue`` of every host
        frame.  Returns dict(pose (n,J,3) mm, pose_norm, com (n,3), com3D (n,3), M (n,3,3)[, crop])."""
        torch = self.torch
        n = len(lastcoms)
        if isinstance(frames, np.ndarray):
            if ndvalue is None:
                ndvalue = np.array([nd_value(f) for f in frames], f32)
            frames = torch.from_numpy(np.ascontiguousarray(frames, f32)).to(self.pose_eng.dev, non_blocking=True)
        elif ndvalue is None:
            raise ValueError("device-resident frames need an explicit ndvalue (the sensor's undefined-depth value)")
        coms = np.asarray(lastcoms)
        if self.comrefNet is not None:
            coms = self.refine(frames, coms)
        M, com3D = self.crop(frames, coms, ndvalue, right_hand)
        out = self.pose_eng.forward_device(deterministic=True)
        jts = out.cpu().numpy()[:n]
        jj = jts.reshape(n, -1, 3).copy()
        if right_hand:
            jj[:, :, 0] *= f32(-1.)
        pose = (jj * f32(self.cube[2]) / f32(2.) + com3D[:, None, :]).astype(f32)
        res = dict(pose=pose, pose_norm=jj, com=coms, com3D=com3D, M=M)
        if return_crops:
            res['crop'] = self.pose_eng.t_ins[0].buf[:n].reshape(n, self.dsize[1], self.dsize[0]).cpu().numpy()
        return res
