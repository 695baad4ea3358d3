"""HandDetector - the augmentation and cascade subset (reference: src/util/handdetector.py: comToBounds
:204-226, comToTransform :228-258, getCrop :260-296, resizeCrop :336-351, cropArea3D :382-490, track :511-533,
refineCoM :634-676, moveCoM :678-710, rotateHand :712-747, scaleHand :750-780, recropHand :782-803).  The geometry
(3x3 matrices, crop bounds, label transforms) stays on the host in fp64 exactly as the reference computes it; every
pixel operation runs on the device: the cv2.warpAffine / cv2.warpPerspective nearest-neighbour gathers, the
z-thresholds and the CoM normalisation of the augmentation in ``dpp_augment_fwd`` (csrc/augment.cu), the window /
padding / cv2.resize-NN / paste / normalisation of the cascade's crops in ``dpp_recrop_fwd`` (csrc/recrop.cu).
``aug_record`` turns one sample's draw into the ``dpp_aug_rec`` the augmentation kernel consumes (``NetTrainer``
batches those per macro batch); ``dpp_b200.cascade`` builds ``dpp_crop_rec`` batches for the cascade.

Contour-based hand detection (detect / estimateHandsize / calculateCoM) is out of scope (CPU, serial)."""
import math
import numpy as np

from data.transformations import rotatePoint2D
from dpp_b200.lib import AUG_REC_DTYPE

f32 = np.float32
f64 = np.float64


def invert3x3_cv(S):
    """cv::invert of a 3x3 CV_64F matrix (closed-form cofactors) - the inverse
    cv2.warpPerspective applies to the forward matrix; its last bits decide NN ties."""
    S = np.asarray(S, f64)
    d = S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0]) \
        + S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0])
    d = 1. / d
    t = np.empty(9, f64)
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


def rotation_inverse_affine(center, angle_deg):
    """cv2.getRotationMatrix2D(center, angle, 1) followed by the 2x3 inversion cv2.warpAffine
    performs; returns {i00, i01, b0, i10, i11, b1} (fp64)."""
    a = float(angle_deg) * (math.pi / 180.)
    alpha, beta = math.cos(a), math.sin(a)
    cx, cy = float(center[0]), float(center[1])
    m00, m01, m02 = alpha, beta, (1 - alpha) * cx - beta * cy
    m10, m11, m12 = -beta, alpha, beta * cx + (1 - alpha) * cy
    D = m00 * m11 - m01 * m10
    D = 1. / D if D != 0 else 0.
    i00, i11 = m11 * D, m00 * D
    i01, i10 = -m01 * D, -m10 * D
    b0 = -i00 * m02 - i01 * m12
    b1 = -i10 * m02 - i11 * m12
    return np.array([i00, i01, b0, i10, i11, b1, 0., 0., 0.], f64)


class HandDetector(object):
    RESIZE_BILINEAR = 0
    RESIZE_CV2_NN = 1
    RESIZE_CV2_LINEAR = 2

    def __init__(self, dpt, fx, fy, importer=None, refineNet=None):
        self.dpt = dpt
        self.maxDepth = min(1500, dpt.max()) if dpt is not None else 1500
        self.minDepth = max(10, dpt.min()) if dpt is not None else 10
        self.fx = fx
        self.fy = fy
        self.refineNet = refineNet
        self.importer = importer
        self.resizeMethod = self.RESIZE_CV2_NN

    # -- crop geometry ------------------------------------------------------------------
    def comToBounds(self, com, size):
        """handdetector.py:204-226; ``com[0]*com[2]`` is a float32 product in the reference."""
        if np.isclose(com[2], 0.):
            print("Warning: CoM ill-defined!")
            xstart = self.dpt.shape[0] // 4
            xend = xstart + self.dpt.shape[0] // 2
            ystart = self.dpt.shape[1] // 4
            yend = ystart + self.dpt.shape[1] // 2
            zstart = self.minDepth
            zend = self.maxDepth
        else:
            if np.asarray(com).dtype == f32:     # com out of joint3DToImg: the products are float32
                c2 = f64(f32(com[2]))
                p0 = f64(f32(com[0]) * f32(com[2]))
                p1 = f64(f32(com[1]) * f32(com[2]))
            else:                                # com out of detect() / a python tuple: float64 throughout
                c2 = f64(com[2])
                p0 = f64(com[0]) * c2
                p1 = f64(com[1]) * c2
            zstart = c2 - f64(size[2]) / 2.
            zend = c2 + f64(size[2]) / 2.
            xstart = int(np.floor((p0 / self.fx - f64(size[0]) / 2.) / c2 * self.fx + 0.5))
            xend = int(np.floor((p0 / self.fx + f64(size[0]) / 2.) / c2 * self.fx + 0.5))
            ystart = int(np.floor((p1 / self.fy - f64(size[1]) / 2.) / c2 * self.fy + 0.5))
            yend = int(np.floor((p1 / self.fy + f64(size[1]) / 2.) / c2 * self.fy + 0.5))
        return xstart, xend, ystart, yend, zstart, zend

    def comToTransform(self, com, size, dsize=(128, 128)):
        """handdetector.py:228-258 (py2 integer division; the sz[1]/sz[0] swap is the reference's)."""
        xstart, xend, ystart, yend, _, _ = self.comToBounds(com, size)
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
        xstart = int(np.floor(dsize[0] / 2. - sz[1] / 2.))
        ystart = int(np.floor(dsize[1] / 2. - sz[0] / 2.))
        off = np.eye(3)
        off[0, 2] = xstart
        off[1, 2] = ystart
        return np.dot(off, np.dot(scale, trans))

    # -- one sample's augmentation record + labels ---------------------------------------
    def aug_record(self, src_index, mode_name, off, rot, sc, com, cube, M, gt3Dcrop, raw=False):
        """Geometry of NetTrainer.augmentCrop for one sample (nettrainer.py:948-995 + the
        HandDetector methods it calls).  Returns (record, curLabel (J,3) f32, cube', com', M')."""
        di = self.importer
        rec = np.zeros((), dtype=AUG_REC_DTYPE)
        rec['src_index'] = src_index
        half = f32(f64(cube[2]) / 2.)
        rec['half_old'] = half
        rec['comz_old'] = f32(com[2])
        new_com, new_cube, Mnew = com, cube, M
        mode = 0
        joints = gt3Dcrop
        if mode_name == 'com' and not np.allclose(off, 0.):
            new_com = di.joint3DToImg(di.jointImgTo3D(com).astype(f64) + np.asarray(off, f64))
            if not (np.allclose(com[2], 0.) or np.allclose(new_com[2], 0.)):
                Mnew = self.comToTransform(new_com, cube, (128, 128))
                H = np.dot(Mnew, np.linalg.inv(M))           # inv of the float32 M, as in moveCoM
                rec['m'] = invert3x3_cv(H)
                _, _, _, _, zs, ze = self.comToBounds(new_com, cube)
                rec['zstart'], rec['zend'] = f32(zs), f32(ze)
                mode = 2
            joints = ((gt3Dcrop + di.jointImgTo3D(com)) - di.jointImgTo3D(new_com)).astype(f32)
        elif mode_name == 'rot' and not np.allclose(rot, 0.):
            rot = np.mod(rot, 360)
            rec['m'] = rotation_inverse_affine((64, 64), -rot)
            mode = 1
            com3D = di.jointImgTo3D(com)
            joint_2D = di.joints3DToImg((gt3Dcrop + com3D).astype(f32))
            data_2D = np.zeros_like(joint_2D)
            for k in range(data_2D.shape[0]):
                data_2D[k] = rotatePoint2D(joint_2D[k], com[0:2], rot)
            joints = (di.jointsImgTo3D(data_2D) - com3D).astype(f32)
        elif mode_name == 'sc' and not np.allclose(sc, 1.):
            new_cube = [f64(s) * f64(sc) for s in cube]
            if not np.allclose(com[2], 0.):
                Mnew = self.comToTransform(com, new_cube, (128, 128))
                H = np.dot(Mnew, np.linalg.inv(M))
                rec['m'] = invert3x3_cv(H)
                _, _, _, _, zs, ze = self.comToBounds(com, cube)     # z-threshold with the OLD cube
                rec['zstart'], rec['zend'] = f32(zs), f32(ze)
                mode = 2
        elif mode_name not in ('com', 'rot', 'sc', 'none'):
            raise NotImplementedError()
        rec['mode'] = mode + (16 if raw else 0)
        rec['bg'] = f32(f64(new_com[2]) + f64(new_cube[2]) / 2.)
        rec['lo'] = f32(f64(new_com[2]) - f64(new_cube[2]) / 2.)
        rec['comz_new'] = f32(new_com[2])
        rec['half_new'] = f32(f64(new_cube[2]) / 2.)
        curLabel = (joints / f32(f64(new_cube[2]) / 2.)).astype(f32)
        return rec, curLabel, np.asarray(new_cube), new_com, Mnew

    # -- reference-signature methods on single images (slow path: one launch per call) ----
    def _warp_raw(self, dpt, rec):
        import torch
        from dpp_b200.augment import run_records
        rec = rec.copy()
        rec['src_index'] = 0
        rec['half_old'] = 1.0
        rec['comz_old'] = 0.0
        rec['mode'] = (int(rec['mode']) & 15) | 16
        out = run_records(np.ascontiguousarray(dpt, f32)[None], np.array([rec]))
        return out[0]

    def moveCoM(self, dpt, cube, com, off, joints3D, M, pad_value=0):
        if np.allclose(off, 0.):
            return dpt, joints3D, com, M
        rec, _, _, new_com, Mnew = self.aug_record(0, 'com', off, 0., 1., com, cube, M, joints3D, raw=True)
        new_dpt = self._warp_raw(dpt, rec) if (int(rec['mode']) & 15) == 2 else dpt
        di = self.importer
        new_joints3D = ((joints3D + di.jointImgTo3D(com)) - di.jointImgTo3D(new_com)).astype(f32)
        return new_dpt, new_joints3D, new_com, Mnew

    def rotateHand(self, dpt, cube, com, rot, joints3D, pad_value=0):
        if np.allclose(rot, 0.):
            return dpt, joints3D, rot
        rec, lab, _, _, _ = self.aug_record(0, 'rot', np.zeros(3), rot, 1., com, cube, np.eye(3, dtype=f32), joints3D,
                                            raw=True)
        new_dpt = self._warp_raw(dpt, rec)
        # labels in mm: undo the /(cube_z/2) of aug_record exactly is not possible in f32, recompute
        rotm = np.mod(rot, 360)
        com3D = self.importer.jointImgTo3D(com)
        joint_2D = self.importer.joints3DToImg((joints3D + com3D).astype(f32))
        data_2D = np.zeros_like(joint_2D)
        for k in range(data_2D.shape[0]):
            data_2D[k] = rotatePoint2D(joint_2D[k], com[0:2], rotm)
        return new_dpt, (self.importer.jointsImgTo3D(data_2D) - com3D).astype(f32), rotm

    def scaleHand(self, dpt, cube, com, sc, joints3D, M, pad_value=0):
        if np.allclose(sc, 1.):
            return dpt, joints3D, cube, M
        rec, _, new_cube, _, Mnew = self.aug_record(0, 'sc', np.zeros(3), 0., sc, com, cube, M, joints3D, raw=True)
        new_dpt = self._warp_raw(dpt, rec) if (int(rec['mode']) & 15) == 2 else dpt
        return new_dpt, joints3D, list(new_cube), Mnew

    # -- cascade: crop around a CoM (single-frame, reference-signature methods; one launch per call) ------------
    def getNDValue(self):
        """handdetector.py:122-130: mode of the undefined-depth pixels of ``self.dpt``."""
        from dpp_b200.cascade import nd_value
        return nd_value(self.dpt)

    def _frame_dev(self):
        import torch
        from dpp_b200.lib import DppError
        if not torch.cuda.is_available():
            raise DppError("HandDetector crops run in dpp_recrop_fwd; there is no CPU fallback")
        if getattr(self, '_dpt_dev', None) is None:
            self._dpt_dev = torch.from_numpy(np.ascontiguousarray(self.dpt, f32)[None]).cuda()
        return self._dpt_dev

    def cropArea3D(self, com=None, size=(250, 250, 250), dsize=(128, 128), docom=False):
        """handdetector.py:382-490 for a given CoM (``docom=False``, as the realtime pipeline calls it): returns
        (crop float32 dsize, M 3x3, com).  ``com=None`` / ``docom=True`` need calculateCoM (out of scope)."""
        import torch
        from dpp_b200.cascade import pose_records, run_crop_records
        if len(size) != 3 or len(dsize) != 2:
            raise ValueError("Size must be 3D and dsize 2D bounding box")
        if com is None or docom is True:
            raise NotImplementedError("CoM estimation from the depth map is outside the B200 path")
        di = self.importer
        if di is None:        # the raw crop does not depend on the camera model (only the caller's normalisation does)
            import types
            di = types.SimpleNamespace(ux=0., uy=0., fx=1., fy=1., flip_y=False)
        coms = np.asarray(com)[None]
        rec, M, _ = pose_records(coms, size, self.fx, self.fy, di, self.dpt.shape, self.getNDValue(), dsize)
        rec['flags'] = 0                                     # raw crop in mm: the caller normalises
        frames = self._frame_dev()
        out = torch.empty((1, dsize[1], dsize[0]), dtype=torch.float32, device=frames.device)
        run_crop_records(frames, rec, out)
        return out[0].cpu().numpy(), M[0], com

    def refineCoM(self, cropped, size, com):
        """handdetector.py:634-676: normalise the dsize crop, take its 1/2 and 1/4 centre crops, run the
        refinement net; returns the 3D offset in mm."""
        import torch
        from dpp_b200.cascade import run_crop_records
        from dpp_b200.lib import CROP_REC_DTYPE, CROP_NORMALISE, CROP_CLAMP
        if self.refineNet.cfgParams.numInputs != 3:
            raise NotImplementedError("Number of inputs is {}".format(self.refineNet.cfgParams.numInputs))
        cropped = np.ascontiguousarray(cropped, f32)
        h, w = cropped.shape
        rec = np.zeros(1, dtype=CROP_REC_DTYPE)              # identity window over the given crop
        rec['wb'], rec['hb'], rec['rw'], rec['rh'] = w, h, w, h
        rec['flags'] = CROP_NORMALISE | CROP_CLAMP
        rec['zstart'], rec['zend'] = -np.inf, np.inf
        rec['hi'] = f32(f64(com[2]) + f64(size[2]) / 2.)
        rec['lo'] = f32(f64(com[2]) - f64(size[2]) / 2.)
        rec['comz'] = f32(com[2])
        rec['half'] = f32(f64(size[2]) / 2.)
        rec['ifx'] = rec['ify'] = 1.
        dev = torch.from_numpy(cropped[None]).cuda()
        x0 = torch.empty((1, h, w), dtype=torch.float32, device=dev.device)
        x1 = torch.empty((1, h // 2, w // 2), dtype=torch.float32, device=dev.device)
        x2 = torch.empty((1, h // 4, w // 4), dtype=torch.float32, device=dev.device)
        run_crop_records(dev, rec, x0, x1, x2)
        jts = self.refineNet.computeOutput([t.cpu().numpy()[:, None] for t in (x0, x1, x2)])
        return jts[0] * f32(f64(size[2]) / 2.)

    def track(self, com, size=(250, 250, 250), dsize=(128, 128), doHandSize=True):
        """handdetector.py:511-533: refine the CoM with the refinement net.  Hand-size estimation from contours
        (``doHandSize=True``) is CPU contour work outside the B200 path."""
        import torch
        from dpp_b200.cascade import refine_records, run_crop_records
        if doHandSize is True:
            raise NotImplementedError("hand-size estimation (cv2 contours) is outside the B200 path")
        if self.refineNet is None or self.importer is None:
            raise RuntimeError("Need refineNet for this")
        rec = refine_records(np.asarray(com)[None], size, self.fx, self.fy, self.dpt.shape, dsize)
        rec['flags'] = 0                                     # raw window stretched to dsize, in mm
        frames = self._frame_dev()
        rz = torch.empty((1, dsize[1], dsize[0]), dtype=torch.float32, device=frames.device)
        run_crop_records(frames, rec, rz)
        newCom3D = (self.refineCoM(rz[0].cpu().numpy(), size, com) + self.importer.jointImgTo3D(com)).astype(f32)
        com = self.importer.joint3DToImg(newCom3D)
        if np.allclose(com, 0.):
            x0, y0, wb, hb = int(rec['xstart'][0]), int(rec['ystart'][0]), int(rec['wb'][0]), int(rec['hb'][0])
            yy, xx = y0 + hb // 2, x0 + wb // 2
            inside = 0 <= yy < self.dpt.shape[0] and 0 <= xx < self.dpt.shape[1]
            com[2] = self.dpt[yy, xx] if inside else 0.
        return com, size

    # -- pose sampling for the PCA prior ----------------------------------------------------------------------------
    POSE_MODE_CODES = {'none': 0, 'rot': 1, 'sc': 2, 'com': 3, 'rot+com': 4, 'com+rot': 4, 'rot+com+sc': 5,
                       'rot+sc+com': 5}

    @staticmethod
    def sampleRandomPoses(importer, rng, base_poses, base_com, base_cube, num_poses, aug_modes, retall=False,
                          rot3D=False, sigma_com=None, sigma_sc=None, rot_range=None):
        """handdetector.py:805-909.  The five random arrays are drawn on the host from ``rng`` in the reference's
        order (:837-841, so the stream stays compatible); the per-pose arithmetic runs in ``dpp_sample_poses``
        (csrc/poses.cu).  ``rot3D=True`` needs the transforms3d package and is never used by the entry scripts."""
        import ctypes as C
        import torch
        from dpp_b200.lib import lib, DppError
        if rot3D is not False:
            raise NotImplementedError("rot3D (transforms3d euler2mat) is not part of the B200 path")
        sigma_com = 5. if sigma_com is None else sigma_com
        sigma_sc = 0.02 if sigma_sc is None else sigma_sc
        rot_range = 180. if rot_range is None else rot_range
        all_modes = ['none', 'rot', 'sc', 'com', 'rot+com', 'com+rot', 'rot+com+sc', 'rot+sc+com', 'sc+rot+com',
                     'sc+com+rot', 'com+sc+rot', 'com+rot+sc']
        assert all([aug_modes[i] in all_modes for i in range(len(aug_modes))])
        n = int(num_poses)
        modes = rng.randint(0, len(aug_modes), n)
        ridxs = rng.randint(0, base_poses.shape[0], n)
        off = rng.randn(n, 3) * sigma_com
        sc = np.fabs(rng.randn(n) * sigma_sc + 1.)
        rot = rng.uniform(-rot_range, rot_range, size=(n, 3))
        if not torch.cuda.is_available():
            raise DppError("sampleRandomPoses runs in dpp_sample_poses; there is no CPU fallback")
        only_none = (aug_modes == ['none'])
        if only_none:                                  # :843-847: the base poses themselves, normalised
            n = base_poses.shape[0]
            codes, ridxs = np.zeros(n, np.int32), np.arange(n, dtype=np.int32)
            off, sc, rot = np.zeros((n, 3)), np.ones(n), np.zeros((n, 3))
        else:
            # the four orderings the reference compares as list == str (:893) end in NotImplementedError there too
            names = [aug_modes[m] for m in np.unique(modes)]
            bad = [m for m in names if m not in HandDetector.POSE_MODE_CODES]
            if bad:
                raise NotImplementedError(bad[0])
            table = np.array([HandDetector.POSE_MODE_CODES.get(m, -1) for m in aug_modes], np.int32)
            codes = table[modes]
        alpha = rot[:, 0] * np.pi / 180.
        cos_sin = np.stack([np.cos(alpha), np.sin(alpha)], axis=1)
        dev = torch.device('cuda', torch.cuda.current_device())

        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        J = int(base_poses.shape[1])
        t_in = [up(base_poses, f32), up(base_com, f32), up(base_cube, f32), up(codes, np.int32), up(ridxs, np.int32),
                up(off, f64), up(sc, f64), up(cos_sin, f64)]
        new_poses = torch.empty((n, J, 3), dtype=torch.float32, device=dev)
        new_com = torch.empty((n, 3), dtype=torch.float32, device=dev)
        new_cube = torch.empty((n, 3), dtype=torch.float32, device=dev)
        lib.dpp_sample_poses(*[C.c_void_p(t.data_ptr()) for t in t_in], float(importer.fx), float(importer.fy),
                             float(importer.ux), float(importer.uy), 1 if getattr(importer, 'flip_y', False) else 0,
                             C.c_void_p(new_poses.data_ptr()), C.c_void_p(new_com.data_ptr()),
                             C.c_void_p(new_cube.data_ptr()), n, J, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        poses = new_poses.cpu().numpy().astype(base_poses.dtype, copy=False)
        if only_none:
            return (poses, base_com, base_cube) if retall is True else poses
        if retall is True:
            return poses, new_com.cpu().numpy().astype(base_poses.dtype, copy=False), \
                new_cube.cpu().numpy().astype(base_poses.dtype, copy=False), rot
        return poses

    # -- a whole batch of augmentation records at once -------------------------------------------------------------
    def _bounds_batch(self, com, size):
        """comToBounds for float32 CoMs (n,3) and per-sample sizes (n,3); non-degenerate depths only."""
        com = np.asarray(com, f32)
        c2 = com[:, 2].astype(f64)
        p0 = (com[:, 0] * com[:, 2]).astype(f64)               # float32 products
        p1 = (com[:, 1] * com[:, 2]).astype(f64)
        s = np.asarray(size).astype(f64) / 2.
        zstart, zend = c2 - s[:, 2], c2 + s[:, 2]
        xstart = np.floor((p0 / self.fx - s[:, 0]) / c2 * self.fx + 0.5).astype(np.int64)
        xend = np.floor((p0 / self.fx + s[:, 0]) / c2 * self.fx + 0.5).astype(np.int64)
        ystart = np.floor((p1 / self.fy - s[:, 1]) / c2 * self.fy + 0.5).astype(np.int64)
        yend = np.floor((p1 / self.fy + s[:, 1]) / c2 * self.fy + 0.5).astype(np.int64)
        return xstart, xend, ystart, yend, zstart, zend

    def _transform_batch(self, com, size, dsize=(128, 128)):
        """comToTransform for n samples: (n,3,3) float64."""
        xstart, xend, ystart, yend, _, _ = self._bounds_batch(com, size)
        wb, hb = xend - xstart, yend - ystart
        wide = wb > hb
        s = np.where(wide, dsize[0] / wb.astype(f64), dsize[1] / hb.astype(f64))
        sz0 = np.where(wide, dsize[0], wb * dsize[1] // hb)
        sz1 = np.where(wide, hb * dsize[0] // wb, dsize[1])
        xs = np.floor(dsize[0] / 2. - sz1 / 2.)                # the reference's sz[1] / sz[0] swap
        ys = np.floor(dsize[1] / 2. - sz0 / 2.)
        Mt = np.zeros((len(wb), 3, 3), f64)
        Mt[:, 0, 0] = s
        Mt[:, 1, 1] = s
        Mt[:, 2, 2] = 1.
        Mt[:, 0, 2] = s * (-xstart) + xs
        Mt[:, 1, 2] = s * (-ystart) + ys
        return Mt

    @staticmethod
    def _invert3x3_cv_batch(S):
        d = S[:, 0, 0] * (S[:, 1, 1] * S[:, 2, 2] - S[:, 1, 2] * S[:, 2, 1]) \
            - S[:, 0, 1] * (S[:, 1, 0] * S[:, 2, 2] - S[:, 1, 2] * S[:, 2, 0]) \
            + S[:, 0, 2] * (S[:, 1, 0] * S[:, 2, 1] - S[:, 1, 1] * S[:, 2, 0])
        d = 1. / d
        t = np.empty((S.shape[0], 9), f64)
        t[:, 0] = (S[:, 1, 1] * S[:, 2, 2] - S[:, 1, 2] * S[:, 2, 1]) * d
        t[:, 1] = (S[:, 0, 2] * S[:, 2, 1] - S[:, 0, 1] * S[:, 2, 2]) * d
        t[:, 2] = (S[:, 0, 1] * S[:, 1, 2] - S[:, 0, 2] * S[:, 1, 1]) * d
        t[:, 3] = (S[:, 1, 2] * S[:, 2, 0] - S[:, 1, 0] * S[:, 2, 2]) * d
        t[:, 4] = (S[:, 0, 0] * S[:, 2, 2] - S[:, 0, 2] * S[:, 2, 0]) * d
        t[:, 5] = (S[:, 0, 2] * S[:, 1, 0] - S[:, 0, 0] * S[:, 1, 2]) * d
        t[:, 6] = (S[:, 1, 0] * S[:, 2, 1] - S[:, 1, 1] * S[:, 2, 0]) * d
        t[:, 7] = (S[:, 0, 1] * S[:, 2, 0] - S[:, 0, 0] * S[:, 2, 1]) * d
        t[:, 8] = (S[:, 0, 0] * S[:, 1, 1] - S[:, 0, 1] * S[:, 1, 0]) * d
        return t

    def _to3d(self, pts):
        """importer.jointImgTo3D on (..., 3): float64 expression, float32 store."""
        di = self.importer
        p = np.asarray(pts)
        c = p.astype(f64)
        ret = np.zeros(p.shape, f32)
        ret[..., 0] = (c[..., 0] - di.ux) * c[..., 2] / di.fx
        if di.flip_y:
            ret[..., 1] = (di.uy - c[..., 1]) * c[..., 2] / di.fy
        else:
            ret[..., 1] = (c[..., 1] - di.uy) * c[..., 2] / di.fy
        ret[..., 2] = p[..., 2]
        return ret

    def _toimg(self, pts):
        """importer.joint3DToImg on (..., 3): the quotient in the sample's own dtype, then float64, float32 store."""
        di = self.importer
        p = np.asarray(pts)
        nz = p[..., 2] != 0.
        z = np.where(nz, p[..., 2], p.dtype.type(1.))
        q0 = (p[..., 0] / z).astype(f64)
        q1 = (p[..., 1] / z).astype(f64)
        ret = np.zeros(p.shape, f32)
        ret[..., 0] = np.where(nz, q0 * di.fx + di.ux, di.ux)
        ret[..., 1] = np.where(nz, (di.uy - q1 * di.fy) if di.flip_y else (q1 * di.fy + di.uy), di.uy)
        ret[..., 2] = np.where(nz, p[..., 2], 0.)
        return ret

    def aug_records_batch(self, src_index, mode_names, off, rot, sc, com, cube, M, gt3Dcrop):
        """``aug_record`` for n samples at once - same results bit for bit (tests/test_host_logic.py), ~25x faster than
        the per-sample loop, which matters because one B200 consumes > 25 000 records per second.  mode_names: n
        strings; off (n,3), rot (n,), sc (n,) float64 draws; com (n,3) float32 image coordinates; cube (n,3) float32;
        M (n,3,3) float32; gt3Dcrop (n,J,3) float32.  Returns (records (n,), curLabel (n,J,3) float32).
        Everything elementwise is vectorised; the few operations whose result bits depend on the library routine
        (3x3 ``numpy.dot`` / ``numpy.linalg.inv``, ``math.cos`` / ``numpy.cos`` on scalars) stay per-sample calls."""
        n = len(src_index)
        com = np.asarray(com, f32)
        cube = np.asarray(cube, f32)
        M = np.asarray(M, f32)
        gt = np.asarray(gt3Dcrop, f32)
        off = np.asarray(off, f64).reshape(n, 3)
        rot = np.asarray(rot, f64).reshape(n)
        sc = np.asarray(sc, f64).reshape(n)
        names = np.asarray(mode_names)
        for m in np.unique(names):
            if m not in ('com', 'rot', 'sc', 'none'):
                raise NotImplementedError()
        if np.isclose(com[:, 2], 0.).any():                   # ill-defined CoMs take the per-sample path
            out = [self.aug_record(src_index[i], names[i], off[i], rot[i], sc[i], com[i], cube[i], M[i], gt[i])
                   for i in range(n)]
            return np.array([o[0] for o in out]), np.stack([o[1] for o in out])
        rec = np.zeros(n, dtype=AUG_REC_DTYPE)
        rec['src_index'] = src_index
        rec['half_old'] = (cube[:, 2].astype(f64) / 2.).astype(f32)
        rec['comz_old'] = com[:, 2]
        new_com = com.copy()
        new_cube = cube.astype(f64)
        joints = gt.copy()
        mode = np.zeros(n, np.int32)

        def homography(idx, Mnew):
            for k, i in enumerate(idx):
                H = np.dot(Mnew[k], np.linalg.inv(M[i]))
                rec['m'][i] = invert3x3_cv(H)

        is_com = (names == 'com') & ~np.isclose(off, 0.).all(axis=1)
        if is_com.any():
            idx = np.nonzero(is_com)[0]
            c3 = self._to3d(com[idx])
            nc = self._toimg(c3.astype(f64) + off[idx])
            new_com[idx] = nc
            ok = ~np.isclose(nc[:, 2], 0.)
            if ok.any():
                j = idx[ok]
                homography(j, self._transform_batch(nc[ok], cube[j]))
                _, _, _, _, zs, ze = self._bounds_batch(nc[ok], cube[j])
                rec['zstart'][j], rec['zend'][j] = zs.astype(f32), ze.astype(f32)
                mode[j] = 2
            joints[idx] = ((gt[idx] + c3[:, None, :]) - self._to3d(nc)[:, None, :]).astype(f32)
        is_rot = (names == 'rot') & ~np.isclose(rot, 0.)
        if is_rot.any():
            idx = np.nonzero(is_rot)[0]
            r = np.mod(rot[idx], 360)
            for k, i in enumerate(idx):
                rec['m'][i] = rotation_inverse_affine((64, 64), -r[k])
            mode[idx] = 1
            c3 = self._to3d(com[idx])
            j2d = self._toimg((gt[idx] + c3[:, None, :]).astype(f32))
            alpha = r * np.pi / 180.
            ca = np.array([np.cos(a) for a in alpha])[:, None]
            sa = np.array([np.sin(a) for a in alpha])[:, None]
            pp0 = j2d[..., 0] - com[idx, 0][:, None]            # float32
            pp1 = j2d[..., 1] - com[idx, 1][:, None]
            d2 = np.zeros_like(j2d)
            d2[..., 0] = pp0.astype(f64) * ca - pp1.astype(f64) * sa
            d2[..., 1] = pp0.astype(f64) * sa + pp1.astype(f64) * ca
            d2[..., 2] = j2d[..., 2]
            d2[..., 0] += com[idx, 0][:, None]
            d2[..., 1] += com[idx, 1][:, None]
            joints[idx] = (self._to3d(d2) - c3[:, None, :]).astype(f32)
        is_sc = (names == 'sc') & ~np.isclose(sc, 1.)
        if is_sc.any():
            idx = np.nonzero(is_sc)[0]
            new_cube[idx] = cube[idx].astype(f64) * sc[idx][:, None]
            homography(idx, self._transform_batch(com[idx], new_cube[idx]))
            _, _, _, _, zs, ze = self._bounds_batch(com[idx], cube[idx])      # z-threshold with the OLD cube
            rec['zstart'][idx], rec['zend'][idx] = zs.astype(f32), ze.astype(f32)
            mode[idx] = 2
        rec['mode'] = mode
        half_new = (new_cube[:, 2] / 2.).astype(f32)
        rec['bg'] = (new_com[:, 2].astype(f64) + new_cube[:, 2] / 2.).astype(f32)
        rec['lo'] = (new_com[:, 2].astype(f64) - new_cube[:, 2] / 2.).astype(f32)
        rec['comz_new'] = new_com[:, 2]
        rec['half_new'] = half_new
        return rec, (joints / half_new[:, None, None]).astype(f32)
