"""Camera models of the dataset importers (reference: src/data/importers.py: DepthImporter
:51-150, ICVLImporter :186-210, MSRA15Importer :536-570 + :756-793, NYUImporter :880-920 +
:1187-1224).  Only the pin-hole projections and per-dataset constants are on the hot path
(augmentation label math); the file readers are out of scope (datasets are absent):
``loadSequence`` keeps its signature; with ``basepath=None`` - or with a basepath and the explicit
opt-in DPP_SYNTHETIC=1 (a loud warning names the directory that is NOT read) - it returns a
deterministic synthetic ``NamedImgSequence`` (``data.synthetic``) so that the entry scripts' data
preparation runs unchanged; with a basepath and no opt-in it raises.

dtype discipline: the reference ran on NumPy 1.x value-based casting, where
``np.float32 scalar (op) python float`` is computed in float64 and only the store into the
``np.float32`` result array rounds.  NumPy 2 would keep float32, so every expression below casts
explicitly."""
import numpy as np

f64 = np.float64


class DepthImporter(object):
    def __init__(self, fx, fy, ux, uy, hand=None):
        self.fx = fx
        self.fy = fy
        self.ux = ux
        self.uy = uy
        self.depth_map_size = (320, 240)
        self.refineNet = None
        self.crop_joint_idx = 0
        self.hand = hand
        self.flip_y = False

    def jointsImgTo3D(self, sample):
        ret = np.zeros((sample.shape[0], 3), np.float32)
        for i in range(sample.shape[0]):
            ret[i] = self.jointImgTo3D(sample[i])
        return ret

    def jointImgTo3D(self, sample):
        ret = np.zeros((3,), np.float32)
        ret[0] = (f64(sample[0]) - self.ux) * f64(sample[2]) / self.fx
        if self.flip_y:
            ret[1] = (self.uy - f64(sample[1])) * f64(sample[2]) / self.fy
        else:
            ret[1] = (f64(sample[1]) - self.uy) * f64(sample[2]) / self.fy
        ret[2] = sample[2]
        return ret

    def joints3DToImg(self, sample):
        ret = np.zeros((sample.shape[0], 3), np.float32)
        for i in range(sample.shape[0]):
            ret[i] = self.joint3DToImg(sample[i])
        return ret

    def joint3DToImg(self, sample):
        ret = np.zeros((3,), np.float32)
        if sample[2] == 0.:
            ret[0] = self.ux
            ret[1] = self.uy
            return ret
        sample = np.asarray(sample)
        if sample.dtype == np.float32:          # f32/f32 stays f32 before meeting the python floats
            q0 = f64(np.float32(sample[0]) / np.float32(sample[2]))
            q1 = f64(np.float32(sample[1]) / np.float32(sample[2]))
        else:
            q0 = f64(sample[0]) / f64(sample[2])
            q1 = f64(sample[1]) / f64(sample[2])
        ret[0] = q0 * self.fx + self.ux
        ret[1] = (self.uy - q1 * self.fy) if self.flip_y else (q1 * self.fy + self.uy)
        ret[2] = sample[2]
        return ret

    # -- sequences ----------------------------------------------------------------------------------------------
    _dataset_name = None            # key of data.synthetic.DATASETS; set by the dataset importers

    def _synthetic_sequence(self, seqName, Nmax, shuffle, rng, cube):
        """``loadSequence`` without dataset files.  The readers of the reference (importers.py:233-372, :597-735,
        :943-1087: PNG / binary depth maps, annotation files, hand detection, pickle cache) are out of scope and the
        datasets are absent; what the rest of the pipeline consumes is the returned ``NamedImgSequence`` of
        ``DepthFrame``s, which ``data.synthetic.generate_sequence`` builds with the same fields, dtypes and value
        ranges.  The sequence is deterministic per (dataset, sequence name); count: ``Nmax`` or DPP_SYNTH_FRAMES."""
        import os
        import zlib
        from data import synthetic
        from data.basetypes import NamedImgSequence
        # A caller that names a dataset directory expects ITS frames.  The readers are not part of this package, so
        # never hand back synthetic frames silently: with a basepath the caller must opt in (DPP_SYNTHETIC=1), and
        # even then every call says what it returns.  ``basepath=None`` (data.synthetic, tests) is the explicit
        # synthetic importer and stays quiet.
        if self.basepath is not None:
            if os.environ.get('DPP_SYNTHETIC', '0') != '1':
                raise NotImplementedError(
                    "%s.loadSequence(%r): the dataset file readers of the reference (src/data/importers.py) are not part "
                    "of this package - basepath=%r is NOT read.  Set DPP_SYNTHETIC=1 to run on deterministic synthetic "
                    "frames instead (results then say nothing about the real dataset), or construct the importer with "
                    "basepath=None." % (type(self).__name__, seqName, self.basepath))
            print("WARNING: %s.loadSequence(%r) returns SYNTHETIC frames (DPP_SYNTHETIC=1); basepath %r is not read%s"
                  % (type(self).__name__, seqName, self.basepath,
                     " although it exists" if os.path.exists(self.basepath) else ""))
        if cube is None:
            cube = self.default_cubes[seqName]
        else:
            assert isinstance(cube, tuple)
            assert len(cube) == 3
        n = int(os.environ.get('DPP_SYNTH_FRAMES', '256'))
        if Nmax != float('inf'):
            n = min(n, int(Nmax))
        seed = zlib.crc32(('%s/%s' % (self._dataset_name, seqName)).encode()) % (2 ** 31)
        seq = synthetic.generate_sequence(self._dataset_name, n, seed=seed, cube=cube, seq_name=seqName)
        data = list(seq.data)
        if shuffle and rng is not None:
            print("Shuffling")
            rng.shuffle(data)                           # importers.py:368-370
        return NamedImgSequence(seqName, data, {'cube': cube})

    def getCameraProjection(self):
        ret = np.zeros((4, 4), np.float32)
        ret[0, 0] = self.fx
        ret[1, 1] = -self.fy if self.flip_y else self.fy
        ret[2, 2] = 1.
        ret[0, 2] = self.ux
        ret[1, 2] = self.uy
        ret[3, 2] = 1.
        return ret


class ICVLImporter(DepthImporter):
    def __init__(self, basepath=None, useCache=True, cacheDir='./cache/', refineNet=None, hand=None):
        super(ICVLImporter, self).__init__(241.42, 241.42, 160., 120., hand)
        self.depth_map_size = (320, 240)
        self.basepath = basepath
        self.numJoints = 16
        self.crop_joint_idx = 0
        self.refineNet = refineNet
        self.default_cubes = {'train': (250, 250, 250), 'test_seq_1': (250, 250, 250), 'test_seq_2': (250, 250, 250)}
        self.sides = {'train': 'right', 'test_seq1': 'right', 'test_seq_2': 'right'}       # (sic) importers.py:211
        self._dataset_name = 'ICVL'

    def loadSequence(self, seqName, subSeq=None, Nmax=float('inf'), shuffle=False, rng=None, docom=False, cube=None):
        if (subSeq is not None) and (not isinstance(subSeq, list)):
            raise TypeError("subSeq must be None or list")
        return self._synthetic_sequence(seqName, Nmax, shuffle, rng, cube)


class MSRA15Importer(DepthImporter):
    def __init__(self, basepath=None, useCache=True, cacheDir='./cache/', refineNet=None, detectorNet=None,
                 derotNet=None, hand=None):
        super(MSRA15Importer, self).__init__(241.42, 241.42, 160., 120., hand)
        self.flip_y = True
        self.depth_map_size = (320, 240)
        self.basepath = basepath
        self.refineNet = refineNet
        self.numJoints = 21
        self.crop_joint_idx = 5
        self.default_cubes = {'P0': (200, 200, 200), 'P1': (200, 200, 200), 'P2': (200, 200, 200),
                              'P3': (180, 180, 180), 'P4': (180, 180, 180), 'P5': (180, 180, 180),
                              'P6': (170, 170, 170), 'P7': (160, 160, 160), 'P8': (150, 150, 150)}
        self.sides = dict((k, 'right') for k in self.default_cubes)
        self._dataset_name = 'MSRA15'

    def loadSequence(self, seqName, subSeq=None, Nmax=float('inf'), shuffle=False, rng=None, docom=False, cube=None):
        if (subSeq is not None) and (not isinstance(subSeq, list)):
            raise TypeError("subSeq must be None or list")
        return self._synthetic_sequence(seqName, Nmax, shuffle, rng, cube)


class NYUImporter(DepthImporter):
    def __init__(self, basepath=None, useCache=True, cacheDir='./cache/', refineNet=None, allJoints=False, hand=None):
        super(NYUImporter, self).__init__(588.03, 587.07, 320., 240., hand)
        self.flip_y = True
        self.depth_map_size = (640, 480)
        self.basepath = basepath
        self.allJoints = allJoints
        self.numJoints = 36
        self.crop_joint_idx = 32 if allJoints else 13
        self.default_cubes = {'train': (300, 300, 300), 'test_1': (300, 300, 300), 'test_2': (250, 250, 250),
                              'test': (300, 300, 300), 'train_synth': (300, 300, 300),
                              'test_synth_1': (300, 300, 300), 'test_synth_2': (250, 250, 250),
                              'test_synth': (300, 300, 300)}
        self.sides = dict((k, 'right') for k in self.default_cubes)
        self.restrictedJointsEval = [0, 3, 6, 9, 12, 15, 18, 21, 24, 25, 27, 30, 31, 32]
        self.refineNet = refineNet
        self._dataset_name = 'NYU'

    def loadSequence(self, seqName, Nmax=float('inf'), shuffle=False, rng=None, docom=False, cube=None):
        return self._synthetic_sequence(seqName, Nmax, shuffle, rng, cube)
