"""Dataset - turns sequences of cropped depth frames into the network's input / label stacks (reference:
src/data/dataset.py:40-111 ``Dataset.imgStackDepthOnly``, :114-150 the per-dataset subclasses).

``imgStackDepthOnly(seqName)`` returns ``(imgStack (N,1,H,W) float32 in [-1, 1], labelStack (N,J,3) float32)``:
per frame ``dpt == 0 -> com_z + cube_z/2``, ``(dpt - com_z) / (cube_z/2)`` and ``gt3Dcrop / (cube_z/2)``
(dataset.py:97-103).  The pixel work is the same normalisation the cascade's crop kernel applies, so the stack is
produced on the device by ``dpp_recrop_fwd`` with identity windows (csrc/recrop.cu); labels are one float32 division.
File readers live in the importers and are out of scope (datasets absent): sequences come from
``data.synthetic.generate_sequence`` or from any code that fills ``NamedImgSequence`` tuples."""
import numpy

from data.importers import ICVLImporter, MSRA15Importer, NYUImporter


class Dataset(object):
    def __init__(self, imgSeqs=None, localCache=True):
        self.localCache = localCache
        self._imgSeqs = [] if imgSeqs is None else imgSeqs
        self._imgStacks = {}
        self._labelStacks = {}

    @property
    def imgSeqs(self):
        return self._imgSeqs

    @imgSeqs.setter
    def imgSeqs(self, value):
        self._imgSeqs = value
        self._imgStacks = {}

    def imgSeq(self, seqName):
        for seq in self._imgSeqs:
            if seq.name == seqName:
                return seq
        return []

    def imgStackDepthOnly(self, seqName, normZeroOne=False):
        imgSeq = None
        for seq in self._imgSeqs:
            if seq.name == seqName:
                imgSeq = seq
                break
        if imgSeq is None:
            return []
        if seqName not in self._imgStacks:
            imgStack, labelStack = self._stack(imgSeq, normZeroOne)
            if not self.localCache:
                return imgStack, labelStack
            self._imgStacks[seqName] = imgStack
            self._labelStacks[seqName] = labelStack
        return self._imgStacks[seqName], self._labelStacks[seqName]

    @staticmethod
    def _stack(imgSeq, normZeroOne):
        import torch
        from dpp_b200.cascade import run_crop_records
        from dpp_b200.lib import CROP_REC_DTYPE, CROP_NORMALISE, DppError
        if normZeroOne:
            raise NotImplementedError("normZeroOne stacks ([0, 1] range) are not used by the entry scripts")
        if not torch.cuda.is_available():
            raise DppError("Dataset.imgStackDepthOnly normalises on the device; there is no CPU fallback")
        f32, f64 = numpy.float32, numpy.float64
        n = len(imgSeq.data)
        frames = numpy.stack([numpy.asarray(fr.dpt, 'float32') for fr in imgSeq.data])
        h, w = frames.shape[1:]
        cube_z = imgSeq.config['cube'][2]
        comz = numpy.array([fr.com[2] for fr in imgSeq.data])
        rec = numpy.zeros(n, dtype=CROP_REC_DTYPE)           # identity window over every stored crop
        rec['src_index'] = numpy.arange(n)
        rec['wb'], rec['hb'], rec['rw'], rec['rh'] = w, h, w, h
        rec['flags'] = CROP_NORMALISE
        rec['zstart'], rec['zend'] = -numpy.inf, numpy.inf
        rec['hi'] = (comz.astype(f64) + cube_z / 2.).astype(f32)      # com[2] + (cube[2] / 2.), stored float32
        rec['comz'] = comz.astype(f32)
        rec['half'] = f32(cube_z / 2.)
        rec['ifx'] = rec['ify'] = 1.
        dev = torch.from_numpy(frames).cuda()
        out = torch.empty((n, h, w), dtype=torch.float32, device=dev.device)
        run_crop_records(dev, rec, out)
        imgStack = out.cpu().numpy().reshape(n, 1, h, w)
        labelStack = numpy.stack([numpy.asarray(fr.gt3Dcrop, dtype='float32') for fr in imgSeq.data]) / f32(cube_z / 2.)
        return imgStack, labelStack.astype('float32')


class ICVLDataset(Dataset):
    def __init__(self, imgSeqs=None, basepath=None, localCache=True):
        super(ICVLDataset, self).__init__(imgSeqs, localCache)
        self.lmi = ICVLImporter('../../data/ICVL/' if basepath is None else basepath)


class MSRA15Dataset(Dataset):
    def __init__(self, imgSeqs=None, basepath=None, localCache=True):
        super(MSRA15Dataset, self).__init__(imgSeqs, localCache)
        self.lmi = MSRA15Importer('../../data/MSRA15/' if basepath is None else basepath)


class NYUDataset(Dataset):
    def __init__(self, imgSeqs=None, basepath=None, localCache=True):
        super(NYUDataset, self).__init__(imgSeqs, localCache)
        self.lmi = NYUImporter('../../data/NYU/' if basepath is None else basepath)
