"""Small host utilities the entry scripts and the trainer import (reference: src/util/helpers.py:35-155).
Pure NumPy bookkeeping - nothing here is on the device path.  ``cartesian`` / ``gaussian_kernel`` / ``rgb_to_gray``
serve the reference's plotting and heat-map code paths, which are out of scope, and are provided for import
compatibility only."""
import numpy


def shuffle_many_inplace(arrays, random_state=None):
    """Apply ONE random permutation to the first axis of every array, in place (helpers.py:87-108; the MSRA15
    cross-validation script shuffles images, labels and side arrays together with it).  The permutation is the
    reference's Fisher-Yates walk - ``rng.randint(i + 1)`` for i = n-1 .. 1 - so a given RandomState produces the same
    order as the reference."""
    if random_state is None:
        rng = numpy.random.mtrand._rand
    elif isinstance(random_state, numpy.random.RandomState):
        rng = random_state
    else:
        raise ValueError("random_state must be None or numpy RandomState")
    n = arrays[0].shape[0]
    assert all(a.shape[0] == n for a in arrays[1:])
    for hi in range(n - 1, 0, -1):
        pick = rng.randint(hi + 1)
        for a in arrays:
            a[[hi, pick]] = a[[pick, hi]]


def chunks(l, n):
    """successive slices of ``l`` with at most ``n`` items (helpers.py:145-155)"""
    for start in range(0, len(l), n):
        yield l[start:start + n]


def cartesian(arrays, out=None):
    """all combinations of the given 1-D arrays, first array varying slowest (helpers.py:35-84)"""
    arrays = [numpy.asarray(a) for a in arrays]
    grids = numpy.meshgrid(*arrays, indexing='ij')
    table = numpy.stack([g.reshape(-1) for g in grids], axis=1).astype(arrays[0].dtype)
    if out is not None:
        out[...] = table
        return out
    return table


def gaussian_kernel(kernel_shape, sigma=None):
    """normalised 2-D Gaussian of edge ``kernel_shape`` (helpers.py:111-133; OpenCV's default sigma rule)"""
    if sigma is None:
        sigma = 0.3 * ((kernel_shape - 1.) * 0.5 - 1.) + 0.8
    mid = numpy.floor(kernel_shape / 2.)
    ax = numpy.arange(kernel_shape, dtype='float64') - mid
    xx, yy = numpy.meshgrid(ax, ax, indexing='ij')
    kern = (1. / (2. * numpy.pi * sigma ** 2.) * numpy.exp(-(xx ** 2. + yy ** 2.) / (2. * sigma ** 2.))).astype('float32')
    return kern / numpy.sum(kern)


def rgb_to_gray(rgb):
    assert len(rgb) == 3, "rgb should be 3, got {}".format(len(rgb))
    g = 0.21 * rgb[0] + 0.72 * rgb[1] + 0.07 * rgb[2]
    return numpy.asarray([g, g, g])
