"""Small point transforms (reference: src/data/transformations.py:58-88)."""
import numpy


def transformPoint2D(pt, M):
    pt2 = numpy.dot(numpy.asarray(M).reshape((3, 3)), numpy.asarray([pt[0], pt[1], 1]))
    return numpy.asarray([pt2[0] / pt2[2], pt2[1] / pt2[2]])


def transformPoints2D(pts, M):
    ret = pts.copy()
    for i in range(pts.shape[0]):
        ret[i, 0:2] = transformPoint2D(pts[i, 0:2], M)
    return ret


def rotatePoint2D(p1, center, angle):
    """Rotate (u,v,d) about ``center`` by ``angle`` degrees (transformations.py:71-88): the
    subtraction/addition of the centre happen in the point's own dtype (float32 in the caller),
    the rotation itself in float64, stored back to the point's dtype."""
    alpha = numpy.float64(angle) * numpy.pi / 180.
    pp = p1.copy()
    pp[0:2] -= center[0:2]
    pr = numpy.zeros_like(pp)
    pr[0] = numpy.float64(pp[0]) * numpy.cos(alpha) - numpy.float64(pp[1]) * numpy.sin(alpha)
    pr[1] = numpy.float64(pp[0]) * numpy.sin(alpha) + numpy.float64(pp[1]) * numpy.cos(alpha)
    pr[2] = pp[2]
    ps = pr
    ps[0:2] += center[0:2]
    return ps


def rotatePoints2D(pts, center, angle):
    """Rotate every (u,v,d) row of ``pts`` about ``center`` by ``angle`` degrees (transformations.py:91-103)."""
    ret = pts.copy()
    for i in range(pts.shape[0]):
        ret[i] = rotatePoint2D(pts[i], center, angle)
    return ret
