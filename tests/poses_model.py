"""NumPy transcription of ``k_sample_poses`` (deep-prior-pp_b200/csrc/poses.cu), vectorised over (pose, joint).
TEST INFRASTRUCTURE ONLY: lets the CPU suite run the host wrapper (HandDetector.sampleRandomPoses) end to end with the
kernel emulated; written from the kernel, not from the oracle, so that the two restatements check each other."""
import numpy as np

f32, f64 = np.float32, np.float64


def _to_img(cam, s):
    fx, fy, ux, uy, flip = cam
    s = s.astype(f32)
    nz = s[..., 2] != 0
    z = np.where(nz, s[..., 2], f32(1))
    q0 = (s[..., 0] / z).astype(f64)
    q1 = (s[..., 1] / z).astype(f64)
    u = (q0 * fx + ux).astype(f32)
    v = ((uy - q1 * fy) if flip else (q1 * fy + uy)).astype(f32)
    out = np.stack([np.where(nz, u, f32(ux)), np.where(nz, v, f32(uy)), np.where(nz, s[..., 2], f32(0))], axis=-1)
    return out.astype(f32)


def _to_3d(cam, s):
    fx, fy, ux, uy, flip = cam
    s0, s1, s2 = s[..., 0].astype(f64), s[..., 1].astype(f64), s[..., 2].astype(f64)
    x = ((s0 - ux) * s2 / fx).astype(f32)
    y = (((uy - s1) * s2 / fy) if flip else ((s1 - uy) * s2 / fy)).astype(f32)
    return np.stack([x, y, s[..., 2].astype(f32)], axis=-1)


def run(base_poses, base_com, base_cube, mode, ridx, off, sc, cos_sin, cam):
    n, J = len(mode), base_poses.shape[1]
    p = base_poses[ridx].astype(f32)                       # (n, J, 3)
    com = base_com[ridx].astype(f32)[:, None, :]
    cube = base_cube[ridx].astype(f32)
    m = np.asarray(mode)[:, None, None]
    moved = (m == 3) | (m == 4) | (m == 5)
    ncom = np.where(moved, (com.astype(f64) + off[:, None, :]).astype(f32), com)
    s = sc.astype(f32)
    ncube = np.where((np.asarray(mode) == 2)[:, None], cube * s[:, None], cube).astype(f32)
    half = (ncube[:, 2].astype(f64) / 2.).astype(f32)[:, None, None]
    shifted = ((p + com) - ncom).astype(f32)
    scaled = np.where(m == 5, shifted * s[:, None, None], shifted).astype(f32)
    q = np.where(m == 1, p + ncom, scaled + com).astype(f32)
    centre = np.where(m == 1, _to_img(cam, com), _to_img(cam, ncom))      # (n, 1, 3)
    uvd = _to_img(cam, q)
    pp0 = (uvd[..., 0] - centre[..., 0]).astype(f32)
    pp1 = (uvd[..., 1] - centre[..., 1]).astype(f32)
    ca, sa = cos_sin[:, 0][:, None], cos_sin[:, 1][:, None]
    pr0 = (pp0.astype(f64) * ca - pp1.astype(f64) * sa).astype(f32)
    pr1 = (pp0.astype(f64) * sa + pp1.astype(f64) * ca).astype(f32)
    rot2d = np.stack([pr0 + centre[..., 0], pr1 + centre[..., 1], uvd[..., 2]], axis=-1).astype(f32)
    xyz = _to_3d(cam, rot2d)
    rotated = (xyz - np.where(m == 1, ncom, com)).astype(f32)
    o = np.where((m == 0) | (m == 2), p, np.where(m == 3, shifted, rotated)).astype(f32)
    return (o / half).astype(f32), ncom[:, 0, :].astype(f32), ncube
