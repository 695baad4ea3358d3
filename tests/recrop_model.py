"""NumPy model of ``k_recrop`` (deep-prior-pp_b200/csrc/recrop.cu): interprets ``dpp_crop_rec`` records exactly as the
kernel does.  TEST INFRASTRUCTURE ONLY - it lets the CPU suite check the host-side record preparation
(dpp_b200/cascade.py) against the oracle without a GPU; the GPU suite checks the kernel itself."""
import numpy as np

f32 = np.float32


def _axis_table(n_out, p, n_dst, n_src, s0, inv, lim):
    d = np.arange(n_out) - p
    s = np.full(n_out, -1, np.int64)
    ok = (d >= 0) & (d < n_dst)
    idx = np.minimum(np.floor(d[ok].astype(np.float64) * inv).astype(np.int64), n_src - 1) + s0
    idx[(idx < 0) | (idx >= lim)] = -2
    s[ok] = idx
    return s


def run(frames, recs, H=128, W=128, centre_crops=False):
    n = len(recs)
    out0 = np.zeros((n, H, W), f32)
    for i, r in enumerate(recs):
        fr = frames[int(r['src_index'])]
        Hf, Wf = fr.shape
        sx = _axis_table(W, int(r['px']), int(r['rw']), int(r['wb']), int(r['xstart']), float(r['ifx']), Wf)
        sy = _axis_table(H, int(r['py']), int(r['rh']), int(r['hb']), int(r['ystart']), float(r['ify']), Hf)
        SY, SX = np.meshgrid(sy, sx, indexing='ij')
        fill = (SX == -1) | (SY == -1)
        inside = (SX >= 0) & (SY >= 0)
        v = np.zeros((H, W), f32)
        v[inside] = fr[SY[inside], SX[inside]]
        m1 = (v < r['zstart']) & (v != 0) & ~fill
        m2 = (v > r['zend']) & (v != 0) & ~fill
        v[m1] = r['zstart']
        v[m2] = 0.
        v[fill] = r['fill']
        flags = int(r['flags'])
        if flags & 1:
            v[v == 0] = r['hi']
            if flags & 2:
                v[v >= r['hi']] = r['hi']
                v[v <= r['lo']] = r['lo']
            v = ((v - f32(r['comz'])) / f32(r['half'])).astype(f32)
        if flags & 4:
            v = v[:, ::-1]
        out0[i] = v
    if not centre_crops:
        return out0
    y1, x1, y2, x2 = (H - H // 2) // 2, (W - W // 2) // 2, (H - H // 4) // 2, (W - W // 4) // 2
    return (out0, np.ascontiguousarray(out0[:, y1:y1 + H // 2, x1:x1 + W // 2]),
            np.ascontiguousarray(out0[:, y2:y2 + H // 4, x2:x2 + W // 4]))
