"""NumPy transcription of ``k_augment`` (deep-prior-pp_b200/csrc/augment.cu): interprets ``dpp_aug_rec`` records the way
the kernel does, with the oracle's index models for the two warps.  TEST INFRASTRUCTURE ONLY (used by
tests/cascade_dryrun.py to run the trainer's epoch pipeline on the CPU; the kernel itself is checked on the GPU by
tests/test_gpu_augment.py)."""
import numpy as np

from oracle import augment as OA

f32 = np.float32


def run(crops, recs, H=128, W=128):
    out = np.zeros((len(recs), H, W), f32)
    for i, r in enumerate(recs):
        src = (crops[int(r['src_index'])].astype(f32) * f32(r['half_old']) + f32(r['comz_old'])).astype(f32)
        premax = src.max()
        mode = int(r['mode']) & 15
        m = np.asarray(r['m'], np.float64)
        if mode == 0:
            v = src.copy()
        elif mode == 1:
            Y, X, inside = OA.warp_affine_nn_indices(m[:6], W, H)
            v = np.zeros((H, W), f32)
            v[inside] = src[Y[inside], X[inside]]
        else:
            Y, X, inside = OA.warp_perspective_nn_indices(m, W, H)
            v = np.zeros((H, W), f32)
            v[inside] = src[Y[inside], X[inside]]
            v[np.abs(v - f32(32000.)) <= f32(0.32000001)] = 0.
            m1 = (v < r['zstart']) & (v != 0)
            m2 = (v > r['zend']) & (v != 0)
            v[m1] = r['zstart']
            v[m2] = 0.
        if not (int(r['mode']) & 16):
            v[v == premax] = r['bg']
            v[v == 0] = r['bg']
            v[v >= r['bg']] = r['bg']
            v[v <= r['lo']] = r['lo']
            v = ((v - f32(r['comz_new'])) / f32(r['half_new'])).astype(f32)
        out[i] = v
    return out
