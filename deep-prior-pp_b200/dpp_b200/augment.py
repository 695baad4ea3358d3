"""Host wrapper of dpp_augment_fwd: upload crops + records, launch, download (utility path used
by the reference-signature HandDetector methods and by tests; the trainer keeps everything on the
device)."""
import ctypes as C
import numpy as np

from .lib import lib, AUG_REC_DTYPE, DppError


def run_records_device(crops_dev, recs_np, out_dev=None):
    """crops_dev: torch CUDA (N,H,W) f32; recs_np: structured array (n,) AUG_REC_DTYPE.
    Returns torch CUDA (n,H,W)."""
    import torch
    recs_np = np.ascontiguousarray(recs_np, dtype=AUG_REC_DTYPE)
    n = recs_np.shape[0]
    H, W = int(crops_dev.shape[-2]), int(crops_dev.shape[-1])
    if n == 0:
        return torch.empty((0, H, W), dtype=torch.float32, device=crops_dev.device)
    rec_dev = torch.from_numpy(recs_np.view(np.uint8).reshape(n, -1).copy()).to(crops_dev.device)
    if out_dev is None:
        out_dev = torch.empty((n, H, W), dtype=torch.float32, device=crops_dev.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.dpp_augment_fwd(C.c_void_p(crops_dev.data_ptr()), C.c_void_p(rec_dev.data_ptr()),
                        C.c_void_p(out_dev.data_ptr()), n, H, W, st)
    return out_dev


def run_records(crops_np, recs_np):
    import torch
    if not torch.cuda.is_available():
        raise DppError("dpp_augment_fwd needs a CUDA device; there is no CPU fallback")
    crops = torch.from_numpy(np.ascontiguousarray(crops_np, np.float32)).cuda()
    return run_records_device(crops, recs_np).cpu().numpy()
