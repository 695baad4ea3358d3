#!/usr/bin/env python
"""BASELINE config 5: CoM-refine + posereg inference cascade, batch 1024, 1 x B200 - p50 latency + frames/s.

One "step" = one batch of synthetic NYU depth frames (640x480, integer mm) through
  dpp_recrop_fwd (window -> 128x128 + centre crops, normalised) -> ScaleNet forward -> host: refined CoMs ->
  dpp_recrop_fwd (aspect-preserving crop, normalised) -> ResNet type 1 (30-D bottleneck + prior layer) forward ->
  poses (B, 14, 3) in mm on the host
(reference: src/util/realtimehandposepipeline.py:296-370, src/test_realtimepipeline.py:61-67; one frame at a time
there).  Prints ONE JSON line in bench.py's format:
  value : frames/s with the frames resident in HBM (host geometry + the two small device->host reads included -
          they are part of the path);  p50_ms / p90_ms : latency of one batch
  e2e   : frames in pinned host memory, copied to the device inside the timed region, poses read back
  roofline : k_recrop (HBM-bound), algorithmic bytes = 4 B per sampled pixel read + 4 B per pixel written
  cpu_baseline : the oracle cascade (cv2 + torch-CPU nets, one frame at a time like the reference's loop) on a
          bounded sample of the same frames.
Not the driver's headline (that is bench.py's training step); `python bench.py --workload cascade` forwards here."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'deep-prior-pp_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "cascade inference frames/sec (CoM-refine ScaleNet + re-crop + ResNet posereg), 640x480 depth frames"
FX, FY = 588., 587.          # src/test_realtimepipeline.py:66
J = 14


def make_frames(batch, unique=64, seed=23455):
    from data import synthetic
    fr = synthetic.generate_frames('NYU', min(unique, batch), seed=seed, edge_fraction=0.25)
    reps = (batch + fr['frames'].shape[0] - 1) // fr['frames'].shape[0]
    rng = np.random.RandomState(seed + 1)
    frames = np.tile(fr['frames'], (reps, 1, 1))[:batch]
    lastcom = np.tile(fr['lastcom'], (reps, 1))[:batch]
    lastcom = lastcom + rng.uniform(-2, 2, lastcom.shape) * np.array([1., 1., 2.])     # distinct windows per copy
    return fr, np.ascontiguousarray(frames), lastcom


def cpu_baseline(fr, frames, lastcom, n_frames, threads):
    import torch
    from oracle import cascade as OC, augment as OA, nets as ON
    torch.set_num_threads(threads)
    opose = ON.build_resnet(np.random.RandomState(23455), type=1, batchSize=1, numJoints=J, nDims=3)
    oref = ON.build_scalenet(np.random.RandomState(23455), type=1, batchSize=1, numJoints=1, nDims=3)
    cam = OA.Camera(**OA.NYU_CAM)

    def refine_fn(xs):
        with torch.no_grad():
            return oref.forward([torch.from_numpy(x) for x in xs], deterministic=True)[0].numpy()

    def pose_fn(x):
        with torch.no_grad():
            return opose.forward(torch.from_numpy(x), deterministic=True)[0].numpy()
    OC.cascade_frame(frames[0], lastcom[0], fr['cube'], cam, FX, FY, refine_fn, pose_fn, use_cv2=True)   # warm-up
    t0 = time.time()
    for i in range(n_frames):
        OC.cascade_frame(frames[i], lastcom[i], fr['cube'], cam, FX, FY, refine_fn, pose_fn, use_cv2=True)
    return n_frames / (time.time() - t0)


def recrop_roofline(torch, casc, frames_dev, lastcom, cube, di, hbm_peak, how):
    """k_recrop timed alone with CUDA events on the launching stream: the pose-net crop of the whole batch,
    repeated over DIFFERENT output buffers / frames larger than L2 (1024 frames = 1.26 GB)."""
    from dpp_b200 import cascade as PC
    from dpp_b200.lib import lib
    n = len(lastcom)
    rec, _, _ = PC.pose_records(lastcom, cube, FX, FY, di, frames_dev.shape[1:], 0.)
    rec_dev = torch.from_numpy(rec.view(np.uint8).reshape(n, -1).copy()).to(frames_dev.device)
    out = casc.pose_eng.t_ins[0].buf
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    Hf, Wf = int(frames_dev.shape[1]), int(frames_dev.shape[2])

    def launch():
        lib.dpp_recrop_fwd(C.c_void_p(frames_dev.data_ptr()), C.c_void_p(rec_dev.data_ptr()),
                           C.c_void_p(out.data_ptr()), None, None, n, Hf, Wf, 128, 128, st)
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sampled = float((rec['rw'].astype(np.int64) * rec['rh']).sum())      # pixels gathered from the frames
    alg = 4.0 * sampled + 4.0 * n * 128 * 128
    gbs = alg / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "k_recrop: pose-net crop of %d frames, one launch" % n, "achieved": gbs,
            "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
            "peak_source": "%s HBM copy bandwidth (MEASURED_PEAKS.json)" % how, "ms_per_launch": ms,
            "launches_timed": reps, "alg_bytes_per_launch": alg,
            "note": "a sampled output row touches a contiguous span of ~wb*4 bytes of a frame row, so DRAM traffic is "
                    "up to wb/128 x the algorithmic read bytes; L2 sees only the first use of each frame"}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-frames', type=int, default=24)
    args, _ = ap.parse_known_args(argv)
    import torch
    from dpp_b200 import cascade as PC
    from net.resnet import ResNet, ResNetParams
    from net.scalenet import ScaleNet, ScaleNetParams
    sys.path.insert(0, ROOT)
    from bench import ClockSampler, peaks, cpu_threads
    B = args.batch
    torch.cuda.set_device(0)
    fr, frames, lastcom = make_frames(B)
    di, cube = fr['importer'], fr['cube']
    pose = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=1, nChan=1, wIn=128, hIn=128, batchSize=B,
                                                                      numJoints=J, nDims=3))
    ref = ScaleNet(np.random.RandomState(23455), cfgParams=ScaleNetParams(type=1, nChan=1, wIn=128, hIn=128, batchSize=B,
                                                                         resizeFactor=2, numJoints=1, nDims=3))
    casc = PC.Cascade(pose, ref, di, FX, FY, cube)
    frames_dev = torch.from_numpy(frames).cuda()
    frames_pin = torch.from_numpy(frames).pin_memory()
    stage = torch.empty_like(frames_dev)

    def step_resident():
        return casc.run(frames_dev, lastcom, ndvalue=0.)

    def step_e2e():
        stage.copy_(frames_pin, non_blocking=True)
        return casc.run(stage, lastcom, ndvalue=0.)

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        lat = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            t0 = time.perf_counter()
            fn()                                  # ends with the device->host read of the poses: synchronous
            lat.append((time.perf_counter() - t0) * 1e3)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, lat      # device clock over the K batches, host gaps included

    clocks = ClockSampler(0)
    clocks.start()
    ms, lat = timed(step_resident)
    ms_e2e, lat_e2e = timed(step_e2e)
    clk = clocks.stop()
    burst, sustained, hbm, how = peaks()
    roof = recrop_roofline(torch, casc, frames_dev, lastcom, cube, di, hbm, how)
    n_launch = 2                                  # the two crop launches
    for eng in (casc.ref_eng, casc.pose_eng):     # forward launches of OUR library per net
        for op in eng.ops:
            n_launch += {'conv': 1, 'convpool': 1, 'fc': 2, 'bn_apply': 1, 'concat': len(op.get('srcs', []))}.get(op['kind'], 1)
    cpu = None
    if not args.no_cpu_baseline:
        th = cpu_threads()
        nf = min(args.cpu_frames, B)
        v = cpu_baseline(fr, frames, lastcom, nf, th)
        cpu = {"value": v, "unit": "frames/s", "cores": th, "kind": "port",
               "sample": "%d frames of the same batch, one at a time (cv2 crops + torch-CPU ScaleNet / ResNet, the "
                         "reference's per-frame loop) after 1 warm-up frame" % nf}
    h2d = int(frames.nbytes) + 2 * B * 88
    out = {
        "metric": METRIC, "value": B / (ms / 1e3), "unit": "frames/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "p50_ms": float(np.percentile(lat, 50)),
        "p90_ms": float(np.percentile(lat, 90)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32x3", "data": "synthetic",
        "config": {"workload": "CoM-refine (ScaleNet type 1) + posereg (ResNet type 1, 14 joints) inference cascade, "
                               "batch %d synthetic NYU 640x480 frames" % B, "global_batch": B,
                   "l2": "the batch of frames is %.2f GB (L2: 126 MB)" % (frames.nbytes / 1e9),
                   "timing": "CUDA events around the K batches (host geometry between the two stages and the two "
                             "device->host reads are part of the path); p50/p90 from the host clock around each batch, "
                             "which ends with a synchronous read of the poses"},
        "e2e": {"value": B / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e,
                "p50_ms": float(np.percentile(lat_e2e, 50)), "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": B * 3 * 4 + B * J * 3 * 4},
        "gpu_launches": n_launch * args.steps, "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(out))


if __name__ == '__main__':
    main()
