#!/usr/bin/env python
"""bench.py - the driver's measurement contract for the DeepPrior++ hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic NYU crops:
  dpp_augment_fwd (rot/com/none, from the HBM-resident ORIGINAL crops) -> ResNet (type 0,
  nDims 30) forward -> cost -> backward -> [NCCL all-reduce] -> ADAM -> BN running-stat EMA.
Workload = BASELINE.json configs[1]: "NYU posereg_embedding ResNet training, batch 128 synthetic
depth, 1xB200".  For N > 1 every rank runs the same per-GPU batch (weak scaling, global batch
128*N) with one gradient all-reduce per step.

  value : whole-job frames/s with all inputs resident in HBM when the timed region starts
  e2e   : the same metric through the reference-facing API with HOST buffers: per step the
          batch's crops + augmentation records + labels are copied from pinned host memory and
          the cost is read back (what trainer.train_model() returns to the caller)
  --impl reference : the CPU oracle restatement (torch-CPU + cv2) of the same step on the host
          cores (the reference's own Theano path cannot run here: no Theano/Python 2).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'deep-prior-pp_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "training frames/sec, 128x128 depth crops (augment + ResNet fwd/bwd + ADAM)"
UNIT = "frames/s"
TRAIN_GFLOP_PER_FRAME = 0.7229     # BASELINE.md section 2 (fwd + dgrad + wgrad, no stem dgrad)
B = 128
E = 30
N_RESIDENT = 2048                  # crops resident in HBM: 2048 * 64 KiB = 134 MB > 126 MB L2
AUG_MODES = ['com', 'rot', 'none']


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops', 1590.0), d.get('bf16_tflops_sustained', 1400.0), d.get('hbm_gbs', 6650.0), 'measured'
    return 1590.0, 1400.0, 6650.0, 'fallback'


class ClockSampler(object):
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def shutdown_distributed(dist, eng=None):
    """Leave a multi-rank run promptly.  Destroying the NCCL communicator while captured graphs that contain its
    collectives are alive can block for minutes at interpreter exit (seen on 2 x B200: JSON printed at once, the
    ranks exited only when the launcher's timeout fired).  Drop the graphs first, synchronise, and arm a watchdog
    that ends the process if the tear-down still stalls - the measurement is complete and printed by then."""
    import torch
    sys.stdout.flush()
    timer = threading.Timer(60.0, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    try:
        if eng is not None:
            eng._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    timer.cancel()


def make_workload(seed=23455):
    """synthetic NYU set + PCA stand-in + per-sample augmentation records for many steps"""
    from data import synthetic
    ds = synthetic.generate('NYU', N_RESIDENT, seed=seed)
    comp, mean = synthetic.random_orthonormal_pca(E, ds['gt3D'].shape[1] * 3, seed=1)
    return ds, comp, mean


def records_for(ds, comp, mean, idxs, rng):
    """host side of one batch: draws in the reference order (nettrainer.py:954-957), dpp_aug_rec records and
    embedded labels in one vectorised pass (HandDetector.aug_records_batch)"""
    hd = ds['hd']
    draws = []
    for _ in idxs:
        mode = rng.randint(0, len(AUG_MODES)); off = rng.randn(3) * 5.; rot = rng.uniform(-180., 180.)
        sc = abs(1. + rng.randn() * 0.02)
        draws.append((mode, off, rot, sc))
    idxs = np.asarray(idxs)
    com = hd._toimg(ds['com3D'][idxs])
    recs, labs = hd.aug_records_batch(idxs, [AUG_MODES[d[0]] for d in draws], np.array([d[1] for d in draws]),
                                      np.array([d[2] for d in draws]), np.array([d[3] for d in draws]), com,
                                      ds['cube'][idxs], ds['M'][idxs], ds['gt3Dcrop'][idxs])
    ys = np.dot(labs.reshape(len(idxs), -1).astype(np.float64) - mean, comp.T)
    return recs, np.asarray(ys, np.float32), draws


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle restatement (reference-equivalent CPU path)
# ---------------------------------------------------------------------------------------------
def cpu_threads():
    # torch-CPU convolutions on 8x8..64x64 maps stop scaling (and regress) beyond ~16 threads
    return min(os.cpu_count() or 1, int(os.environ.get('DPP_CPU_THREADS', '16')))


def cpu_step_time(ds, comp, mean, steps, threads):
    import torch
    from oracle import nets as ON, augment as OA
    torch.set_num_threads(threads)
    ON.set_dtype(torch.float32)          # the timed CPU arm runs in the reference's precision
    onet = ON.build_resnet(np.random.RandomState(23455), type=0, batchSize=B, numJoints=1, nDims=E)
    adam = ON.Adam(onet.params)
    cam = OA.Camera(**OA.NYU_CAM)
    ohd = OA.Hand(cam, use_cv2=True)
    rng = np.random.RandomState(99)
    times = []
    for s in range(steps):
        idxs = rng.randint(0, N_RESIDENT, B)
        draws = [OA.draw_aug_params(rng, len(AUG_MODES)) for _ in idxs]
        t0 = time.time()
        x, y = OA.augment_poses(ds['x'], ds['com3D'], ds['cube'], ds['M'], ds['gt3Dcrop'], list(idxs), draws,
                                AUG_MODES, cam, ohd, pca_mean=mean, pca_components=comp)
        ON.train_step(onet, adam, torch.from_numpy(x), torch.from_numpy(y), 1e-4, 1, E)
        times.append(time.time() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = cpu_threads()
    ds, comp, mean = make_workload()
    steps = max(1, min(args.steps, 6))
    warm = 1 if args.warmup > 0 else 0
    times = cpu_step_time(ds, comp, mean, steps + warm, threads)[warm:]
    ms = 1000. * float(np.mean(times))
    val = B / (ms / 1000.)
    sample = "%d step(s) of the batch-128 workload (cv2 augmentation + torch-CPU ResNet fwd/bwd/ADAM), %d thread(s)" % (
        len(times), threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "NYU posereg_embedding ResNet(type 0, 30-D embedding) training, batch 128, "
                               "aug rot/com/none", "global_batch": B, "note": "CPU oracle restatement of the "
                   "reference path (Theano cannot run here); host cores only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import ctypes as C
    from dpp_b200.lib import lib, AUG_REC_DTYPE
    from dpp_b200.engine import Engine
    from net.resnet import ResNet, ResNetParams
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    precision = int(os.environ.get('DPP_PRECISION', '1'))
    ds, comp, mean = make_workload(seed=23455 + rank)
    net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=B, numJoints=1, nDims=E))
    eng = Engine(net, precision=precision)
    net._eng = eng
    eng._alloc_training()
    if world > 1:
        eng.set_world(world, lambda g: dist.all_reduce(g))
    total = args.warmup + args.steps
    rng = np.random.RandomState(1234 + rank)
    nrec = min(total, 64)                    # distinct record sets, cycled
    recs_all, ys_all = [], []
    for s in range(nrec):
        idxs = rng.randint(0, N_RESIDENT, B)
        r, y, _ = records_for(ds, comp, mean, idxs, rng)
        recs_all.append(r); ys_all.append(y)
    crops_dev = torch.from_numpy(ds['x'][:, 0].copy()).to(dev)
    recs_np = np.concatenate(recs_all)
    recs_dev = torch.from_numpy(recs_np.view(np.uint8).reshape(len(recs_np), -1).copy()).to(dev)
    ys_dev = torch.from_numpy(np.concatenate(ys_all)).to(dev)
    rec_bytes = np.dtype(AUG_REC_DTYPE).itemsize
    eng.set_lr(1e-4)
    launches = {'n': 0}

    def step_resident(s):
        k = s % nrec
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        lib.dpp_augment_fwd(C.c_void_p(crops_dev.data_ptr()), C.c_void_p(recs_dev.data_ptr() + k * B * rec_bytes),
                            C.c_void_p(eng.t_in.buf.data_ptr()), B, 128, 128, st)
        eng.y_in.copy_(ys_dev[k * B:(k + 1) * B], non_blocking=True)
        eng.train_step(None, use_graph=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for s in range(warmup):
            fn(s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(warmup + s)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms / steps

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_step = timed(step_resident, args.steps, max(args.warmup, 3))
    value = B * world / (ms_step / 1000.)

    # ---- e2e: host buffers in, cost out, every step
    # two staging sets: the host->device copies of batch s+1 run on a copy stream underneath the compute of batch s
    stage_x = [torch.empty((B, 128, 128), dtype=torch.float32, device=dev) for _ in range(2)]
    stage_r = [torch.empty((B, rec_bytes), dtype=torch.uint8, device=dev) for _ in range(2)]
    stage_y = [torch.empty((B, E), dtype=torch.float32, device=dev) for _ in range(2)]
    host_batches = []
    for k in range(nrec):
        r = recs_all[k].copy()
        src = r['src_index'].copy()
        r['src_index'] = np.arange(B, dtype=np.int32)        # records index the staged batch
        host_batches.append((torch.from_numpy(ds['x'][src, 0].copy()).pin_memory(),
                             torch.from_numpy(r.view(np.uint8).reshape(B, -1).copy()).pin_memory(),
                             torch.from_numpy(ys_all[k].copy()).pin_memory()))
    cost_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]        # staging set b holds a complete batch
    consumed = [torch.cuda.Event(), torch.cuda.Event()]      # the step that read staging set b has consumed it
    state = {'next': None}

    def upload(s):
        b = s & 1
        hx, hr, hy = host_batches[s % nrec]
        copy_stream.wait_event(consumed[b])
        with torch.cuda.stream(copy_stream):
            stage_x[b].copy_(hx, non_blocking=True)
            stage_r[b].copy_(hr, non_blocking=True)
            stage_y[b].copy_(hy, non_blocking=True)
            copied[b].record(copy_stream)

    def step_e2e(s):
        b = s & 1
        main = torch.cuda.current_stream()
        if state['next'] != s:               # first step of a run: nothing was prefetched
            upload(s)
        upload(s + 1)                        # the next batch travels while this one is computed
        state['next'] = s + 1
        main.wait_event(copied[b])
        st = C.c_void_p(main.cuda_stream)
        lib.dpp_augment_fwd(C.c_void_p(stage_x[b].data_ptr()), C.c_void_p(stage_r[b].data_ptr()),
                            C.c_void_p(eng.t_in.buf.data_ptr()), B, 128, 128, st)
        eng.y_in.copy_(stage_y[b], non_blocking=True)
        consumed[b].record(main)
        cost = eng.train_step(None, use_graph=True)
        cost_host.copy_(cost, non_blocking=False)            # the float train_model() returns

    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3)) if not args.no_e2e else float('nan')
    e2e_val = B * world / (ms_e2e / 1000.)

    clk = clocks.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel class, timed live with CUDA events on this stream
    roof = measure_dominant_kernel(eng, torch) if not args.no_roofline else None
    # ---- kernel launches per step (counted from the engine's op list)
    n_launch = count_launches(eng) + 1

    # ---- e2e again, with the host-side record preparation INSIDE the timed region (single GPU only: an exception on
    # one rank must not leave the others waiting in a collective).  Extra evidence, never the headline: it runs after
    # every contract measurement has been taken, and any failure is reported in the key instead of breaking the line.
    e2e_prep = None
    if world == 1 and not args.no_e2e:
        try:
            rng2 = np.random.RandomState(777)
            hx = torch.empty((B, 128, 128), dtype=torch.float32).pin_memory()
            hr = torch.empty((B, rec_bytes), dtype=torch.uint8).pin_memory()
            hy = torch.empty((B, E), dtype=torch.float32).pin_memory()
            xs_host = ds['x'][:, 0]
            copied2 = torch.cuda.Event()

            def prep():
                idxs = rng2.randint(0, N_RESIDENT, B)
                r, yv, _ = records_for(ds, comp, mean, idxs, rng2)
                r = r.copy()
                src = r['src_index'].copy()
                r['src_index'] = np.arange(B, dtype=np.int32)
                hx.numpy()[...] = xs_host[src]
                hr.numpy()[...] = r.view(np.uint8).reshape(B, -1)
                hy.numpy()[...] = yv

            def step_prep(s):
                main = torch.cuda.current_stream()
                stage_x[0].copy_(hx, non_blocking=True)
                stage_r[0].copy_(hr, non_blocking=True)
                stage_y[0].copy_(hy, non_blocking=True)
                copied2.record(main)
                st = C.c_void_p(main.cuda_stream)
                lib.dpp_augment_fwd(C.c_void_p(stage_x[0].data_ptr()), C.c_void_p(stage_r[0].data_ptr()),
                                    C.c_void_p(eng.t_in.buf.data_ptr()), B, 128, 128, st)
                eng.y_in.copy_(stage_y[0], non_blocking=True)
                cost = eng.train_step(None, use_graph=True)
                copied2.synchronize()            # the pinned buffers have been read: the next batch may overwrite them
                prep()                           # host: draws + records + labels of the NEXT batch, under this step
                cost_host.copy_(cost, non_blocking=False)
            prep()
            ms_prep = timed(step_prep, args.steps, max(args.warmup, 3))
            e2e_prep = {"value": B / (ms_prep / 1000.), "unit": UNIT, "ms_per_step": ms_prep,
                        "note": "as e2e, plus the random draws, augmentation records and embedded labels of every batch "
                                "computed on one host thread inside the timed region (overlapping the previous step)"}
        except Exception as exc:                 # pragma: no cover
            e2e_prep = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
    if rank != 0:
        if dist is not None:
            shutdown_distributed(dist, eng)
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = cpu_threads()
        t = cpu_step_time(ds, comp, mean, 3, threads)[1:]
        v = B / float(np.mean(t))
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 steps of the batch-128 workload (cv2 augmentation + torch-CPU ResNet fwd/bwd/ADAM) after "
                         "1 warm-up step"}
    h2d = B * 128 * 128 * 4 + B * rec_bytes + B * E * 4
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {0: "f32", 1: "tf32x3", 2: "tf32"}[precision], "data": "synthetic",
        "config": {"workload": "NYU posereg_embedding ResNet(type 0, 30-D embedding) training, batch 128/GPU, "
                               "aug rot/com/none, %d resident crops" % N_RESIDENT,
                   "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "inputs (134 MB of crops + 0.9 GB of activations per step) exceed the 126 MB L2",
                   "precision_mode": precision},
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4,
                "note": "per step: crops + augmentation records + labels copied from pinned host memory, cost read "
                        "back; the records are built on the host BEFORE the timed region (vectorised, ~2 ms per batch "
                        "on one thread, i.e. less than a step); e2e_with_host_prep times them too"},
        "e2e_with_host_prep": e2e_prep,
        "gpu_launches": n_launch * args.steps,
        "clocks": clk,
        "roofline": roof,
        "cpu_baseline": cpu,
        "tensor_fraction_whole_step": value / world * TRAIN_GFLOP_PER_FRAME / 1000. / peaks()[1],
    }
    print(json.dumps(out))
    if dist is not None:
        shutdown_distributed(dist, eng)


def count_launches(eng):
    """kernels of OUR library launched per training step (checked against the ncu launch list: 275 for the ResNet)"""
    n = 1 + 2 + 1 + 1      # loss, adam + tick, weight-image pack, ema (the two arena fills are memset nodes)
    n += len(eng.bns)      # one BN-backward apply per BatchNorm
    for op in eng.ops:
        k = op['kind']
        if k == 'conv':
            n += 3                  # fwd, wgrad, dgrad
        elif k == 'convpool':
            n += 2
        elif k == 'fc':
            n += 2 + 3              # fwd: GEMM + epilogue; bwd: pre-pass + 2 GEMMs
        elif k == 'bn_apply':
            n += 2                  # materialised BN+ReLU and its backward reduce
    return n


def measure_dominant_kernel(eng, torch):
    """Roofline of the dominant kernel, timed live with CUDA events on the launching stream.

    The dominant kernel of the step is the implicit-GEMM convolution (k_conv_tc in precision 1/2, k_igemm in
    precision 0: 126 of the ~285 launches and ~half of the GPU time, see profiles/).  All 63 ConvLayer forward
    launches of the batch-128 net are issued back to back between two events (their 0.9 GB of activations exceed
    the 126 MB L2, so no flush is needed) and the average launch is compared with HBM speed: per launch the
    ALGORITHMIC bytes are 4 * (input + output [+ residual]) elements - at ~20 FLOP/B the
    layers sit far below the tensor ridge, so the bound is HBM; the tensor-pipe fraction is
    reported next to it."""
    import ctypes as C
    from dpp_b200.lib import lib
    convs = [op for op in eng.ops if op['kind'] == 'conv']
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    calls, bytes_alg, macs = [], 0.0, 0.0
    for op in convs:
        d = eng._conv_desc(op)
        L = op['layer']
        bnref = eng._bnref(op['in_bn'], op['src'], True) if op['in_bn'] is not None else None
        stats = eng._stats_ptr(op['out_bn']) if op['out_bn'] is not None else None
        res = op['residual']
        calls.append((d, op['src'].buf.data_ptr(), bnref, eng.pview(L.W).data_ptr(), eng.pview(L.b).data_ptr(),
                      res.buf.data_ptr() if res is not None else None, op['dst'].buf.data_ptr(), stats))
        bytes_alg += 4.0 * (d.N * d.H * d.W * d.Cin + d.N * d.Ho * d.Wo * d.Cout * (2 if res is not None else 1))
        macs += float(d.N) * d.Ho * d.Wo * d.Cout * d.k * d.k * d.Cin

    def chain():
        for d, x, bnref, w, b, r, y, stats in calls:
            lib.dpp_conv2d_fwd(C.byref(d), C.c_void_p(x), C.byref(bnref) if bnref else None, C.c_void_p(w),
                               C.c_void_p(b), C.c_void_p(r) if r else None, C.c_void_p(y), stats, st)
    for _ in range(3):
        chain()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        chain()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * len(calls))          # average launch
    burst, sustained, hbm, how = peaks()
    per_launch_bytes = bytes_alg / len(calls)
    gbs = per_launch_bytes / (ms * 1e-3) / 1e9
    tflops = 2.0 * macs / len(calls) / (ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get('conv_fwd_avg_bytes_per_launch')
        except Exception:
            traffic = None
    kname = {0: 'k_igemm', 1: 'k_conv_tc<*,2>', 2: 'k_conv_tc<*,1>'}[eng.precision]
    return {"bound": "hbm", "kernel": "%s: the %d ConvLayer forward launches of the step, back to back" % (kname, len(calls)),
            "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": traffic,
            "peak_source": "%s HBM copy bandwidth (MEASURED_PEAKS.json)" % how,
            "ms_per_launch": ms, "launches_timed": reps * len(calls), "alg_bytes_per_launch": per_launch_bytes,
            "tensor_tflops": tflops, "tensor_peak_tflops_tf32": burst / 2.0,
            "tensor_frac_tf32": tflops * (3.0 if eng.precision == 1 else 1.0) / (burst / 2.0),
            "note": "tensor_frac counts the MMA work actually issued (3 TF32 MMAs per product in 3xTF32 mode)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true', help='skip the live kernel timing (profiler runs)')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer arm (profiler runs)')
    ap.add_argument('--workload', default='train', choices=['train', 'cascade'],
                    help="train = BASELINE configs[1] (the headline); cascade = configs[4], tools/bench_cascade.py")
    args, rest = ap.parse_known_args()
    if args.workload == 'cascade':
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import bench_cascade
        return bench_cascade.main(['--steps', str(args.steps), '--warmup', str(args.warmup)] + rest
                                  + (['--no-cpu-baseline'] if args.no_cpu_baseline else []))
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
