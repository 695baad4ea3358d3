#!/usr/bin/env python
"""bench.py - the driver's measurement contract for the DeepPrior++ hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic depth crops:
  dpp_augment_fwd (from the HBM-resident ORIGINAL crops) -> ResNet (type 0, nDims 30) forward -> cost ->
  backward -> [NCCL gradient exchange] -> ADAM -> BN running-stat EMA.

Workloads (BASELINE.json configs):
  train   (default) configs[1]: NYU posereg_embedding ResNet training, batch 128 per GPU, aug com/rot/none.
          N > 1: every rank runs the same per-GPU batch (weak scaling, global batch 128*N).
  icvl512 configs[2]: ICVL 16-joint, GLOBAL batch 512 split over the N ranks (strong scaling), gradient exchange in
          stage-ordered buckets; --syncbn sums the BatchNorm statistics over the ranks as well.
  poseregnet  the network the reference's entry scripts actually train (PoseRegNet type 0), batch 128 per GPU.
  msra15  configs[3]: MSRA15 21-joint (y-flipped projection, per-subject cubes), aug com/rot/sc/none, batch 128 per GPU.
  cascade configs[4]: tools/bench_cascade.py.

  value : whole-job frames/s with all inputs resident in HBM when the timed region starts
  e2e   : the same metric through HOST buffers: per step the random draws, the augmentation records and the
          embedded labels are computed on the host INSIDE the timed region, the batch's crops + records + labels are
          copied from pinned host memory and the cost is read back (what trainer.train_model() returns)
  trainer_api : frames/s of PoseRegNetTrainer.train() itself (the reference-facing call), epochs of a resident set
  strong : (default workload, every N) the icvl512 step at this N, so that the driver's 1/2/4/8 runs also hold a
          strong-scaling curve
  --impl reference : the CPU oracle restatement (torch-CPU + cv2) of the same step on the host cores (the
          reference's own Theano path cannot run here: no Theano / Python 2).
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
STRONG_GLOBAL_B = 512              # BASELINE config 3
AUG_MODES = ['com', 'rot', 'none']

WORKLOADS = {
    # name: (dataset, aug modes, per-GPU batch ('B' = bench.B) or None = global batch / N, global batch, description)
    'train': ('NYU', ['com', 'rot', 'none'], 'B', None,
              "NYU posereg_embedding ResNet(type 0, 30-D embedding) training, batch 128/GPU, aug rot/com/none"),
    'icvl512': ('ICVL', ['com', 'rot', 'none'], None, 'STRONG',
                "ICVL 16-joint posereg_embedding ResNet(type 0, 30-D embedding) training, GLOBAL batch 512 split over "
                "the ranks, aug rot/com/none"),
    'poseregnet': ('NYU', ['com', 'rot', 'none'], 'B', None,
                   "NYU posereg_embedding PoseRegNet(type 0: 3 x 8-filter conv+pool, FC 1024 - dropout - FC 1024 - dropout - 30-D "
                   "embedding) training - the network the reference's main_*_posereg_embedding.py scripts build -, batch "
                   "128/GPU, aug rot/com/none"),
    'msra15': ('MSRA15', ['com', 'rot', 'sc', 'none'], 'B', None,
               "MSRA15 21-joint crossval config, ResNet(type 0, 30-D embedding) training, batch 128/GPU, "
               "aug rot/scale/com/none on the device"),
}


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


def shutdown_distributed(dist, engs=()):
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
        for eng in engs:
            if eng is not None:
                eng._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    timer.cancel()


def make_workload(seed=23455, dataset='NYU', n=None):
    """synthetic set + PCA stand-in"""
    from data import synthetic
    ds = synthetic.generate(dataset, N_RESIDENT if n is None else n, seed=seed)
    comp, mean = synthetic.random_orthonormal_pca(E, ds['gt3D'].shape[1] * 3, seed=1)
    return ds, comp, mean


def records_for(ds, comp, mean, idxs, rng, aug_modes=None):
    """host side of one batch: draws in the reference order (nettrainer.py:954-957), dpp_aug_rec records and
    embedded labels in one vectorised pass (HandDetector.aug_records_batch)"""
    aug_modes = AUG_MODES if aug_modes is None else aug_modes
    hd = ds['hd']
    draws = []
    for _ in idxs:
        mode = rng.randint(0, len(aug_modes)); off = rng.randn(3) * 5.; rot = rng.uniform(-180., 180.)
        sc = abs(1. + rng.randn() * 0.02)
        draws.append((mode, off, rot, sc))
    idxs = np.asarray(idxs)
    com = hd._toimg(ds['com3D'][idxs])
    recs, labs = hd.aug_records_batch(idxs, [aug_modes[d[0]] for d in draws], np.array([d[1] for d in draws]),
                                      np.array([d[2] for d in draws]), np.array([d[3] for d in draws]), com,
                                      ds['cube'][idxs], ds['M'][idxs], ds['gt3Dcrop'][idxs])
    ys = np.dot(labs.reshape(len(idxs), -1).astype(np.float64) - mean, comp.T)
    return recs, np.asarray(ys, np.float32), draws


# ---- host preparation in worker processes (the reference augments in 8 worker processes, trainer/nettrainer.py:59,666-689)
_PREP = {}


def _prep_init(dataset, n_resident, seed, aug_modes):
    """worker initialiser: every worker rebuilds the (deterministic) synthetic set and the PCA stand-in once"""
    global N_RESIDENT
    N_RESIDENT = n_resident
    _PREP['ds'], _PREP['comp'], _PREP['mean'] = make_workload(seed=seed, dataset=dataset, n=n_resident)
    _PREP['aug'] = list(aug_modes)


def _prep_job(job):
    """one batch: random draws, augmentation records, embedded labels (the host work of one step).  job = (seed, nb) returns
    the arrays; job = (seed, nb, shm name, slot, slot bytes) also gathers the batch's crops and writes crops | records |
    labels into staging slot `slot` of the shared-memory ring (page-locked by the parent) and returns the slot number."""
    seed, nb = job[:2]
    rng = np.random.RandomState(seed)
    idxs = rng.randint(0, _PREP['ds']['x'].shape[0], nb)
    r, yv, _ = records_for(_PREP['ds'], _PREP['comp'], _PREP['mean'], idxs, rng, _PREP['aug'])
    r = r.copy()
    src = r['src_index'].astype(np.int64)
    r['src_index'] = np.arange(nb, dtype=np.int32)                    # records index the staged batch
    rb = r.view(np.uint8).reshape(nb, -1)
    yv = np.ascontiguousarray(yv, np.float32)
    if len(job) == 2:
        return src, rb.copy(), yv
    name, slot, slot_bytes = job[2:]
    shm = _PREP.get('shm')
    if shm is None or shm.name != name:
        from multiprocessing import shared_memory
        shm = _PREP['shm'] = shared_memory.SharedMemory(name=name)
    base = slot * slot_bytes
    xb = nb * 128 * 128 * 4
    x = np.ndarray((nb, 128, 128), np.float32, buffer=shm.buf, offset=base)
    np.take(_PREP['ds']['x'][:, 0], src, axis=0, out=x)
    np.ndarray(rb.shape, np.uint8, buffer=shm.buf, offset=base + xb)[...] = rb
    np.ndarray(yv.shape, np.float32, buffer=shm.buf, offset=base + xb + rb.size)[...] = yv
    return slot


def prep_workers():
    """worker processes per rank for the e2e host preparation: DPP_PREP_WORKERS, default min(4, cores / ranks - 1)"""
    e = os.environ.get('DPP_PREP_WORKERS')
    if e is not None:
        return max(0, int(e))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    world = int(os.environ.get('WORLD_SIZE', '1'))
    return max(0, min(4, cores // max(world, 1) - 1))


def first_batch(ds, comp, mean, nb, seed=1234):
    """the batch the cost check runs on: same recipe in bench.py (device) and tests/golden/make_bench_golden.py (oracle)"""
    rng = np.random.RandomState(seed)
    idxs = rng.randint(0, ds['x'].shape[0], nb)
    r, y, draws = records_for(ds, comp, mean, idxs, rng)
    return idxs, r, y, draws


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle restatement (reference-equivalent CPU path)
# ---------------------------------------------------------------------------------------------
def cpu_threads():
    """all host cores the process may use (DPP_CPU_THREADS overrides)"""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    return int(os.environ.get('DPP_CPU_THREADS', str(avail)))


_AUG_STATE = None


def _aug_worker(job):
    """one of the para_num_proc = 8 augmentation workers of the reference (trainer/nettrainer.py:59,666-689)"""
    from oracle import augment as OA
    ds, comp, mean = _AUG_STATE
    idxs, draws = job
    cam = OA.Camera(**OA.NYU_CAM)
    ohd = OA.Hand(cam, use_cv2=True)
    return OA.augment_poses(ds['x'], ds['com3D'], ds['cube'], ds['M'], ds['gt3Dcrop'], list(idxs), draws,
                            AUG_MODES, cam, ohd, pca_mean=mean, pca_components=comp)


def cpu_step_time(ds, comp, mean, steps, threads, workers=8):
    import multiprocessing
    import torch
    from oracle import nets as ON, augment as OA
    global _AUG_STATE
    _AUG_STATE = (ds, comp, mean)
    pool = multiprocessing.get_context('fork').Pool(workers) if workers > 1 else None     # forked before torch threads spin up
    torch.set_num_threads(threads)
    ON.set_dtype(torch.float32)          # the timed CPU arm runs in the reference's precision
    onet = ON.build_resnet(np.random.RandomState(23455), type=0, batchSize=B, numJoints=1, nDims=E)
    adam = ON.Adam(onet.params)
    rng = np.random.RandomState(99)
    times = []
    try:
        for s in range(steps):
            idxs = rng.randint(0, ds['x'].shape[0], B)
            draws = [OA.draw_aug_params(rng, len(AUG_MODES)) for _ in idxs]
            t0 = time.time()
            if pool is not None:
                per = (B + workers - 1) // workers
                parts = pool.map(_aug_worker, [(idxs[i:i + per], draws[i:i + per]) for i in range(0, B, per)])
                x = np.concatenate([p[0] for p in parts]); y = np.concatenate([p[1] for p in parts])
            else:
                x, y = _aug_worker((idxs, draws))
            ON.train_step(onet, adam, torch.from_numpy(x), torch.from_numpy(y), 1e-4, 1, E)
            times.append(time.time() - t0)
    finally:
        if pool is not None:
            pool.terminate()
    return times


def cpu_arm(ds, comp, mean, steps, warm):
    """times the CPU restatement with ALL host cores and with 16 threads (torch-CPU convolutions on 8x8..64x64 maps
    stop scaling somewhere in between) and reports the faster of the two - the reference arm gets every advantage"""
    allc = cpu_threads()
    tried = {}
    for th in sorted(set([allc, min(allc, 16)]), reverse=True):
        t = cpu_step_time(ds, comp, mean, steps + warm, th)[warm:]
        tried[th] = B / float(np.mean(t))
    best = max(tried, key=lambda k: tried[k])
    return tried[best], best, tried


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ds, comp, mean = make_workload()
    steps = max(1, min(args.steps, 4))
    warm = 1 if args.warmup > 0 else 0
    val, threads, tried = cpu_arm(ds, comp, mean, steps, warm)
    ms = 1000. * B / val
    sample = "%d step(s) of the batch-128 workload (cv2 augmentation in 8 worker processes + torch-CPU ResNet " \
             "fwd/bwd/ADAM); frames/s by torch thread count: %s; host cores available: %d" % (
                 steps, {k: round(v, 1) for k, v in tried.items()}, cpu_threads())
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS['train'][4], "global_batch": B,
                   "note": "CPU oracle restatement of the reference path (Theano cannot run here); host cores only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
class Ctx(object):
    """device, ranks and the timing helper shared by the measurements of one run"""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
            self.dist = dist
        self.dev = torch.device('cuda', self.local)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed steps, then exactly `steps` timed ones between barrier + synchronize, CUDA events, max over ranks"""
        torch = self.torch
        for s in range(warmup):
            fn(s)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(warmup + s)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.dist is not None:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms / steps


class StepRunner(object):
    """One training workload on this rank: engine, resident crops, pre-built record sets, the step functions."""

    def __init__(self, ctx, dataset, aug_modes, batch, total_steps, syncbn=False, seed0=23455, net_kind='resnet'):
        import ctypes as C
        torch = ctx.torch
        from dpp_b200.lib import lib, AUG_REC_DTYPE
        from dpp_b200.engine import Engine
        from net.resnet import ResNet, ResNetParams
        self.C, self.lib, self.ctx, self.nb, self.aug_modes = C, lib, ctx, batch, aug_modes
        self.precision = int(os.environ.get('DPP_PRECISION', '1'))
        self.dataset, self.seed = dataset, seed0 + ctx.rank
        self.ds, self.comp, self.mean = make_workload(seed=seed0 + ctx.rank, dataset=dataset)
        if net_kind == 'poseregnet':
            from net.poseregnet import PoseRegNet, PoseRegNetParams
            net = PoseRegNet(np.random.RandomState(23455), cfgParams=PoseRegNetParams(type=0, nChan=1, wIn=128, hIn=128,
                                                                                      batchSize=batch, numJoints=1, nDims=E))
        else:
            net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=batch, numJoints=1, nDims=E))
        self.eng = eng = Engine(net, precision=self.precision)
        net._eng = eng
        eng._alloc_training()
        if ctx.world > 1:
            dist = ctx.dist
            eng.set_world(ctx.world, lambda g: dist.all_reduce(g), rank=ctx.rank, syncbn=syncbn, dist=dist)
        self.rng = np.random.RandomState(1234 + ctx.rank)
        self.nrec = min(total_steps, 64)             # distinct record sets, cycled
        self.recs_all, self.ys_all = [], []
        for s in range(self.nrec):
            idxs = self.rng.randint(0, N_RESIDENT, batch)
            r, y, _ = records_for(self.ds, self.comp, self.mean, idxs, self.rng, aug_modes)
            self.recs_all.append(r); self.ys_all.append(y)
        dev = ctx.dev
        self.crops_dev = torch.from_numpy(self.ds['x'][:, 0].copy()).to(dev)
        recs_np = np.concatenate(self.recs_all)
        self.recs_dev = torch.from_numpy(recs_np.view(np.uint8).reshape(len(recs_np), -1).copy()).to(dev)
        self.ys_dev = torch.from_numpy(np.concatenate(self.ys_all)).to(dev)
        self.rec_bytes = np.dtype(AUG_REC_DTYPE).itemsize
        eng.set_lr(1e-4)

    def augment(self, crops_ptr, recs_ptr):
        C, torch = self.C, self.ctx.torch
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.lib.dpp_augment_fwd(C.c_void_p(crops_ptr), C.c_void_p(recs_ptr), C.c_void_p(self.eng.t_in.buf.data_ptr()),
                                 self.nb, 128, 128, st)

    def step_resident(self, s):
        k = s % self.nrec
        nb = self.nb
        self.augment(self.crops_dev.data_ptr(), self.recs_dev.data_ptr() + k * nb * self.rec_bytes)
        self.eng.y_in.copy_(self.ys_dev[k * nb:(k + 1) * nb], non_blocking=True)
        self.eng.train_step(None, use_graph=True)

    def cost_check(self):
        """ONE step through the timed code path (CUDA graph, batch of the full size) on a fixed batch from the initial
        weights, its cost compared with the float64 oracle's value for the same batch (tests/golden/bench_first_step.json,
        written by tests/golden/make_bench_golden.py): the measured path is itself checked, not only its small-batch
        relatives.  Only defined for the default workload on rank 0's data (seed 23455)."""
        torch = self.ctx.torch
        p = os.path.join(ROOT, 'tests', 'golden', 'bench_first_step.json')
        if not os.path.exists(p):
            return {"skipped": "no golden file"}
        gold = json.load(open(p))
        if gold.get('batch') != self.nb or gold.get('n_resident') != N_RESIDENT:
            return {"skipped": "golden value is for batch %s of %s resident crops" % (gold.get('batch'), gold.get('n_resident'))}
        idxs, r, y, _ = first_batch(self.ds, self.comp, self.mean, self.nb)
        dev = self.ctx.dev
        rd = torch.from_numpy(r.view(np.uint8).reshape(len(r), -1).copy()).to(dev)
        self.augment(self.crops_dev.data_ptr(), rd.data_ptr())
        self.eng.y_in.copy_(torch.from_numpy(y).to(dev))
        cost = float(self.eng.train_step(None, use_graph=True).cpu()[0])
        rel = abs(cost - gold['cost']) / abs(gold['cost'])
        return {"engine": cost, "oracle": gold['cost'], "rel": rel, "ok": bool(rel < 1e-4), "tolerance": 1e-4,
                "batch": self.nb, "golden": "tests/golden/bench_first_step.json"}


def measure_e2e(run, ctx, steps, warmup, with_prep):
    """host buffers in, cost out, every step.  with_prep: the draws, the augmentation records and the embedded labels of
    batch s+1 are computed on the host (one thread) while step s runs on the device - all inside the timed region."""
    import ctypes as C
    torch = ctx.torch
    eng, nb, dev = run.eng, run.nb, ctx.dev
    rec_bytes = run.rec_bytes
    stage_x = [torch.empty((nb, 128, 128), dtype=torch.float32, device=dev) for _ in range(2)]
    stage_r = [torch.empty((nb, rec_bytes), dtype=torch.uint8, device=dev) for _ in range(2)]
    stage_y = [torch.empty((nb, E), dtype=torch.float32, device=dev) for _ in range(2)]
    cost_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    pool = None
    if with_prep:
        rng2 = np.random.RandomState(777 + ctx.rank)
        nwork = prep_workers()
        if nwork > 0:
            try:
                import multiprocessing
                pool = multiprocessing.get_context('spawn').Pool(
                    nwork, initializer=_prep_init, initargs=(run.dataset, run.ds['x'].shape[0], run.seed, list(run.aug_modes)))
                pool.map(_prep_job, [(1, nb)] * nwork)                # workers up and initialised before anything is timed
            except Exception as ex:                                   # no worker processes on this box: one host thread
                print("bench: host-preparation workers unavailable (%s); preparing on the main thread" % ex, file=sys.stderr)
                pool = None
        ring = None
        if pool is not None:
            try:
                # staging ring in shared memory, page-locked here: the workers write a finished batch (crops gathered from
                # their copy of the set, records, labels) straight into it; the parent only issues the copies
                from multiprocessing import shared_memory
                xb, rbts, ybts = nb * 128 * 128 * 4, nb * rec_bytes, nb * E * 4
                slot_bytes = (xb + rbts + ybts + 4095) // 4096 * 4096
                nslot = 2 * nwork + 2
                try:                                              # a small /dev/shm (container default: 64 MB) cannot hold the ring:
                    vfs = os.statvfs('/dev/shm')                  # writing past it would kill the workers with SIGBUS
                    if vfs.f_bavail * vfs.f_frsize < 2 * nslot * slot_bytes * max(ctx.world, 1):
                        raise RuntimeError("/dev/shm has %d MB free" % (vfs.f_bavail * vfs.f_frsize >> 20))
                except OSError:
                    pass
                shm = shared_memory.SharedMemory(create=True, size=nslot * slot_bytes)
                whole = torch.frombuffer(shm.buf, dtype=torch.uint8)
                rc = torch.cuda.cudart().cudaHostRegister(whole.data_ptr(), nslot * slot_bytes, 0)
                if int(rc) != 0:
                    raise RuntimeError("cudaHostRegister failed (%s)" % rc)
                ring = {'shm': shm, 'whole': whole, 'slot_bytes': slot_bytes, 'nslot': nslot, 'free': list(range(nslot)),
                        'inflight': [], 'xb': xb, 'rb': rbts, 'yb': ybts, 'busy': {}}
            except Exception as ex:
                print("bench: shared-memory staging unavailable (%s); gathering on the main thread" % ex, file=sys.stderr)
                ring = None
        hx = [torch.empty((nb, 128, 128), dtype=torch.float32).pin_memory() for _ in range(2)]
        hr = [torch.empty((nb, rec_bytes), dtype=torch.uint8).pin_memory() for _ in range(2)]
        hy = [torch.empty((nb, E), dtype=torch.float32).pin_memory() for _ in range(2)]
        xs_host = run.ds['x'][:, 0]
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        copy_stream = torch.cuda.Stream(device=dev)
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        xs_t = torch.from_numpy(xs_host)
        pending = []                                                  # batches being prepared by the workers
        jobno = [0]

        def submit():
            jobno[0] += 1
            pending.append(pool.apply_async(_prep_job, ((1000003 * (ctx.rank + 1) + jobno[0], nb),)))

        def ring_submit():
            r_ = ring
            while r_['free'] and len(r_['inflight']) < 2 * nwork:
                slot = r_['free'].pop(0)
                ev = r_['busy'].pop(slot, None)
                if ev is not None:
                    ev.synchronize()                                  # the copies out of this slot have completed
                jobno[0] += 1
                r_['inflight'].append(pool.apply_async(
                    _prep_job, ((1000003 * (ctx.rank + 1) + jobno[0], nb, r_['shm'].name, slot, r_['slot_bytes']),)))

        def ring_upload(b):
            r_ = ring
            ring_submit()
            slot = r_['inflight'].pop(0).get(timeout=300)             # a finished batch (a dead worker must not hang the run)
            o = slot * r_['slot_bytes']
            w = r_['whole']
            copy_stream.wait_event(consumed[b])
            with torch.cuda.stream(copy_stream):
                stage_x[b].view(torch.uint8).view(-1).copy_(w[o:o + r_['xb']], non_blocking=True)
                stage_r[b].view(-1).copy_(w[o + r_['xb']:o + r_['xb'] + r_['rb']], non_blocking=True)
                stage_y[b].view(torch.uint8).view(-1).copy_(w[o + r_['xb'] + r_['rb']:o + r_['xb'] + r_['rb'] + r_['yb']],
                                                            non_blocking=True)
                copied[b].record(copy_stream)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            r_['busy'][slot] = ev
            r_['free'].append(slot)
            ring_submit()

        def prep(b):
            if ring is not None:
                return                                                # the batch is prepared by a worker, see ring_upload
            if pool is not None:
                while len(pending) < 2 * nwork:                       # keep the workers busy a few batches ahead
                    submit()
                src, rbytes, yv = pending.pop(0).get(timeout=300)
                torch.index_select(xs_t, 0, torch.from_numpy(src), out=hx[b])     # the batch's crops into pinned memory
                hr[b].numpy()[...] = rbytes
                hy[b].numpy()[...] = yv
                return
            idxs = rng2.randint(0, N_RESIDENT, nb)
            r, yv, _ = records_for(run.ds, run.comp, run.mean, idxs, rng2, run.aug_modes)
            r = r.copy()
            src = r['src_index'].copy()
            r['src_index'] = np.arange(nb, dtype=np.int32)            # records index the staged batch
            hx[b].numpy()[...] = xs_host[src]
            hr[b].numpy()[...] = r.view(np.uint8).reshape(nb, -1)
            hy[b].numpy()[...] = yv

        def upload(b):
            if ring is not None:
                return ring_upload(b)
            copy_stream.wait_event(consumed[b])                       # the step that read staging set b is done with it
            with torch.cuda.stream(copy_stream):
                stage_x[b].copy_(hx[b], non_blocking=True)
                stage_r[b].copy_(hr[b], non_blocking=True)
                stage_y[b].copy_(hy[b], non_blocking=True)
                copied[b].record(copy_stream)

        state = {'ready': None}

        def step(s):
            b = s & 1
            main = torch.cuda.current_stream()
            if state['ready'] != s:                                   # first step of a run: nothing prepared yet
                prep(b); upload(b)
            main.wait_event(copied[b])
            run.augment(stage_x[b].data_ptr(), stage_r[b].data_ptr())
            eng.y_in.copy_(stage_y[b], non_blocking=True)
            consumed[b].record(main)
            cost = eng.train_step(None, use_graph=True)
            if ring is None:
                copied[b ^ 1].synchronize()                           # pinned set b^1 has been read by its last upload
            prep(b ^ 1)                                               # host work of the NEXT batch, under this step
            upload(b ^ 1)
            state['ready'] = s + 1
            cost_host.copy_(cost, non_blocking=False)                 # the float train_model() returns
        for e in copied + consumed:
            e.record()
        if pool is not None:
            note = "per step, inside the timed region: random draws + augmentation records + label embedding of one batch in %d " \
                   "worker processes (the reference augments in 8), the batch's crops gathered into %s, crops + " \
                   "records + labels copied to the device, cost read back" % (
                       nwork, "a page-locked shared-memory ring by the workers" if ring is not None else "pinned memory")
        else:
            note = "per step, inside the timed region: random draws + augmentation records + label embedding on one host " \
                   "thread (overlapping the previous step), crops + records + labels copied from pinned host memory, cost read back"
    else:
        host_batches = []
        for k in range(run.nrec):
            r = run.recs_all[k].copy()
            src = r['src_index'].copy()
            r['src_index'] = np.arange(nb, dtype=np.int32)
            host_batches.append((torch.from_numpy(run.ds['x'][src, 0].copy()).pin_memory(),
                                 torch.from_numpy(r.view(np.uint8).reshape(nb, -1).copy()).pin_memory(),
                                 torch.from_numpy(run.ys_all[k].copy()).pin_memory()))
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        state = {'next': None}

        def upload(s):
            b = s & 1
            hx_, hr_, hy_ = host_batches[s % run.nrec]
            copy_stream.wait_event(consumed[b])
            with torch.cuda.stream(copy_stream):
                stage_x[b].copy_(hx_, non_blocking=True)
                stage_r[b].copy_(hr_, non_blocking=True)
                stage_y[b].copy_(hy_, non_blocking=True)
                copied[b].record(copy_stream)

        def step(s):
            b = s & 1
            main = torch.cuda.current_stream()
            if state['next'] != s:
                upload(s)
            upload(s + 1)
            state['next'] = s + 1
            main.wait_event(copied[b])
            run.augment(stage_x[b].data_ptr(), stage_r[b].data_ptr())
            eng.y_in.copy_(stage_y[b], non_blocking=True)
            consumed[b].record(main)
            cost = eng.train_step(None, use_graph=True)
            cost_host.copy_(cost, non_blocking=False)
        note = "records built on the host BEFORE the timed region; per step crops + records + labels copied from pinned " \
               "host memory, cost read back"
    ms = ctx.timed(step, steps, warmup)
    if pool is not None:
        torch.cuda.synchronize()
        pool.terminate()
        pool.join()
        if ring is not None:
            torch.cuda.cudart().cudaHostUnregister(ring['whole'].data_ptr())
            whole = ring.pop('whole')
            del whole
            try:
                ring['shm'].close()
                ring['shm'].unlink()
            except Exception:
                pass
    return ms, note


class _PcaProj(object):
    """sklearn-PCA stand-in handed to the trainer (module level: it travels to the record workers)"""
    def __init__(self, comp, mean):
        self.comp, self.mean = comp, mean

    def transform(self, lab):
        return np.dot(np.asarray(lab, np.float64) - self.mean, self.comp.T)


def measure_trainer_api(ctx, epochs=6, n_train=2048):
    """frames/s of PoseRegNetTrainer.train() - the call the reference's entry scripts make - on a resident synthetic NYU
    set: per epoch 16 minibatches of 128, the augmented set regenerated on the device from fresh host records
    (force_macrobatch_reload, as main_nyu_posereg_embedding.py:112 sets it), the cost returned to the host every
    minibatch; validation switched off inside the timed epochs (validation_frequency beyond the run)."""
    import contextlib
    import io
    from data import synthetic
    from net.resnet import ResNet, ResNetParams
    from trainer.poseregnettrainer import PoseRegNetTrainer, PoseRegNetTrainerParams
    ds, comp, mean = make_workload(seed=4242, n=n_train)
    y = np.dot(ds['gt3Dcrop'].reshape(n_train, -1).astype(np.float64) - mean, comp.T).astype(np.float32)
    net = ResNet(np.random.RandomState(23455), cfgParams=ResNetParams(type=0, batchSize=B, numJoints=1, nDims=E))
    p = PoseRegNetTrainerParams()
    p.batch_size, p.learning_rate, p.weightreg_factor = B, 1e-4, 0.0
    p.force_macrobatch_reload, p.para_augment, p.para_num_proc = True, True, 8        # the reference's default: 8 record workers
    p.validation_frequency = 10 ** 9
    p.snapshot_last = 10 ** 9

    p.augment_fun_params = {'fun': 'augment_poses', 'args': {'normZeroOne': False, 'di': ds['importer'], 'hd': ds['hd'],
                                                              'aug_modes': AUG_MODES, 'proj': _PcaProj(comp, mean)}}
    with contextlib.redirect_stdout(io.StringIO()):
        tr = PoseRegNetTrainer(net, p, np.random.RandomState(23455), './')
        tr.verbose = False
        tr.setData(ds['x'], y, ds['x'][:B], y[:B])
        tr.addStaticData({'val_data_y3D': ds['gt3D'][:B], 'pca_data': comp.astype(np.float32), 'mean_data': mean.astype(np.float32)})
        tr.addManagedData({'train_data_cube': ds['cube'], 'train_data_com': ds['com3D'], 'train_data_M': ds['M'],
                           'train_gt3Dcrop': ds['gt3Dcrop']})
        tr.compileFunctions()
        tr.poseNet.save = lambda *a, **k: None               # no snapshot files from a benchmark
        tr.train(n_epochs=1)                                  # warm-up epoch: graph capture, first augmentation
        ctx.torch.cuda.synchronize()
        t0 = time.time()
        tr.train(n_epochs=epochs)
        ctx.torch.cuda.synchronize()
        dt = time.time() - t0
    frames = epochs * tr.getNumFullMiniBatches() * B
    net._eng.release()
    return {"value": frames / dt, "unit": UNIT, "epochs": epochs, "minibatches_per_epoch": tr.getNumFullMiniBatches(),
            "wall_s": dt,
            "note": "PoseRegNetTrainer.train() on %d resident crops: per epoch new draws + records in 8 worker processes, the "
                    "augmented set regenerated on the device, one train_model() call and one cost read-back per "
                    "minibatch, plus the initial validation pass of each train() call" % n_train}


def run_b200(args):
    ctx = Ctx()
    torch = ctx.torch
    rank, world = ctx.rank, ctx.world
    dataset, aug_modes, per_gpu, global_b, descr = WORKLOADS[args.workload]
    per_gpu = B if per_gpu == 'B' else per_gpu
    global_b = STRONG_GLOBAL_B if global_b == 'STRONG' else global_b
    if per_gpu is None:
        if global_b % world:
            raise SystemExit("global batch %d does not divide over %d ranks" % (global_b, world))
        nb, scaling = global_b // world, "strong"
    else:
        nb, scaling = per_gpu, "weak"
    warm = max(args.warmup, 3)
    total = warm + args.steps
    run = StepRunner(ctx, dataset, aug_modes, nb, total, syncbn=args.syncbn,
                     net_kind='poseregnet' if args.workload == 'poseregnet' else 'resnet')
    eng = run.eng
    check = None
    if args.workload == 'train' and not args.no_cost_check:
        # every rank steps (the training step of a data-parallel job contains collectives: all ranks or none); the
        # golden value is that of rank 0's data, and the cost is taken before the gradient exchange
        check = run.cost_check()
        if rank != 0:
            check = None
    clocks = ClockSampler(ctx.local)
    if rank == 0:
        clocks.start()
    ms_step = ctx.timed(run.step_resident, args.steps, warm)
    value = nb * world / (ms_step / 1000.)

    ms_e2e, e2e_note, ms_pre = float('nan'), "", None
    if not args.no_e2e:
        ms_e2e, e2e_note = measure_e2e(run, ctx, args.steps, warm, with_prep=True)
        ms_pre, _ = measure_e2e(run, ctx, args.steps, warm, with_prep=False)
    e2e_val = nb * world / (ms_e2e / 1000.)
    clk = clocks.stop() if rank == 0 else None

    roof = measure_dominant_kernel(eng, torch) if (not args.no_roofline and args.workload != 'poseregnet') else None
    n_launch = count_launches(eng) + 1

    # strong-scaling sample inside the default run: BASELINE config 3 (global batch 512 split over the ranks)
    strong = None
    if args.workload == 'train' and not args.no_strong:
        try:
            gb = STRONG_GLOBAL_B
            if gb % world == 0:
                eng._graphs.clear()
                r2 = StepRunner(ctx, 'ICVL', ['com', 'rot', 'none'], gb // world, 3 + min(args.steps, 10))
                ms2 = ctx.timed(r2.step_resident, min(args.steps, 10), 3)
                strong = {"workload": WORKLOADS['icvl512'][4], "global_batch": gb, "per_gpu_batch": gb // world,
                          "n_gpus": world, "ms_per_step": ms2, "value": gb / (ms2 / 1000.), "unit": UNIT,
                          "scaling": "strong", "batchnorm": "per-replica statistics",
                          "buckets": [list(b) for b in sorted(getattr(r2.eng, '_exchanged', []))]}
                r2.eng._graphs.clear()
                if world > 1:
                    r3 = StepRunner(ctx, 'ICVL', ['com', 'rot', 'none'], gb // world, 3 + min(args.steps, 10), syncbn=True)
                    ms3 = ctx.timed(r3.step_resident, min(args.steps, 10), 3)
                    strong["syncbn_ms_per_step"] = ms3
                    strong["syncbn_value"] = gb / (ms3 / 1000.)
                    r3.eng._graphs.clear()
                    r4 = StepRunner(ctx, 'ICVL', ['com', 'rot', 'none'], gb // world, 3 + min(args.steps, 10), syncbn='p2p')
                    ms4 = ctx.timed(r4.step_resident, min(args.steps, 10), 3)
                    r4.eng.check_barriers()
                    strong["syncbn_p2p_ms_per_step"] = ms4
                    strong["syncbn_p2p_value"] = gb / (ms4 / 1000.)
                    strong["syncbn_note"] = "syncbn: one NCCL all-reduce per BatchNorm and direction (122 per step); " \
                                            "syncbn_p2p: dpp_stats_exchange, a one-shot exchange over NVLink peer memory"
                    r4.eng._graphs.clear()
        except Exception as exc:                 # pragma: no cover - extra evidence must not break the contract line
            if world > 1:
                raise
            strong = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}

    trainer_api = None
    if world == 1 and args.workload == 'train' and not args.no_e2e and not args.no_trainer_api:
        try:
            eng._graphs.clear()
            trainer_api = measure_trainer_api(ctx)
        except Exception as exc:                 # pragma: no cover
            trainer_api = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    if rank != 0:
        if ctx.dist is not None:
            shutdown_distributed(ctx.dist, [eng])
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, threads, tried = cpu_arm(run.ds, run.comp, run.mean, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 steps of the batch-128 workload (cv2 augmentation in 8 worker processes + torch-CPU ResNet "
                         "fwd/bwd/ADAM) after 1 warm-up step; frames/s by torch thread count: %s; host cores "
                         "available: %d" % ({k: round(x, 1) for k, x in tried.items()}, cpu_threads())}
    h2d = nb * 128 * 128 * 4 + nb * run.rec_bytes + nb * E * 4
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": {0: "f32", 1: "tf32x3", 2: "tf32"}[run.precision], "data": "synthetic",
        "config": {"workload": "%s, %d resident crops" % (descr, N_RESIDENT),
                   "global_batch": nb * world, "per_gpu_batch": nb, "parallelism": "dp%d" % world,
                   "batchnorm": ("SyncBN (statistics summed over the ranks, %s)" % ('peer memory' if args.syncbn == 'p2p' else 'NCCL')) if (args.syncbn and world > 1) else
                                ("per-replica statistics" if world > 1 else "single device"),
                   "l2": "inputs (134 MB of crops + 0.9 GB of activations per step) exceed the 126 MB L2",
                   "precision_mode": run.precision},
        "e2e": None if args.no_e2e else {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                                         "d2h_bytes_per_step": 4, "note": e2e_note},
        "e2e_records_prebuilt": None if ms_pre is None else {"value": nb * world / (ms_pre / 1000.), "unit": UNIT,
                                                               "ms_per_step": ms_pre},
        "trainer_api": trainer_api,
        "cost_check": check,
        "strong": strong,
        "gpu_launches": n_launch * args.steps,
        "clocks": clk,
        "roofline": roof,
        "cpu_baseline": cpu,
        "tensor_fraction_whole_step": (value / world * TRAIN_GFLOP_PER_FRAME / 1000. / peaks()[1]) if args.workload != 'poseregnet' else None,
    }
    print(json.dumps(out))
    if check is not None and check.get('ok') is False:
        print("bench.py: the timed path's first-step cost differs from the oracle: %r" % (check,), file=sys.stderr)
        if ctx.dist is not None:
            shutdown_distributed(ctx.dist, [eng])
        sys.exit(3)
    if ctx.dist is not None:
        shutdown_distributed(ctx.dist, [eng])


def count_launches(eng):
    """kernels of OUR library launched per training step (checked against the ncu launch list)"""
    n = 1 + 2 + 1 + 1      # loss, adam + tick, weight-image pack, ema (the two arena fills are memset nodes)
    n += eng.launches_bn_bwd() if hasattr(eng, 'launches_bn_bwd') else len(eng.bns)
    n += eng.launches_wgrad() if hasattr(eng, 'launches_wgrad') else len([o for o in eng.ops if o['kind'] == 'conv'])
    for op in eng.ops:
        k = op['kind']
        if k == 'conv':
            n += 2                  # fwd, dgrad (backward-weights: counted above, grouped or per layer)
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
    precision 0: 126 of the launches and ~40 % of the GPU time, see profiles/).  All 63 ConvLayer forward
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
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get('conv_fwd_avg_bytes_per_launch')
            traffic_src = "COMMITTED figure, not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum per " \
                          "launch from the ncu pass recorded in profiles/ncu_traffic.json (%s)" % tj.get('source', 'see profiles/README.md')
        except Exception:
            traffic = None
    kname = {0: 'k_igemm', 1: 'k_conv_tc<*,2>', 2: 'k_conv_tc<*,1>'}[eng.precision]
    return {"bound": "hbm", "kernel": "%s: the %d ConvLayer forward launches of the step, back to back" % (kname, len(calls)),
            "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": traffic,
            "traffic_source": traffic_src,
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
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer arms (profiler runs)')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling sample of the default workload')
    ap.add_argument('--no-trainer-api', action='store_true', help='skip the PoseRegNetTrainer.train() measurement')
    ap.add_argument('--no-cost-check', action='store_true')
    ap.add_argument('--syncbn', nargs='?', const=True, default=False, choices=[True, 'nccl', 'p2p'],
                    help="N > 1: sum the BatchNorm statistics over the ranks (NCCL per BatchNorm, or 'p2p': peer-memory exchange)")
    ap.add_argument('--workload', default='train', choices=['train', 'icvl512', 'msra15', 'poseregnet', 'cascade'],
                    help="train = BASELINE configs[1] (the headline); icvl512 = configs[2]; msra15 = configs[3]; "
                         "cascade = configs[4], tools/bench_cascade.py")
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
