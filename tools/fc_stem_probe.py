#!/usr/bin/env python
"""Kernel probe (development tool, GPU only): times the ResNet stem (dpp_convpool_fwd / _bwd, 5x5 1->32 + pool 2)
and the HiddenLayer entry points (dpp_fc_fwd / dpp_fc_bwd) at the benchmarked batch-128 shapes, CUDA events over
20 back-to-back calls with a 256 MB L2 flush write between rounds."""
import ctypes as C
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-prior-pp_b200'))
from dpp_b200.lib import lib  # noqa: E402


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def timed(fn, reps=20, rounds=3):
    """us per call; PROBE_COLD=1: a 256 MB write before EVERY call (weights come from HBM, as inside a training step)"""
    flush = torch.empty(64 * 1024 * 1024, device='cuda')
    if os.environ.get('PROBE_COLD', '0') == '1':
        tot = []
        for _ in range(reps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot.append(e0.elapsed_time(e1) * 1e3)
        tot.sort()
        return tot[len(tot) // 2]
    best = 1e9
    for _ in range(rounds):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def main():
    B = int(os.environ.get('PROBE_B', '128'))
    precision = int(os.environ.get('DPP_PRECISION', '1'))
    g = torch.Generator(device='cuda').manual_seed(1)
    # ---- stem
    x = torch.randn(B, 128, 128, 1, device='cuda', generator=g)
    w = torch.randn(25, 32, device='cuda', generator=g) * 0.2
    b = torch.randn(32, device='cuda', generator=g) * 0.1
    y = torch.zeros(B, 64, 64, 32, device='cuda')
    am = torch.zeros(B, 64, 64, 32, dtype=torch.uint8, device='cuda')
    stats = torch.zeros(64, dtype=torch.float64, device='cuda')
    dy = torch.randn(B, 64, 64, 32, device='cuda', generator=g)
    dw = torch.zeros(25, 32, device='cuda')
    db = torch.zeros(32, device='cuda')
    t = timed(lambda: lib.dpp_convpool_fwd(P(x), P(w), P(b), P(y), P(am), P(stats), B, 128, 128, 1, 32, 5, 2, 2, 0, None))
    print("stem fwd   %7.1f us  (%.1f TFLOP/s fp32)" % (t, 2.0 * B * 128 * 128 * 25 * 32 / t * 1e-6))
    t = timed(lambda: lib.dpp_convpool_bwd(P(x), P(w), P(y), P(am), P(dy), P(dw), P(db), None, B, 128, 128, 1, 32, 5, 2, 2, 0, None))
    print("stem bwd_w %7.1f us" % t)
    # ---- FC layers of the ResNet
    for (n_in, n_out, relu) in ((16384, 1024, 1), (1024, 1024, 1), (1024, 30, 0)):
        xx = torch.randn(B, n_in, device='cuda', generator=g)
        ww = torch.randn(n_in, n_out, device='cuda', generator=g) * (1.0 / n_in) ** 0.5
        bb = torch.randn(n_out, device='cuda', generator=g) * 0.1
        yy = torch.zeros(B, n_out, device='cuda')
        go = torch.randn(B, n_out, device='cuda', generator=g)
        dww = torch.zeros(n_in, n_out, device='cuda')
        dbb = torch.zeros(n_out, device='cuda')
        dxx = torch.zeros(B, n_in, device='cuda')
        scratch = torch.zeros(B, n_out, device='cuda')
        tf = timed(lambda: lib.dpp_fc_fwd(P(xx), P(ww), P(bb), P(yy), B, n_in, n_out, relu, None, 1.0, precision, None))
        tb = timed(lambda: lib.dpp_fc_bwd_ex(P(xx), P(ww), P(yy), P(go), P(dww), P(dbb), P(dxx), P(scratch), B, n_in, n_out, relu,
                                             None, 1.0, precision, 1, None))
        tw = timed(lambda: lib.dpp_fc_bwd_ex(P(xx), P(ww), P(yy), P(go), P(dww), P(dbb), None, P(scratch), B, n_in, n_out, relu,
                                             None, 1.0, precision, 1, None))
        mb = n_in * n_out * 4e-6
        print("fc %5d -> %4d: fwd %7.1f us (%.0f GB/s of weights)   bwd (pre + dW + dx) %7.1f us (%.0f GB/s over 2 weight-sized streams)"
              "   pre + dW %7.1f us" % (n_in, n_out, tf, mb / tf * 1e3, tb, 2 * mb / tb * 1e3, tw))


TAGS = {1: 'entry', 2: 'setup done', 3: 'role done', 10: 'T slot free', 20: 'landed', 21: 'stage free', 22: 'stored + arrived',
        52: 'E ld16 done', 53: 'E 16 stores issued', 40: 'M acc free', 41: 'M A full', 42: 'M B full', 43: 'M committed', 50: 'E acc full', 51: 'E done'}


def timeline_fs(B, n_in, n_out, which, maxev=120):
    """in-kernel timeline of CTA 0 of one k_fc_stream launch (prof build: make -C deep-prior-pp_b200/csrc prof)"""
    try:
        setp = lib.raw('dpp_debug_set_prof_fs')
    except Exception:
        print("(no dpp_debug_set_prof_fs in this build)")
        return
    setp.restype = C.c_int
    setp.argtypes = [C.c_void_p]
    g = torch.Generator(device='cuda').manual_seed(1)
    xx = torch.randn(B, n_in, device='cuda', generator=g)
    ww = torch.randn(n_in, n_out, device='cuda', generator=g)
    bb = torch.zeros(n_out, device='cuda')
    yy = torch.zeros(B, n_out, device='cuda')
    go = torch.randn(B, n_out, device='cuda', generator=g)
    dww = torch.zeros(n_in, n_out, device='cuda')
    dbb = torch.zeros(n_out, device='cuda')
    dxx = torch.zeros(B, n_in, device='cuda')
    scratch = torch.zeros(B, n_out, device='cuda')
    prof = torch.zeros(5000, dtype=torch.int64, device='cuda')
    for _ in range(2):      # second run: warm instruction cache
        prof.zero_()
        torch.cuda.synchronize()
        if which == 'fwd':
            setp(prof.data_ptr())
            lib.dpp_fc_fwd(P(xx), P(ww), P(bb), P(yy), B, n_in, n_out, 1, None, 1.0, 1, None)
        else:
            setp(prof.data_ptr())
            lib.dpp_fc_bwd_ex(P(xx), P(ww), P(yy), P(go), P(dww), P(dbb), P(dxx) if which == 'dx' else None, P(scratch), B, n_in,
                              n_out, 1, None, 1.0, 1, 1, None)
        torch.cuda.synchronize()
    setp(None)
    pr = prof.cpu().numpy()
    ev = []
    for base, role in ((0, 'A'), (1000, 'B'), (2000, 'M'), (3000, 'E'), (4000, 'T')):
        i = base
        while i < base + 990 and pr[i] != 0:
            ev.append((int(pr[i + 1]), role, int(pr[i])))
            i += 2
    ev.sort()
    print("== timeline %s %d x %d -> %d (CTA 0; with dx the buffer holds the LAST launch = dx)" % (which, B, n_in, n_out))
    t0 = ev[0][0]
    only = os.environ.get('PROBE_ROLES')
    if only:
        ev = [e for e in ev if e[1] in only]
    for t, role, tag in ev[:maxev]:
        print("  %7d  %s  %s" % (t - t0, role, TAGS.get(tag, tag)))
    print("  ... %d events, last at %d" % (len(ev), ev[-1][0] - t0))


if __name__ == '__main__':
    if os.environ.get('PROBE_TIMELINE', '0') == '1':
        for which in os.environ.get('PROBE_WHICH', 'fwd,dw,dx').split(','):
            timeline_fs(128, 16384, 1024, which, int(os.environ.get('PROBE_EVENTS', '120')))
        sys.exit(0)
    main()
