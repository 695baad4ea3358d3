#!/usr/bin/env python
"""Kernel probe (development tool, GPU only): times dpp_conv2d_fwd / dgrad / wgrad on the ResNet's
batch-128 layer shapes in isolation (CUDA events, warm and L2-flushed) and, with the -DDPP_PROFILE
build (make -C deep-prior-pp_b200/csrc prof; DPP_LIB=.../libdpp_b200_prof.so), prints the in-kernel
clock64 timeline of CTA 0 (producer / MMA / epilogue / loader events)."""
import ctypes as C
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-prior-pp_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from dpp_b200.lib import lib, ConvDesc, BnRef, PackItem  # noqa: E402
from test_gpu_conv_tc import _setup, P  # noqa: E402

SHAPES = {  # N, H, Cin, Cout, k, stride, residual
    'A_3x3_16_16@32': (128, 32, 16, 16, 3, 1, 0),
    'B_1x1_16_64@32+res': (128, 32, 16, 64, 1, 1, 1),
    'C_1x1_64_16@32': (128, 32, 64, 16, 1, 1, 0),
    'D_1x1_256_64@8': (128, 8, 256, 64, 1, 1, 0),
    'E_3x3_64_64@8': (128, 8, 64, 64, 3, 1, 0),
    'F_1x1_64_256@8+res': (128, 8, 64, 256, 1, 1, 1),
    'G_1x1s2_32_16@64': (128, 64, 32, 16, 1, 2, 0),
    'H_3x3_32_32@16': (128, 16, 32, 32, 3, 1, 0),
    'Z_tiny': (1, 8, 64, 64, 1, 1, 0),
}
TAGS = {1: 'entry', 2: 'setup done', 3: 'role done', 4: 'exit', 10: 'P round start', 11: 'P chunk: loads->wait stage',
        12: 'P stage free', 13: 'P arrived', 10: 'P issued', 11: 'P data landed', 20: 'M acc free', 21: 'M B ready', 22: 'M A ready', 23: 'M committed',
        30: 'E side issued', 31: 'E acc ready', 32: 'E tile done',
        14: 'P stored', 15: 'P fenced', 33: 'E tmem loaded', 34: 'E staged', 35: 'E batch stored', 36: 'E batch stats done'}


def timeline(prof, maxev=60):
    ev = []
    for base, role in ((0, 'P'), (1000, 'M'), (2000, 'E'), (3000, 'L')):
        i = base
        while i < base + 990 and prof[i] != 0:
            ev.append((int(prof[i + 1]), role, int(prof[i])))
            i += 2
    if not ev:
        return
    ev.sort()
    t0 = ev[0][0]
    print("    cycles  role event")
    for t, role, tag in ev[:maxev]:
        print("    %7d  %s   %s" % (t - t0, role, TAGS.get(tag, tag)))
    if len(ev) > maxev:
        print("    ... %d more events; last at %d" % (len(ev) - maxev, ev[-1][0] - t0))


def main():
    precision = int(os.environ.get('DPP_PRECISION', '1'))
    which = sys.argv[1:] or list(SHAPES)
    has_prof = False
    try:
        setp = lib.raw('dpp_debug_set_prof')
        setp.restype = C.c_int
        setp.argtypes = [C.c_void_p]
        has_prof = True
    except AttributeError:
        pass
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
    if os.environ.get('PROBE_FLAGS'):
        setf = lib.raw('dpp_debug_set_flags')
        setf.restype = C.c_int
        setf.argtypes = [C.c_int]
        setf(int(os.environ['PROBE_FLAGS']))
        print("debug flags", os.environ['PROBE_FLAGS'])
        has_prof = has_prof and os.environ.get('PROBE_TIMELINE', '0') == '1'
    for name in which:
        N, H, Cin, Cout, k, stride, res = SHAPES[name]
        d, x, w, bias, bn, Ho, keep = _setup(N, H, Cin, Cout, k, stride, precision)
        y = torch.empty(N, Ho, Ho, Cout, device='cuda')
        r = torch.randn(N, Ho, Ho, Cout, device='cuda') if res else None
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device='cuda')

        def launch():
            lib.dpp_conv2d_fwd(C.byref(d), P(x), C.byref(bn), P(w), P(bias), P(r), P(y), P(stats), None)
        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            launch()
        e1.record(); torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / 20 * 1e3
        # the same 20 launches replayed from a CUDA graph: no host launch cost between kernels
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        sp = C.c_void_p(side.cuda_stream)

        def launch_on(stp):
            lib.dpp_conv2d_fwd(C.byref(d), P(x), C.byref(bn), P(w), P(bias), P(r), P(y), P(stats), stp)
        with torch.cuda.stream(side):
            launch_on(sp)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(20):
                launch_on(C.c_void_p(torch.cuda.current_stream().cuda_stream))
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        graph_us = e0.elapsed_time(e1) / 20 * 1e3
        cold = 0.0
        for _ in range(5):
            flush.fill_(1.0)
            e0.record(); launch(); e1.record(); torch.cuda.synchronize()
            cold += e0.elapsed_time(e1) / 5 * 1e3
        byt = 4.0 * (x.numel() + y.numel() * (2 if res else 1))
        print("%-22s fwd warm %7.1f us  graph %7.1f us  cold %7.1f us   alg bytes %6.1f MB -> %5.0f GB/s (graph)" % (
            name, warm, graph_us, cold, byt / 1e6, byt / graph_us / 1e3))
        if os.environ.get('PROBE_BWD', '1') == '1':
            dy = torch.randn(N, Ho, Ho, Cout, device='cuda')
            dx = torch.zeros(N, H, H, Cin, device='cuda')
            dzs = torch.zeros(2 * Cin, dtype=torch.float64, device='cuda')
            dw = torch.zeros(k * k * Cin, Cout, device='cuda')
            db = torch.zeros(Cout, device='cuda')

            def graph_time(fn):
                if os.environ.get('PROBE_EAGER', '0') == '1':       # profiler runs: no graph replays
                    for _ in range(3):
                        fn(None)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(20):
                        fn(None)
                    e1.record(); torch.cuda.synchronize()
                    return e0.elapsed_time(e1) / 20 * 1e3
                with torch.cuda.stream(side):
                    fn(sp)
                torch.cuda.synchronize()
                gg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gg, stream=side):
                    for _ in range(20):
                        fn(C.c_void_p(torch.cuda.current_stream().cuda_stream))
                gg.replay(); torch.cuda.synchronize()
                e0.record(); gg.replay(); e1.record(); torch.cuda.synchronize()
                return e0.elapsed_time(e1) / 20 * 1e3
            t_dg = graph_time(lambda stp: lib.dpp_conv2d_dgrad(C.byref(d), P(dy), P(w), P(dx), 0, C.byref(bn), P(x), P(dzs), stp))
            t_wg = graph_time(lambda stp: lib.dpp_conv2d_wgrad(C.byref(d), P(x), C.byref(bn), P(dy), P(dw), P(db), stp))
            bd = 4.0 * (dy.numel() + 2 * x.numel())
            bw = 4.0 * (dy.numel() + x.numel())
            if os.environ.get('PROBE_WG_TIMELINE', '0') == '1':
                try:
                    setw = lib.raw('dpp_debug_set_prof_wg')
                    setw.restype = C.c_int
                    setw.argtypes = [C.c_void_p]
                    prof = torch.zeros(5000, dtype=torch.int64, device='cuda')
                    setw(prof.data_ptr())
                    lib.dpp_conv2d_wgrad(C.byref(d), P(x), C.byref(bn), P(dy), P(dw), P(db), None)
                    torch.cuda.synchronize()
                    setw(None)
                    print("  wgrad timeline (CTA 0; P = producer warp 0, M = MMA warp, E = warp 4, L = warp 7)")
                    timeline(prof.cpu().numpy(), int(os.environ.get('PROBE_EVENTS', '90')))
                except AttributeError:
                    print("  (no dpp_debug_set_prof_wg in this build)")
            print("%-22s dgrad graph %7.1f us (%5.0f GB/s)   wgrad graph %7.1f us (%5.0f GB/s)" % (
                name, t_dg, bd / t_dg / 1e3, t_wg, bw / t_wg / 1e3))
        if has_prof:
            prof = torch.zeros(5000, dtype=torch.int64, device='cuda')
            setp(prof.data_ptr())
            launch()
            torch.cuda.synchronize()
            setp(None)
            timeline(prof.cpu().numpy(), int(os.environ.get('PROBE_EVENTS', '70')))


if __name__ == '__main__':
    main()
