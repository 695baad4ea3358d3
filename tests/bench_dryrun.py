"""bench.py's B200 arm walked on the CPU (run by tests/test_host_bench.py in a subprocess): the device is faked as in
tests/cascade_dryrun.py (kernels -> NumPy transcriptions, engine -> the oracle net), so every Python statement of
``run_b200`` executes - workload and record preparation, the resident / e2e / e2e-with-host-preparation loops, the JSON
line.  TEST INFRASTRUCTURE ONLY: it guards the driver's measurement script against Python-level regressions; the
numbers it prints mean nothing."""
import contextlib
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(HERE, 'cascade_dryrun.py')).read()
src = src[:src.index("import test_gpu_cascade as T")]          # the fakes only, not the cascade test bodies
exec(compile(src, os.path.join(HERE, 'cascade_dryrun.py'), 'exec'))


class _Event(object):
    def __init__(self, **k):
        self.t = 0.

    def record(self, *a):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


class _Stream(object):
    cuda_stream = 0

    def __init__(self, **k):
        pass

    def wait_event(self, e):
        pass


torch.cuda.Event = _Event                                      # noqa: F821 (torch comes from the exec'd fakes)
torch.cuda.synchronize = lambda *a, **k: None                  # noqa: F821
torch.Tensor.pin_memory = lambda self: self                    # noqa: F821
torch.cuda.Stream = _Stream                                    # noqa: F821
torch.cuda.current_stream = lambda *a, **k: _Stream()          # noqa: F821
torch.cuda.stream = lambda s: contextlib.nullcontext()         # noqa: F821

import dpp_b200.engine as ENG                                  # noqa: E402


class BenchEngine(FakeEngine):                                 # noqa: F821
    def __init__(self, net, precision=None, **kw):
        FakeEngine.__init__(self, net)                         # noqa: F821
        self.precision, self.bns, self._graphs, self._lr = precision, [], {}, 1e-4

    def set_lr(self, lr):
        self._lr = lr

    def set_world(self, *a, **k):
        pass

    def train_step(self, lr=None, use_graph=True):
        return FakeEngine.train_step(self, self._lr if lr is None else lr, use_graph)      # noqa: F821


ENG.Engine = BenchEngine
sys.argv = ['bench.py', '--steps', '1', '--warmup', '1', '--no-roofline', '--no-cpu-baseline', '--no-trainer-api']
sys.path.insert(0, os.path.dirname(HERE))
os.environ['DPP_PREP_WORKERS'] = '0'       # this script has no __main__ guard: spawned workers would re-run it (the pool has its own test)
import bench                                                   # noqa: E402

bench.B, bench.N_RESIDENT, bench.STRONG_GLOBAL_B = 8, 32, 8                              # a batch the CPU oracle steps through in a second


class _NoClocks(object):
    def __init__(self, i):
        pass

    def start(self):
        pass

    def stop(self):
        return {}


bench.ClockSampler = _NoClocks
bench.count_launches = lambda eng: 1
bench.main()

# ---- the cascade bench (tools/bench_cascade.py, BASELINE config 5) the same way ---------------------------------------
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tools'))
import bench_cascade                                           # noqa: E402

bench_cascade.main(['--batch', '8', '--steps', '2', '--warmup', '1', '--cpu-frames', '2'])
