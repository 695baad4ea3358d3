"""bench.py's measurement script walked on the CPU with the device faked (tests/bench_dryrun.py, in a subprocess because
it monkey-patches torch.cuda): the JSON line must carry every key of the driver's contract, and the optional
host-preparation ``e2e`` arm and the strong-scaling sample must have run without raising."""
import json
import os
import subprocess
import sys


def test_bench_b200_arm_walks_on_the_cpu():
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, 'bench_dryrun.py')], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('{"metric": "training')][-1]
    out = json.loads(line)
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
        assert k in out, k
    assert out['config']['workload'] and out['e2e']['h2d_bytes_per_step'] > 0 and out['e2e']['d2h_bytes_per_step'] == 4
    assert out['e2e']['value'] > 0 and 'host thread' in out['e2e']['note']      # e2e = host preparation inside the timed region
    assert out['e2e_records_prebuilt']['value'] > 0
    assert out['strong'] is not None and 'error' not in out['strong'] and out['strong']['scaling'] == 'strong'
    assert 'skipped' in out['cost_check']                                          # golden value is for batch 128 / 2048 crops
    casc = json.loads([l for l in r.stdout.splitlines() if l.startswith('{"metric": "cascade')][-1])
    for k in ('value', 'p50_ms', 'p90_ms', 'e2e', 'roofline', 'cpu_baseline', 'gpu_launches', 'config'):
        assert k in casc, k
    assert casc['roofline']['bound'] == 'hbm' and casc['e2e']['h2d_bytes_per_step'] > 8 * 480 * 640 * 4
