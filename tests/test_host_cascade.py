"""Host logic of the inference cascade on the CPU: tests/cascade_dryrun.py runs the bodies of the GPU cascade tests
with the device faked (kernel emulated by tests/recrop_model.py, engines by the oracle nets) in a SUBPROCESS, because
it monkey-patches torch.cuda.  It validates records, batch padding, pointer hand-off, the RealtimeHandposePipeline and
HandDetector surfaces; the CUDA kernels are validated by tests/test_gpu_cascade.py on the GPU."""
import os
import subprocess
import sys


def test_cascade_host_plumbing_dry_run():
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, 'cascade_dryrun.py')], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert 'cascade ok' in r.stdout and 'joint errors ok' in r.stdout and 'recrop MSRA15 ok' in r.stdout
    assert 'evaluation ok' in r.stdout and 'poses ok' in r.stdout and 'dataset ok' in r.stdout
