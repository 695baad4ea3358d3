"""Host logic of the inference cascade on the CPU: tests/cascade_dryrun.py runs the bodies of the GPU cascade tests
with the device faked (kernel emulated by tests/recrop_model.py, engines by the oracle nets) in a SUBPROCESS, because
it monkey-patches torch.cuda.  It validates records, batch padding, pointer hand-off, the RealtimeHandposePipeline and
HandDetector surfaces; the CUDA kernels are validated by tests/test_gpu_cascade.py on the GPU.  Where /root/reference
exists it also EXECUTES THE REFERENCE'S OWN ENTRY SCRIPTS (main_nyu / main_icvl_posereg_embedding.py, py2 -> py3 pass in
memory) against the product package: data preparation, PCA on 1e6 sampled poses, network / trainer set-up, train()
(two epochs, oracle arithmetic behind the emulated device), save, PCA prior layer, computeOutput and the evaluation
metrics all run through the product's classes unchanged."""
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
    # in the build container the reference's own entry scripts run against the product package (host side)
    assert 'entry scripts ok' in r.stdout or 'entry scripts skipped' in r.stdout
