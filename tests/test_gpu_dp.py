"""Data-parallel parity on >= 2 GPUs (skipped on a single-GPU box): tools/dp_check.py under torchrun - the exchanged
gradient equals the sum of the per-rank gradients, bucketed (early) and single (late) exchange agree, replicas stay
bit-identical, and with SyncBN two ranks x B samples reproduce ONE device computing the global batch of 2B (the
single-device reference semantics of net/batchnormlayer.py:154-159)."""
import os
import subprocess
import sys
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dp_check_two_ranks():
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, 'deep-prior-pp_b200'))
    p = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29581', os.path.join(ROOT, 'tools', 'dp_check.py')],
                       capture_output=True, text=True, timeout=600, env=env)
    print(p.stdout[-3000:])
    print(p.stderr[-2000:])
    assert 'DP_CHECK PASS' in p.stdout
