"""The per-thread core of the opt-in 8-filter conv+pool kernels (deep-prior-pp_b200/csrc/convpool8.cuh, shared between
the device kernel and the host) executed on the CPU: tests/convpool8_host_test.cpp emulates the kernel's tile loop
and compares outputs and arg-max cells bit for bit with an independent direct convolution + max-pool, for every
(filter, channels, pool) combination of the ScaleNet / PoseRegNet towers.  The device kernel itself is compared with
the generic kernel on the GPU by tests/test_gpu_convpool8.py."""
import os
import shutil
import subprocess

import pytest


@pytest.mark.skipif(shutil.which('g++') is None, reason="needs g++")
def test_convpool8_thread_code_on_host(tmp_path):
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(str(tmp_path), 'convpool8_host_test')
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-ffp-contract=off', os.path.join(here, 'convpool8_host_test.cpp'),
                           '-o', exe, '-lm'])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'convpool8 host test OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
