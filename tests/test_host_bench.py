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


def test_e2e_prep_workers_return_what_the_main_thread_would():
    """bench.py's e2e host preparation runs in spawned worker processes (_prep_init / _prep_job): a worker's batch is
    the batch records_for builds in-process from the same seed, and the records index the staged crops 0..n-1."""
    code = r"""
import sys, numpy as np, multiprocessing
sys.path.insert(0, %r)
import bench
if __name__ == '__main__':
    pool = multiprocessing.get_context('spawn').Pool(2, initializer=bench._prep_init, initargs=('NYU', 64, 23455, ['com', 'rot', 'none']))
    got = pool.map(bench._prep_job, [(11, 16), (12, 16)])
    # the same job writing into a slot of the shared-memory staging ring: crops | records | labels
    from multiprocessing import shared_memory
    xb, rb, yb = 16 * 128 * 128 * 4, 16 * 112, 16 * 30 * 4
    slot_bytes = (xb + rb + yb + 4095) // 4096 * 4096
    shm = shared_memory.SharedMemory(create=True, size=3 * slot_bytes)
    assert pool.apply(bench._prep_job, ((11, 16, shm.name, 2, slot_bytes),)) == 2
    pool.terminate(); pool.join()
    ds, comp, mean = bench.make_workload(seed=23455, dataset='NYU', n=64)
    o = 2 * slot_bytes
    x = np.ndarray((16, 128, 128), np.float32, buffer=shm.buf, offset=o)
    assert (x == ds['x'][got[0][0], 0]).all()
    assert (np.ndarray((16, 112), np.uint8, buffer=shm.buf, offset=o + xb) == got[0][1]).all()
    assert (np.ndarray((16, 30), np.float32, buffer=shm.buf, offset=o + xb + rb) == got[0][2]).all()
    del x
    shm.close(); shm.unlink()
    for seed, (src, rb, yv) in zip((11, 12), got):
        rng = np.random.RandomState(seed)
        idxs = rng.randint(0, 64, 16)
        r, y, _ = bench.records_for(ds, comp, mean, idxs, rng, ['com', 'rot', 'none'])
        assert (src == r['src_index']).all() and (yv == y).all()
        r = r.copy(); r['src_index'] = np.arange(16, dtype=np.int32)
        assert (rb == r.view(np.uint8).reshape(16, -1)).all()
    print('workers ok')
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'workers ok' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
