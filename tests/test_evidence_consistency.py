"""The committed measurement evidence is self-consistent: the roofline's ``traffic`` figure (profiles/ncu_traffic.json,
quoted by bench.py) is what tools/traffic_from_launches.py extracts from the committed ncu launch list, the launch list
holds the kernels DESIGN.md names, and every committed bench line carries the keys of the driver's contract."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_roofline_traffic_matches_the_committed_launch_list():
    want = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'traffic_from_launches.py'),
                        os.path.join(ROOT, want['source']), str(want.get('step', 0))], capture_output=True, text=True, cwd=ROOT,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads(r.stdout)
    assert got['launches'] == want['launches'] == 63
    assert abs(got['conv_fwd_avg_bytes_per_launch'] - want['conv_fwd_avg_bytes_per_launch']) < 1.0
    bench = json.load(open(os.path.join(ROOT, 'profiles', 'r2b_bench.json')))
    assert abs(bench['roofline']['traffic'] - want['conv_fwd_avg_bytes_per_launch']) < 0.01 * want['conv_fwd_avg_bytes_per_launch']
    assert abs(bench['roofline']['frac'] - bench['roofline']['achieved'] / bench['roofline']['peak']) < 1e-9


def test_launch_list_names_the_kernels_of_the_step():
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    from launch_summary import load
    names = set(n.split('<')[0] for n, _, _, _ in load(os.path.join(ROOT, 'profiles', 'r1_final_launches.csv')))
    for k in ('k_conv_tc', 'k_wgrad_mn', 'k_bn_bwd_apply', 'k_stem_fwd', 'k_gemm_tc', 'k_adam', 'k_augment', 'k_loss_sqerr'):
        assert k in names, k
    rows = load(os.path.join(ROOT, 'profiles', 'r2_final_launches.csv'))
    names2 = set(n.split('<')[0] for n, _, _, _ in rows)
    for k in ('k_conv_tc', 'k_wgrad_group', 'k_bn_bwd_apply', 'k_stem_fwd', 'k_gemm_tc', 'k_adam', 'k_augment', 'k_loss_sqerr'):
        assert k in names2, k
    assert 'k_wgrad_mn' not in names2                  # round 2: the backward-weights GEMMs run grouped
    # ... and between two k_augment launches (one step) only the 4 un-fusable BatchNorm-backward launches are left
    aug = [i for i, r in enumerate(rows) if r[0].startswith('k_augment')]
    step = [r[0].split('<')[0] for r in rows[aug[0]:aug[1]]]
    assert step.count('k_bn_bwd_apply') == 4 and step.count('k_wgrad_group') == 4 and step.count('k_conv_tc') == 126
    # second half of round 2: streaming HiddenLayer GEMMs, fp32 backward-weights launches for the narrow layers
    rows3 = load(os.path.join(ROOT, 'profiles', 'r2b_final_launches.csv'))
    aug = [i for i, r in enumerate(rows3) if r[0].startswith('k_augment')]
    step = [r[0].split('<')[0] for r in rows3[aug[-2]:aug[-1]]]
    assert step.count('k_conv_tc') == 126 and step.count('k_wgrad_group') == 4 and step.count('k_wgrad3') == 2
    assert step.count('k_wgrad1') == 2 and step.count('k_fc_stream') == 6 and step.count('k_gemm_tc') == 0
    assert step.count('k_stem_fwd') == 1 and step.count('k_stem_bwd_w') == 1 and step.count('k_adam') == 1
    casc = set(n.split('<')[0] for n, _, _, _ in load(os.path.join(ROOT, 'profiles', 'r1_cascade_launches.csv')))
    assert {'k_recrop', 'k_convpool_fwd', 'k_conv_tc', 'k_stem_fwd'} <= casc


def test_committed_bench_lines_follow_the_contract():
    keys = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
            'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches')
    for path in glob.glob(os.path.join(ROOT, 'profiles', 'r1_*bench*.json')) + glob.glob(os.path.join(ROOT, 'profiles', 'r2_*bench*.json*')) + \
            glob.glob(os.path.join(ROOT, 'profiles', 'r2b_*bench*.json*')):
        for line in open(path).read().strip().splitlines():
            d = json.loads(line)
            for k in keys:
                assert k in d, (os.path.basename(path), k)
            assert 'workload' in d['config']
            if d['e2e'] is not None:                 # profiler / multi-variant runs skip the host-buffer arm (--no-e2e)
                assert d['e2e']['value'] > 0
            if d.get('impl') != 'reference':
                assert d['gpu_launches'] > 0
                if d.get('roofline'):            # multi-GPU and early lines were taken with --no-roofline
                    assert d['roofline']['bound'] in ('hbm', 'tensor') and 0 < d['roofline']['frac'] < 1
                if d.get('clocks'):
                    assert set(d['clocks']['reasons']) <= {'sw_power_cap'}
