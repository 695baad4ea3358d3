#!/usr/bin/env python
"""Digest of an `ncu --set full` report: the handful of raw metrics DESIGN.md / bench.py quote
(duration, DRAM bytes, L2 hit rate, issue utilisation, shared-memory wavefronts / bank conflicts, occupancy).
usage: python tools/ncu_digest.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_tc.sum', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__grid_size', 'launch__block_size']


def main():
    rep = sys.argv[1]
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        print("kernel:", v[h.index('Kernel Name')])
        for w in WANT:
            if w in h:
                i = h.index(w)
                print("  %-62s %s %s" % (w, v[i], u[i]))
        tc = [(c, v[i]) for i, c in enumerate(h) if ('tensor' in c or 'pipe_tc' in c or 'tmem' in c) and v[i] not in ('', '0')]
        for c, x in tc[:12]:
            print("  %-62s %s" % (c, x))


if __name__ == '__main__':
    main()
