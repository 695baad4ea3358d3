#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel.

usage: python tools/launch_summary.py launches.csv [--steps N] [--step n] [--md]
Prints, per kernel name (template arguments kept, parameter list dropped): launches, total us, average us,
share of the summed GPU time.  With --steps the totals are divided by the number of profiled steps.
"""
import collections
import csv
import re
import sys


def load(path, metric='gpu__time_duration.sum'):
    rows = list(csv.reader(open(path, errors='replace')))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    H = rows[h]
    ki, vi, gi, bi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size'), H.index('Block Size')
    mi = H.index('Metric Name')
    out = []
    for r in rows[h + 1:]:
        if len(r) <= vi or r[mi] != metric:
            continue
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        name = re.sub(r'^(void )?(<unnamed>|\(anonymous namespace\))::', '', r[ki])
        name = re.sub(r'\((?:[^()]|\([^()]*\))*\)\s*$', '', name)
        out.append((name, v, r[gi], r[bi]))
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--') and not a.isdigit()]
    steps = 1
    if '--steps' in sys.argv:
        steps = int(sys.argv[sys.argv.index('--steps') + 1])
        args = [a for a in args if a != str(steps)]
    md = '--md' in sys.argv
    L = load(args[0])
    if '--step' in sys.argv:
        # one steady-state step: the launches between the n-th and (n+1)-th k_adam_tick (a rotation of one step)
        n = int(sys.argv[sys.argv.index('--step') + 1])
        ticks = [i for i, l in enumerate(L) if l[0].startswith('k_adam_tick')]
        L = L[ticks[n] + 1:ticks[n + 1] + 1]
    agg = collections.OrderedDict()
    for name, v, g, b in L:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    items = sorted(agg.items(), key=lambda kv: -kv[1][1])
    if md:
        print('| kernel | launches/step | us/step | avg us | share |')
        print('|---|---|---|---|---|')
    for k, (n, t) in items:
        if md:
            print('| `%s` | %.1f | %.1f | %.2f | %.3f |' % (k[:90], n / steps, t / 1e3 / steps, t / n / 1e3, t / tot))
        else:
            print('%-90s n=%7.1f us=%10.1f avg_us=%8.2f share=%.3f' % (k[:90], n / steps, t / 1e3 / steps, t / n / 1e3, t / tot))
    print(('| **total** | %.1f | %.1f | | |' if md else 'TOTAL n=%.1f us=%.1f') % (len(L) / steps, tot / 1e3 / steps))


if __name__ == '__main__':
    main()
