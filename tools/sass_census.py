#!/usr/bin/env python
"""Per-kernel SASS opcode census of libdpp_b200.so (runs here, no GPU): the Blackwell-native instructions each kernel
contains - UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTMALDG (cp.async.bulk.tensor,
TMA tensor load), UBLKCP (cp.async.bulk), LDGSTS (cp.async), SYNCS (mbarrier), REDG / RED (red.global) - beside HMMA
(legacy mma.sync, must be 0).  usage: python tools/sass_census.py > profiles/r2_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'deep-prior-pp_b200', 'csrc', 'libdpp_b200.so')
OPS = ['UTCHMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'RED', 'HMMA', 'FFMA2', 'FFMA', 'DFMA']


def demangle(name):
    try:
        return subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    counts, cur, lines = collections.OrderedDict(), None, collections.Counter()
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            lines[cur] += 1
            for o in OPS:
                if op == o or op.startswith(o + '.') or (o == 'RED' and op.startswith('REDG')):      # (the regex stops 'FFMA2' from counting as 'FFMA')
                    counts[cur][o] += 1
    print("SASS opcode census of %s (sm_100a), one row per kernel" % os.path.relpath(LIB, ROOT))
    print("%-58s %7s " % ("kernel", "instrs") + " ".join("%7s" % o for o in OPS))
    tot = collections.Counter()
    for fn, c in counts.items():
        d = demangle(fn)
        d = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', d)
        d = re.sub(r'\(.*$', '', d).replace('void ', '')
        print("%-58s %7d " % (d[:58], lines[fn]) + " ".join("%7d" % c[o] for o in OPS))
        tot.update(c)
    print("%-58s %7d " % ("TOTAL", sum(lines.values())) + " ".join("%7d" % tot[o] for o in OPS))


if __name__ == '__main__':
    main()
