#!/usr/bin/env python
"""Hot source lines of an ncu report: joins the report's SASS page (stall samples per instruction) with the
line table of the same build (`nvdisasm -g -c` of the cubin compiled with -lineinfo), by instruction index.

usage: python tools/ncu_hot.py report.ncu-rep file.sass 'k_conv_tcILi64ELi2' [top]"""
import collections
import csv
import re
import subprocess
import sys


def sass_lines(sass_path, kernel_pat):
    """[(line_no, instruction text)] of the kernel whose mangled name contains kernel_pat"""
    out, on, cur = [], False, None
    for ln in open(sass_path, errors='replace'):
        if ln.startswith('//--------------------- .text.'):
            on = kernel_pat in ln
            cur = None
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            # inlined frames are listed innermost first; keep the OUTERMOST (last of a run) that is in conv code
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main():
    rep, sass, pat = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    H = rows[h]
    si, ii = H.index('# Samples'), H.index('Instructions Executed')
    stall_cols = [(i, c[6:]) for i, c in enumerate(H) if c.startswith('stall_') and 'Not Issued' not in c]
    inst = rows[h + 1:]
    lines = sass_lines(sass, pat)
    if len(lines) != len(inst):
        print("WARNING: instruction count mismatch: report %d vs disassembly %d" % (len(inst), len(lines)))
    agg = collections.defaultdict(lambda: [0, collections.Counter(), 0])
    total = 0
    for k, r in enumerate(inst):
        n = int(r[si] or 0)
        total += n
        key = lines[k][0] if k < len(lines) else None
        a = agg[key]
        a[0] += n
        a[2] += int(r[ii] or 0)
        for i, c in stall_cols:
            v = int(r[i] or 0)
            if v:
                a[1][c] += v
    src = {}
    print("total samples %d, %d instructions" % (total, len(inst)))
    for key, (n, st, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ''
        if key:
            f, l = key
            if f not in src:
                try:
                    src[f] = open('deep-prior-pp_b200/csrc/' + f).read().split('\n') if '--src' not in sys.argv else []
                except OSError:
                    src[f] = []
            alt = sys.argv[sys.argv.index('--srcdir') + 1] if '--srcdir' in sys.argv else None
            if alt:
                try:
                    src[f] = open(alt + '/' + f).read().split('\n')
                except OSError:
                    pass
            text = src[f][l - 1].strip()[:100] if l - 1 < len(src[f]) else ''
        print("%6d %5.1f%% inst=%-8d %-22s %-100s | %s" % (n, 100.0 * n / max(total, 1), ex, '%s:%d' % key if key else '?', text,
                                                    ' '.join('%s:%d' % kv for kv in st.most_common(3))))


if __name__ == '__main__':
    main()
