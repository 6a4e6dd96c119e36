#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_me > x.csv
       python tools/ncu_lines.py x.csv [top]"""
import csv
import sys
from collections import defaultdict


SORT = 2 if "--by-samples" in sys.argv else 0
if "--by-samples" in sys.argv: sys.argv.remove("--by-samples")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    fname, hdr = '?', None
    agg = defaultdict(lambda: [0, 0, 0, ''])
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        elif len(r) > 8 and r[0] == 'Line No':
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == '-':      # per-line aggregate row
            try:
                ie = int(r[hdr.index('Instructions Executed')] or 0)
                ex = int(r[hdr.index('L1 Wavefronts Shared Excessive')] or 0)
                sm = int(r[hdr.index('# Samples')] or 0)
            except ValueError:
                continue
            a = agg[(fname, int(r[0]))]
            a[0] += ie; a[1] += ex; a[2] += sm; a[3] = r[1].strip()[:100]
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print('total warp instructions', tot, 'samples', tots)
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
        print('%5.1f%% inst %5.1f%% samp  exc=%9d  %s:%d | %s' % (100 * a[0] / tot, 100 * a[2] / tots, a[1], f, ln, a[3]))


if __name__ == '__main__':
    main()
