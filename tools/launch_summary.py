#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: count / mean / max duration per (kernel, grid)."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            h, st = r, i + 1
            break
    kn, mn, mv, gs = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Grid Size')
    d, inst = collections.OrderedDict(), collections.OrderedDict()
    for r in rows[st:]:
        if len(r) > mv:
            try:
                key, val = (r[kn].split('(')[0], r[gs]), float(r[mv].replace(',', ''))
            except ValueError:
                continue
            if r[mn] == 'gpu__time_duration.sum':
                d.setdefault(key, []).append(val / 1000)
            elif r[mn] == 'smsp__inst_executed.sum':
                inst.setdefault(key, []).append(val)
    tot = sum(sum(v) for v in d.values())
    print('| kernel | grid | launches | mean us | max us | share |' + (' warp inst / launch |' if inst else ''))
    print('|---|---|---|---|---|---|' + ('---|' if inst else ''))
    for (k, g), v in d.items():
        n = inst.get((k, g))
        print('| %s | %s | %d | %.1f | %.1f | %.1f%% |' % (k, g, len(v), sum(v) / len(v), max(v), 100 * sum(v) / tot) + (' %.3e |' % (sum(n) / len(n)) if n else ''))


if __name__ == '__main__':
    main()
