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
    kn, mv, gs = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
    d = collections.OrderedDict()
    for r in rows[st:]:
        if len(r) > mv:
            try:
                d.setdefault((r[kn].split('(')[0], r[gs]), []).append(float(r[mv].replace(',', '')) / 1000)
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    print('| kernel | grid | launches | mean us | max us | share |')
    print('|---|---|---|---|---|---|')
    for (k, g), v in d.items():
        print('| %s | %s | %d | %.1f | %.1f | %.1f%% |' % (k, g, len(v), sum(v) / len(v), max(v), 100 * sum(v) / tot))


if __name__ == '__main__':
    main()
