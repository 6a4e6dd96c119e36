#!/usr/bin/env python
"""Read the per-row CABAC trace written with HB_ENTROPY_TRACE=<file> (debug hook of the encoder): for every frame of the last
batch print the row hand-off ramp, per-row coding time and bytes."""
import sys

import numpy as np


def main():
    raw = np.fromfile(sys.argv[1], dtype=np.uint64)
    n, rows = int(raw[0]), int(raw[1])
    rec = raw[2:].reshape(n, rows, 6).astype(np.int64)
    t0 = rec[:, :, 0].min()
    for i in ([0, 1, n // 2, n - 1] if len(sys.argv) < 3 else [int(a) for a in sys.argv[2:]]):
        st, en, by = rec[i, :, 0] - t0, rec[i, :, 1] - t0, rec[i, :, 2]
        dur = en - st
        print('frame %d: first start %.3f ms, last end %.3f ms, latency %.3f ms, bytes %d' % (i, st.min() / 1e6, en.max() / 1e6, (en.max() - st.min()) / 1e6, by.sum()))
        print('  row time us: mean %.0f max %.0f (row %d) | ramp per row us: mean %.1f max %.1f' % (dur.mean() / 1e3, dur.max() / 1e3, int(dur.argmax()),
              np.diff(st).mean() / 1e3, np.diff(st).max() / 1e3))
        cb, cc, ne = rec[i, :, 3].sum(), rec[i, :, 4].sum(), rec[i, :, 5].sum()
        print('  lane-0 cycles per row (mean): all-lane phase (staging waits, masks, binarisation) %.0f, coding %.0f; list entries %d -> %.0f cycles / '
              'entry in the coding loop, %.0f ns of row time per entry' % (cb / rows, cc / rows, ne, cc / max(1, ne), dur.sum() / max(1, ne)))
        print('  ns per byte: %.0f  | rows: ' % (dur.sum() / max(1, by.sum())) + ' '.join('%d:%dus/%dB' % (r, dur[r] / 1e3, by[r]) for r in range(0, rows, 6)))


if __name__ == '__main__':
    main()
