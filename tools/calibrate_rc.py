"""How well the rate controller's size ESTIMATE (47 nnz + 22 sum floor(log2 |l|) + 104 coded sub-blocks + 160 | 80 per CU, in 1/16 bit;
oracle/hevc_rc.c orc_rc_cu_estimate == csrc/enc_dev.cuh) tracks the real access-unit size.  TEST INFRASTRUCTURE: runs the CPU model.

    python tools/calibrate_rc.py [--size 960x544] [--frames 24]      -> one JSON line per (content class, QP): estimate / actual"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from hevc_b200.synth import content_clip
    from oracle import encoder_model as em
    opts = {a.split('=')[0]: a.split('=')[1] for a in sys.argv[1:] if '=' in a}
    w, h = (int(x) for x in opts.get('--size', '960x544').split('x'))
    n = int(opts.get('--frames', 24))
    for kind in ('base', 'pan', 'static', 'grain', 'clean'):
        frames = content_clip(kind, w, h, n, seed=0)
        for qp in (20, 26, 32, 38):
            enc = em.ModelEncoder(em.make_params(w, h, 8, qp_i=qp, qp_p=qp + 2, keyint=60, hash_sei=False))
            ratios, est_t, act_t = [], 0.0, 0.0
            for y, u, v in frames:
                au, info = enc.encode(y, u, v)
                est, act = info.est_bits16 / 16.0, 8.0 * len(au)
                ratios.append(est / act)
                est_t += est
                act_t += act
            enc.close()
            print(json.dumps({'class': kind, 'qp_i': qp, 'size': f'{w}x{h}', 'frames': n, 'estimate_over_actual_total': round(est_t / act_t, 4),
                              'per_frame_min': round(min(ratios), 3), 'per_frame_max': round(max(ratios), 3),
                              'per_frame_mean_abs_error': round(float(np.mean(np.abs(np.array(ratios) - 1))), 4)}), flush=True)


if __name__ == '__main__':
    main()
