"""R-D harness on the CPU model (TEST INFRASTRUCTURE: runs oracle/, never the product path).

The CUDA encoder is byte-identical to the model (tests/test_gpu_encoder.py), so rate / PSNR / SSIM measured here are the
product's.  Prints one JSON line per (content class, tool set, QP) and the Bjontegaard delta rate of every tool set
against the first one.  The x265 anchor column stays empty: libx265 is not in the image.

    python tools/model_rd.py [--size 640x360] [--frames 30] [--classes base,hardcut,...] [--sets r1,r2] [--out file.jsonl]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

TOOLSETS = {           # name -> model parameter overrides
    'r1': dict(scenecut=0, intra_in_p=0, sao=0, qp_cascade=0),         # the round-1 tool set
    'intra': dict(scenecut=0, intra_in_p=1, sao=0, qp_cascade=0),
    'cut': dict(scenecut=1, intra_in_p=1, sao=0, qp_cascade=0),
    'sao': dict(scenecut=1, intra_in_p=1, sao=1, qp_cascade=0),
    'r2': dict(scenecut=1, intra_in_p=1, sao=1, qp_cascade=1),         # everything this round added
    'r2_c3': dict(scenecut=1, intra_in_p=1, sao=1, qp_cascade=1, qp_p_offset=3),     # cascade with the anchors 3 / 4 above the key frame
    'r2_c4': dict(scenecut=1, intra_in_p=1, sao=1, qp_cascade=1, qp_p_offset=4),
    # experiments: P-frame QP offset against the key frame (default +2)
    'r2_p1': dict(scenecut=1, intra_in_p=1, sao=1, qp_cascade=0, qp_p_offset=1),
    'r2_p3': dict(scenecut=1, intra_in_p=1, sao=1, qp_cascade=0, qp_p_offset=3),
    'r2_p4': dict(scenecut=1, intra_in_p=1, sao=1, qp_cascade=0, qp_p_offset=4),
}


def psnr(a, b, peak):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float(10 * np.log10(peak * peak / max(mse, 1e-12)))


def ssim(a, b, peak):
    """mean SSIM over 8x8 windows with stride 4 (the libavfilter / x264 flavour), luma"""
    a, b = a.astype(np.float64), b.astype(np.float64)
    c1, c2 = (0.01 * peak) ** 2, (0.03 * peak) ** 2

    def box(x):
        s = np.cumsum(np.cumsum(np.pad(x, ((1, 0), (1, 0))), 0), 1)
        w = s[8:, 8:] - s[:-8, 8:] - s[8:, :-8] + s[:-8, :-8]
        return w[::4, ::4] / 64.0
    ma, mb = box(a), box(b)
    va, vb, cab = box(a * a) - ma * ma, box(b * b) - mb * mb, box(a * b) - ma * mb
    return float(np.mean(((2 * ma * mb + c1) * (2 * cab + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2))))


def bd_rate(anchor, test):
    """Bjontegaard delta rate (%) of `test` against `anchor`; each a list of (kbps, quality)"""
    a, t = sorted(anchor, key=lambda p: p[1]), sorted(test, key=lambda p: p[1])
    la, qa = np.log([p[0] for p in a]), np.array([p[1] for p in a])
    lt, qt = np.log([p[0] for p in t]), np.array([p[1] for p in t])
    lo, hi = max(qa.min(), qt.min()), min(qa.max(), qt.max())
    if hi <= lo:
        return None
    pa, pt = np.polyfit(qa, la, 3), np.polyfit(qt, lt, 3)
    ia, it = np.polyint(pa), np.polyint(pt)
    avg = ((np.polyval(it, hi) - np.polyval(it, lo)) - (np.polyval(ia, hi) - np.polyval(ia, lo))) / (hi - lo)
    return float((np.exp(avg) - 1) * 100)


def run_point(frames, w, h, depth, qp, fps, keyint, tools, rc=None):
    from oracle import encoder_model as em
    kw = dict(tools)
    off = kw.pop('qp_p_offset', 2)
    if rc:
        kw.update(rate_control=1, vbv_maxrate_kbps=rc[0], vbv_bufsize_kbit=rc[1])
    enc = em.ModelEncoder(em.make_params(w, h, depth, qp_i=qp, qp_p=min(51, qp + off), keyint=keyint, fps=(fps, 1), hdr10=(depth == 10),
                                         hash_sei=False, **kw))
    sh = depth - 8
    nbytes, ps, ss, idr, intra = 0, [], [], 0, 0
    for y, u, v in frames:
        au, info = enc.encode(y.astype(np.uint16) << sh, u.astype(np.uint16) << sh, v.astype(np.uint16) << sh)
        nbytes += len(au)
        ry = enc.recon()[0][:h, :w]
        ps.append(psnr(ry, y.astype(np.uint16) << sh, 255 << sh))
        ss.append(ssim(ry, y.astype(np.uint16) << sh, 255 << sh))
        idr += info.is_idr
        intra += info.n_intra if not info.is_idr else 0
    enc.close()
    return {'kbps': round(nbytes * 8 / 1000 / (len(frames) / fps), 2), 'psnr_y': round(float(np.mean(ps)), 4),
            'ssim_y': round(float(np.mean(ss)), 5), 'idr_frames': idr, 'intra_cus_in_p': intra}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', default='640x360')
    ap.add_argument('--frames', type=int, default=30)
    ap.add_argument('--depth', type=int, default=8)
    ap.add_argument('--fps', type=int, default=30)
    ap.add_argument('--keyint', type=int, default=90)
    ap.add_argument('--classes', default='base,hardcut,pan,static,grain')
    ap.add_argument('--sets', default='r1,r2')
    ap.add_argument('--qps', default='22,27,32,37')
    ap.add_argument('--vbv', default='', help='maxrate,bufsize in kbit: crf + VBV mode (rate control on; --qps are then the crf-derived key-frame QPs)')
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    from hevc_b200.synth import content_clip
    w, h = (int(x) for x in args.size.split('x'))
    sets = args.sets.split(',')
    out = open(args.out, 'w') if args.out else None
    summary = {}
    for kind in args.classes.split(','):
        frames = content_clip(kind, w, h, args.frames, seed=0)
        curves = {}
        for name in sets:
            for qp in (int(q) for q in args.qps.split(',')):
                rc = tuple(int(x) for x in args.vbv.split(',')) if args.vbv else None
                r = run_point(frames, w, h, args.depth, qp, args.fps, args.keyint, TOOLSETS[name], rc=rc)
                if rc:
                    r['vbv'] = list(rc)
                r.update({'class': kind, 'set': name, 'qp_i': qp, 'size': args.size, 'frames': args.frames, 'depth': args.depth,
                          'x265_anchor': None})
                curves.setdefault(name, []).append(r)
                line = json.dumps(r)
                print(line, flush=True)
                if out:
                    out.write(line + '\n')
        for name in sets[1:]:
            bp = bd_rate([(p['kbps'], p['psnr_y']) for p in curves[sets[0]]], [(p['kbps'], p['psnr_y']) for p in curves[name]])
            bs = bd_rate([(p['kbps'], p['ssim_y']) for p in curves[sets[0]]], [(p['kbps'], p['ssim_y']) for p in curves[name]])
            summary[(kind, name)] = (bp, bs)
            line = json.dumps({'class': kind, 'set': name, 'anchor': sets[0], 'bd_rate_psnr_pct': None if bp is None else round(bp, 2),
                               'bd_rate_ssim_pct': None if bs is None else round(bs, 2)})
            print(line, flush=True)
            if out:
                out.write(line + '\n')
    if out:
        out.close()


if __name__ == '__main__':
    main()
