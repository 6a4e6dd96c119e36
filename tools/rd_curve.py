"""Rate-distortion points of the B200 encoder at constant QP (the x265 anchor column stays empty: libx265 is not in the image).

    python tools/rd_curve.py [clip_type] [frames]      -> JSON lines: qp, kbit/s, PSNR-Y/U/V (dB) from the decoder output"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def psnr(a, b, peak):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float(10 * np.log10(peak * peak / max(mse, 1e-12)))


def main():
    import torch
    from hevc_b200 import _cabi, derive, encoder as E
    from hevc_b200.probe import VideoInfo
    from hevc_b200.synth import CLIP_TYPES, TorchSynthClip
    from oracle import fforacle
    name = sys.argv[1] if len(sys.argv) > 1 else '1080p_sdr'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    w, h, fps, hdr = CLIP_TYPES[name]
    tags = ('bt2020', 'smpte2084', 'bt2020nc') if hdr else ('bt709', 'bt709', 'bt709')
    p = derive.derive_b200_params(VideoInfo(w, h, float(fps), *tags, 'yuv420p', '', '', 0, hdr, None, None, 5.0))
    ctx = _cabi.Context(0)
    clip = TorchSynthClip(w, h, seed=0, device='cuda:0')
    frames = clip.frames(0, n).cpu().numpy()
    sh = p.bit_depth - 8
    lw, cw = w * h, (w // 2) * (h // 2)
    for deblock in (True, False):
        for qp in (22, 27, 32, 37):
            enc = E.B200Encoder(ctx, E.to_c_params(p, qp=(qp, qp + 2), rate_control=False, deblock=deblock), max_batch=n)
            stream, stats = enc.encode(frames, n)
            enc.close()
            dec = fforacle.decode_hevc(stream, verify_hash=False)
            py, pu, pv = [], [], []
            for i, d in enumerate(dec):
                f = frames[i]
                py.append(psnr(d[0], f[:lw].reshape(h, w).astype(np.uint16) << sh, 255 << sh))
                pu.append(psnr(d[1], f[lw:lw + cw].reshape(h // 2, w // 2).astype(np.uint16) << sh, 255 << sh))
                pv.append(psnr(d[2], f[lw + cw:].reshape(h // 2, w // 2).astype(np.uint16) << sh, 255 << sh))
            print(json.dumps({'clip': name, 'frames': n, 'deblock': deblock, 'qp_i': qp, 'qp_p': qp + 2,
                              'kbps': round(len(stream) * 8 / 1000 / (n / fps), 1), 'psnr_y': round(float(np.mean(py)), 3),
                              'psnr_u': round(float(np.mean(pu)), 3), 'psnr_v': round(float(np.mean(pv)), 3)}))
    ctx.close()


if __name__ == '__main__':
    main()
