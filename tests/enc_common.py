"""Shared helpers for the encoder parity tests: build matching parameter blocks for the CPU model and the CUDA path."""
import numpy as np

from hevc_b200.derive import B200Params
from hevc_b200.synth import SynthClip


def b200_params(w, h, depth, keyint=4, hdr10=None):
    hdr10 = (depth == 10) if hdr10 is None else hdr10
    p = B200Params(width=w, height=h, fps_num=30, fps_den=1, bit_depth=depth, profile_idc=2 if depth > 8 else 1, level_idc=120, tier=0,
                   crf=20, vbv_maxrate_kbps=2940, vbv_bufsize_kbit=3528, keyint=keyint, min_keyint=max(2, keyint // 2),
                   colour_primaries=1, transfer_characteristics=1, matrix_coeffs=1)
    if hdr10:
        p.colour_primaries, p.transfer_characteristics, p.matrix_coeffs = 9, 16, 9
        p.hdr10 = p.hrd = p.aud = p.repeat_headers = 1
        p.chroma_loc = 0
        p.master_display = (13250, 34500, 7500, 3000, 34000, 16000, 15635, 16450, 10000000, 50)
        p.max_cll, p.max_fall = 1000, 400
    return p


def model_params(p: B200Params, qp_i, qp_p, hash_sei, rate_control=False, deblock=True, **tools):
    """``tools``: scenecut / intra_in_p / sao / qp_cascade overrides (defaults = hevc_b200.encoder.to_c_params's)"""
    from oracle import encoder_model as em
    t = {'scenecut': 1, 'intra_in_p': 1, 'sao': 1, 'qp_cascade': 1}
    t.update({k: int(v) for k, v in tools.items()})
    m = em.make_params(p.width, p.height, p.bit_depth, qp_i=qp_i, qp_p=qp_p, keyint=p.keyint, fps=(p.fps_num, p.fps_den),
                       hdr10=bool(p.hdr10), hash_sei=hash_sei, level_idc=p.level_idc, tier=p.tier, min_keyint=p.min_keyint,
                       vbv_maxrate_kbps=p.vbv_maxrate_kbps, vbv_bufsize_kbit=p.vbv_bufsize_kbit, rate_control=int(rate_control), deblock=int(deblock), **t)
    return m


def clip_frames(w, h, n, seed=1, noise=2.0):
    clip = SynthClip(w, h, seed=seed, noise=noise)
    return [clip.frame(i) for i in range(n)]


def run_model(p, frames8, qp_i, qp_p, hash_sei=False, force_idr_at=(), rate_control=False, deblock=True, **tools):
    """-> (stream bytes, per-frame AUs, per-frame recon, per-frame (cus, coefs))"""
    from oracle import encoder_model as em
    enc = em.ModelEncoder(model_params(p, qp_i, qp_p, hash_sei, rate_control, deblock, **tools))
    aus, recs, decs, qps, infos = [], [], [], [], []
    sh = p.bit_depth - 8
    for i, (y, u, v) in enumerate(frames8):
        au, info = enc.encode(y.astype(np.uint16) << sh, u.astype(np.uint16) << sh, v.astype(np.uint16) << sh, force_idr=i in force_idr_at)
        aus.append(au)
        qps.append(info.qp)
        infos.append((info.is_idr, info.poc, info.n_intra))
        recs.append(enc.recon())
        decs.append((enc.last_cus(), enc.last_coefs()))
    enc.close()
    run_model.last_qps = qps
    run_model.last_infos = infos
    return b''.join(aus), aus, recs, decs
