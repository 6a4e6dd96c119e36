"""The CPU encoder model is pinned by the normative side: every stream it writes must decode under the FFmpeg
hevc decoder (with decoded-picture-hash verification on) to exactly the model's own reconstruction."""
import numpy as np
import pytest

from hevc_b200.synth import SynthClip
from oracle import encoder_model as em
from oracle import fforacle


def _run(w, h, depth, frames, qp, keyint=4, seed=1, **kw):
    clip = SynthClip(w, h, seed=seed)
    enc = em.ModelEncoder(em.make_params(w, h, depth, qp_i=qp, qp_p=min(51, qp + 2), keyint=keyint, hdr10=(depth == 10), **kw))
    stream, recs, infos = b'', [], []
    for n in range(frames):
        y, u, v = clip.frame(n)
        if depth == 10:
            y, u, v = (a.astype(np.uint16) << 2 for a in (y, u, v))
        au, info = enc.encode(y, u, v)
        stream += au
        recs.append(enc.recon())
        infos.append(info)
    return stream, recs, infos


@pytest.mark.parametrize('w,h,depth,frames,qp', [
    (192, 112, 8, 6, 26), (200, 120, 8, 5, 10), (200, 120, 10, 5, 4), (328, 184, 8, 5, 40), (64, 64, 10, 3, 51),
    (32, 16, 8, 3, 20), (416, 240, 10, 5, 22), (640, 360, 8, 3, 0)])
def test_decoder_matches_model_reconstruction(w, h, depth, frames, qp):
    stream, recs, infos = _run(w, h, depth, frames, qp)
    decoded = fforacle.decode_hevc(stream, verify_hash=True)       # raises on any picture-hash mismatch
    assert len(decoded) == frames
    for dec, rec in zip(decoded, recs):
        for c in range(3):
            assert dec[c].shape == ((h, w) if c == 0 else (h // 2, w // 2))
            assert (dec[c] == rec[c][:dec[c].shape[0], :dec[c].shape[1]]).all()
    assert [i.is_idr for i in infos] == [int(n % 4 == 0) for n in range(frames)]
    assert all(i.psnr_y > (21 if qp > 45 else 22 if qp > 30 else 30) for i in infos)      # sanity floor, not a quality claim


def test_deblocking_filter_on_and_off():
    """both settings decode exactly; the filter changes the reconstruction and helps PSNR at coarse quantisers"""
    a, ra, ia = _run(200, 120, 8, 5, 38, deblock=1)
    b, rb, ib = _run(200, 120, 8, 5, 38, deblock=0)
    for stream, recs in ((a, ra), (b, rb)):
        dec = fforacle.decode_hevc(stream, verify_hash=True)
        assert all((d[c] == r[c][:d[c].shape[0], :d[c].shape[1]]).all() for d, r in zip(dec, recs) for c in range(3))
    assert a != b and np.mean([i.psnr_y for i in ia]) > np.mean([i.psnr_y for i in ib])


def test_corrupted_stream_is_detected():
    stream, _, _ = _run(192, 112, 8, 2, 26)
    bad = bytearray(stream)
    bad[len(bad) // 2] ^= 0x10
    with pytest.raises(fforacle.DecodeError):
        fforacle.decode_hevc(bytes(bad), verify_hash=True)


def test_static_content_is_skipped():
    w, h = 128, 64
    y = np.full((h, w), 90, np.uint8)
    u = np.full((h // 2, w // 2), 100, np.uint8)
    enc = em.ModelEncoder(em.make_params(w, h, 8, qp_i=30, qp_p=30, keyint=10))
    enc.encode(y, u, u)
    au, info = enc.encode(y, u, u)
    assert info.n_skip == (w // 16) * (h // 16) and info.bytes < 100


def test_hdr10_headers_present():
    stream, _, _ = _run(64, 64, 10, 2, 30)
    types = [t for _, t, _ in fforacle.iter_nals(stream)]
    assert types[:4] == [35, 32, 33, 34]            # AUD, VPS, SPS, PPS
    assert types.count(39) >= 5                     # buffering period, pic timing x2, mastering display, content light level
    assert 19 in types and 1 in types and 40 in types


def vbv_underflows(au_bits, maxrate_kbps, bufsize_kbit, fps):
    """Leaky-bucket check of ACTUAL access-unit sizes against the HRD the stream signals (Annex C, VBR / cbr_flag = 0): the
    CPB fills at vbv-maxrate up to vbv-bufsize, the first removal happens at initial_cpb_removal_delay = 0.9 x bufsize / maxrate
    (what the buffering-period SEI carries), one access unit leaves per frame interval.  -> [(frame, missing bits)]"""
    B, per = bufsize_kbit * 1000.0, maxrate_kbps * 1000.0 / fps
    fullness, bad = 0.9 * B, []
    for i, bits in enumerate(au_bits):
        if bits > fullness + 1e-6:
            bad.append((i, bits - fullness))
        fullness = min(B, max(0.0, fullness - bits) + per)
    return bad


@pytest.mark.parametrize('depth,kbps,qp', [(8, 300, 20), (10, 250, 20), (8, 120, 16)])
def test_vbv_conformance_of_actual_sizes(depth, kbps, qp):
    """wherever hrd=1 is signalled (reference core/utils.py:65) the real access-unit sizes -- not the controller's estimates --
    must keep the decoder buffer from underflowing, over several GOPs, at caps that really bind"""
    w, h, n, keyint = 320, 192, 40, 12
    stream, _, infos = _run(w, h, depth, n, qp, keyint=keyint, seed=7, rate_control=1, vbv_maxrate_kbps=kbps,
                            vbv_bufsize_kbit=int(kbps * 1.2), hrd=1)
    assert max(i.qp for i in infos) > qp + 4                      # the cap binds
    bits = [8 * i.bytes for i in infos]
    assert vbv_underflows(bits, kbps, int(kbps * 1.2), 30.0) == []
    # the estimate the controller steers by tracks the real size
    tot_est = sum(i.est_bits16 for i in infos) / 16.0
    assert abs(tot_est - sum(bits)) / sum(bits) < 0.15
