"""GPU end-to-end: convert_video(encoder='b200') is a drop-in for the encode step -- Y4M / container in, hvc1 MP4 out,
Apple-compliance checks pass, an independent demuxer+decoder plays it, and the result dictionary / progress contract holds."""
import json
import struct
import threading

import numpy as np
import pytest

from hevc_b200 import compliance, transcoder, upscale
from hevc_b200.frames import write_y4m
from hevc_b200.synth import SynthClip

pytestmark = pytest.mark.gpu


def _decode_mp4(data):
    from oracle import fforacle
    rep = compliance.inspect(data)
    ps = rep['param_sets']
    annexb = b''.join(b'\0\0\0\1' + ps[t][0] for t in (32, 33, 34))
    pos = rep['chunk_offset']
    for size in rep['sample_sizes']:
        end = pos + size
        while pos < end:
            ln = struct.unpack('>I', data[pos:pos + 4])[0]
            annexb += b'\0\0\0\1' + data[pos + 4:pos + 4 + ln]
            pos += 4 + ln
    return rep, fforacle.decode_hevc(annexb, verify_hash=False)


def _psnr(a, b, peak):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10 * np.log10(peak * peak / max(mse, 1e-9))


@pytest.mark.parametrize('hdr', [False, True])
def test_convert_video_y4m(tmp_path, hdr):
    w, h, n = 320, 192, 12
    clip = SynthClip(w, h, seed=4)
    frames = [clip.frame(i) for i in range(n)]
    src = tmp_path / ('hdr.y4m' if hdr else 'sdr.y4m')
    write_y4m(src, frames, w, h, (30, 1))
    if hdr:        # the reference's "HDR" fixtures are 8-bit clips tagged bt2020/smpte2084 (tests/generate_test_videos.py:33-35)
        (tmp_path / 'hdr.y4m.json').write_text(json.dumps({'color_primaries': 'bt2020', 'color_transfer': 'smpte2084', 'color_space': 'bt2020nc'}))
    ticks = []
    res = transcoder.convert_video(src, tmp_path, progress_callback=lambda *a: ticks.append(a), encoder='b200', device=0)
    assert res == {'file': src.name, 'status': 'SUCCESS', 'quality': res['quality'], 'retries': 0, 'method': 'B200', 'hdr': hdr}
    assert ticks and ticks[-1][1] == ticks[-1][2] and all(t[0] == src.name for t in ticks)
    data = (tmp_path / (src.stem + '.mp4')).read_bytes()
    expect = {'profile_idc': 2 if hdr else 1, 'tier': 0, 'hdr10': hdr, 'master_display': (13250, 34500, 7500, 3000, 34000, 16000, 15635, 16450, 10000000, 50),
              'max_cll': 1000, 'max_fall': 400}
    assert compliance.check_bytes(data, expect) == []
    rep, dec = _decode_mp4(data)
    assert len(dec) == n and dec[0][0].shape == (h, w)
    sh = 2 if hdr else 0
    for i in (0, n - 1):
        assert _psnr(dec[i][0], frames[i][0].astype(np.uint16) << sh, 255 << sh) > 28      # the (mis-scaled) VBV cap of the reference binds at this size
    import cv2
    cap = cv2.VideoCapture(str(tmp_path / (src.stem + '.mp4')))
    assert cap.isOpened() and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == n


def test_convert_video_segment_streams(tmp_path, monkeypatch):
    """HEVC_B200_STREAMS=2: closed-GOP segments of one key-frame interval on two encoder streams of the same GPU; same
    contract (result entry, progress ticks, compliant hvc1 MP4, every frame decodes)"""
    w, h, n = 320, 192, 75                                   # 30 fps -> key-frame interval 30: three segments, the last one short
    clip = SynthClip(w, h, seed=6)
    frames = [clip.frame(i) for i in range(n)]
    src = tmp_path / 'seg.y4m'
    write_y4m(src, frames, w, h, (30, 1))
    monkeypatch.setenv('HEVC_B200_STREAMS', '2')
    ticks = []
    res = transcoder.convert_video(src, tmp_path, progress_callback=lambda *a: ticks.append(a), encoder='b200', device=0)
    assert res['status'] == 'SUCCESS' and res['method'] == 'B200'
    assert ticks and ticks[-1][1] == ticks[-1][2] == n
    data = (tmp_path / 'seg.mp4').read_bytes()
    assert compliance.check_bytes(data, {'profile_idc': 1, 'tier': 0, 'hdr10': False}) == []
    rep, dec = _decode_mp4(data)
    assert len(dec) == n
    for i in (0, 31, n - 1):
        assert _psnr(dec[i][0], frames[i][0].astype(np.uint16), 255) > 28


def test_convert_video_container_input_and_cancel(tmp_path):
    """BGR frames from a container (OpenCV decode) go through the device CSC kernel; the stop event cancels between batches"""
    import cv2
    w, h, n = 256, 144, 40
    src = tmp_path / 'in.avi'
    vw = cv2.VideoWriter(str(src), cv2.VideoWriter_fourcc(*'FFV1'), 25.0, (w, h))
    clip = SynthClip(w, h, seed=6, noise=0.0)
    for i in range(n):
        y, u, v = clip.frame(i)
        yuv = np.concatenate([y.reshape(-1), u.reshape(-1), v.reshape(-1)]).reshape(h * 3 // 2, w)
        vw.write(cv2.cvtColor(yuv, cv2.COLOR_YUV2BGR_I420))
    vw.release()
    res = transcoder.convert_video(src, tmp_path, encoder='b200', device=0)
    assert res['status'] == 'SUCCESS' and res['method'] == 'B200'
    rep, dec = _decode_mp4((tmp_path / 'in.mp4').read_bytes())
    assert len(dec) == n and rep['sps']['width'] == w
    assert _psnr(dec[3][0], clip.frame(3)[0], 255) > 26       # tiny picture: the reference VBV clamp for this level binds hard
    ev = threading.Event()
    ev.set()
    other = tmp_path / 'cancelled'
    other.mkdir()
    assert transcoder.convert_video(src, other, encoder='b200', device=0, stop_event=ev)['status'] == 'CANCELLED'
    assert list(other.iterdir()) == []                      # a cancelled run leaves neither a partial MP4 nor a spool file


def test_upscale_and_encode(tmp_path):
    """config 4 in miniature: geometry rule + scaler fused into the encoder's ingest -> Main10.  Main10 is a bit depth, not a
    transfer function: an SDR BT.709 source keeps SDR signalling (no HDR10 SEI, no mdcv / clli, VUI = the source's tags)."""
    assert upscale.target_geometry(1920, 1080) == (3840, 2160) and upscale.target_geometry(1280, 720) == (1920, 1080)
    w, h, n = 480, 270, 6
    clip = SynthClip(w, h, seed=8, noise=0.0)
    src = tmp_path / 'small.y4m'
    write_y4m(src, [clip.frame(i) for i in range(n)], w, h, (30, 1))
    (tmp_path / 'small.y4m.json').write_text(json.dumps({'color_primaries': 'bt709', 'color_transfer': 'bt709', 'color_space': 'bt709'}))
    res = upscale.process_video(src, tmp_path, target_height=540, device=0)
    assert res['status'] == 'SUCCESS' and (res['width'], res['height']) == (960, 540) and res['hdr'] is False, res
    data = (tmp_path / 'small.mp4').read_bytes()
    rep, dec = _decode_mp4(data)
    sps = rep['sps']
    assert len(dec) == n and dec[0][0].shape == (540, 960) and sps['bit_depth'] == 10 and sps['profile_idc'] == 2
    assert (sps['colour_primaries'], sps['transfer_characteristics'], sps['matrix_coeffs']) == (1, 1, 1)
    assert 137 not in rep['sei'] and 144 not in rep['sei'] and not sps.get('hrd')
    assert 'mdcv' not in rep['entry_boxes'] and 'clli' not in rep['entry_boxes']
    assert compliance.check_bytes(data, {'profile_idc': 2, 'tier': 0, 'hdr10': False}) == []
    from oracle import pixel_ref
    want = pixel_ref.scale_plane(clip.frame(2)[0], 960, 540, 10)
    assert _psnr(dec[2][0], want, 1023) > 33


def test_upscale_hdr_source_keeps_hdr10(tmp_path):
    w, h, n = 480, 270, 4
    clip = SynthClip(w, h, seed=9, noise=0.0)
    src = tmp_path / 'hdr.y4m'
    write_y4m(src, [clip.frame(i) for i in range(n)], w, h, (30, 1))
    (tmp_path / 'hdr.y4m.json').write_text(json.dumps({'color_primaries': 'bt2020', 'color_transfer': 'smpte2084', 'color_space': 'bt2020nc'}))
    res = upscale.process_video(src, tmp_path, target_height=540, device=0)
    assert res['status'] == 'SUCCESS' and res['hdr'] is True
    assert compliance.check_bytes((tmp_path / 'hdr.mp4').read_bytes(), {'profile_idc': 2, 'hdr10': True}) == []


def test_upscale_segments_over_devices(tmp_path):
    """config 4's composition: one clip -> closed-GOP segments of one key-frame interval dealt round-robin to one worker per
    entry of ``devices`` (here the same GPU twice), fused scale + Main10 encode per worker, in-order collection, ONE mux;
    the result decodes as one stream and equals the single-worker encode of the same segments frame for frame."""
    w, h, n = 320, 180, 200                                  # 30 fps -> key-frame interval 90: segments 90 / 90 / 20
    clip = SynthClip(w, h, seed=10)
    frames = [clip.frame(i) for i in range(n)]
    src = tmp_path / 'long.y4m'
    write_y4m(src, frames, w, h, (30, 1))
    ticks = []
    res = upscale.process_video(src, tmp_path, target_height=360, devices=[0, 0], progress_callback=lambda *a: ticks.append(a))
    assert res['status'] == 'SUCCESS' and (res['width'], res['height']) == (640, 360), res
    rep, dec = _decode_mp4((tmp_path / 'long.mp4').read_bytes())
    assert len(dec) == n and ticks and ticks[-1][1] == n
    sync = rep['sync_samples']
    assert sync == [1, 91, 181]                                    # every segment opens with an IDR
    from oracle import pixel_ref
    for i in (0, 89, 90, 199):
        assert _psnr(dec[i][0], pixel_ref.scale_plane(frames[i][0], 640, 360, 10), 1023) > 25      # the reference VBV clamp binds hard at this size


@pytest.mark.parametrize('tagged_hdr', [False, True])
def test_convert_video_10bit_y4m(tmp_path, tagged_hdr):
    """10-bit planar source (HB_PIX_YUV420P16 with src_bit_depth = 10): without HDR tags the reference would hand libx265
    8-bit samples (-pix_fmt yuv420p), so the ingest stage rounds to 8 bits; with HDR tags it is a Main10 encode at full depth."""
    w, h, n = 256, 144, 6
    clip = SynthClip(w, h, seed=12, noise=0.0)
    f8 = [clip.frame(i) for i in range(n)]
    rng = np.random.default_rng(3)
    f10 = [tuple((p.astype(np.uint16) << 2) + rng.integers(0, 4, p.shape, dtype=np.uint16) for p in f) for f in f8]
    src = tmp_path / 'ten.y4m'
    write_y4m(src, f10, w, h, (30, 1), ten_bit=True)
    if tagged_hdr:
        (tmp_path / 'ten.y4m.json').write_text(json.dumps({'color_primaries': 'bt2020', 'color_transfer': 'smpte2084', 'color_space': 'bt2020nc'}))
    res = transcoder.convert_video(src, tmp_path, encoder='b200', device=0)
    assert res['status'] == 'SUCCESS' and res['hdr'] is tagged_hdr
    rep, dec = _decode_mp4((tmp_path / 'ten.mp4').read_bytes())
    assert rep['sps']['bit_depth'] == (10 if tagged_hdr else 8) and len(dec) == n
    if tagged_hdr:
        assert _psnr(dec[2][0], f10[2][0], 1023) > 26
    else:
        want = np.minimum((f10[2][0].astype(np.int32) + 2) >> 2, 255)
        assert _psnr(dec[2][0], want, 255) > 26 and int(dec[2][0].max()) <= 255


def test_convert_video_main10_hdr10_container(tmp_path):
    """a Main10 HDR10 MP4 in (made by the CPU model + muxer) -> convert_video(encoder='b200'): the source is decoded at 10 bits
    through the bundled libavcodec (not 8-bit BGR), the prober picks the HDR10 boxes up, and the output is a compliant Main10
    HDR10 hvc1 MP4 carrying the SAME mastering-display / content-light values, close to the 10-bit source"""
    from tests.test_host_pipeline import _model_mp4
    src, dec_src = _model_mp4(tmp_path, 10, w=256, h=144, n=8, name='hdr_in.mp4')
    out_dir = tmp_path / 'out'
    out_dir.mkdir()
    res = transcoder.convert_video(src, out_dir, encoder='b200', device=0)
    assert res['status'] == 'SUCCESS' and res['hdr'] is True and res['method'] == 'B200'
    data = (out_dir / 'hdr_in.mp4').read_bytes()
    expect = {'profile_idc': 2, 'tier': 0, 'hdr10': True, 'master_display': (13250, 34500, 7500, 3000, 34000, 16000, 15635, 16450, 10000000, 50),
              'max_cll': 1000, 'max_fall': 400}
    assert compliance.check_bytes(data, expect) == []
    rep, dec = _decode_mp4(data)
    assert len(dec) == 8 and rep['sps']['bit_depth'] == 10
    assert _psnr(dec[3][0], dec_src[3][0], 1023) > 30          # against the 10-bit decoded source, full depth kept


@pytest.mark.parametrize('tagged_hdr', [False, True])
def test_convert_video_422_10bit_master(tmp_path, tagged_hdr):
    """a 4:2:2 10-bit master (the ProRes-style source of the reference's users; here Y4M C422p10): libswscale reduces the chroma to 4:2:0 at
    10 bits on the host (what ffmpeg does for -pix_fmt, core/transcoder.py:464), HDR-tagged it becomes a Main10 encode at full depth,
    untagged the reference's 8-bit Main encode"""
    w, h, n = 256, 144, 6
    clip = SynthClip(w, h, seed=14, noise=0.0)
    rng = np.random.default_rng(5)
    frames = []
    for i in range(n):
        y, u, v = clip.frame(i)
        y10 = (y.astype(np.uint16) << 2) + rng.integers(0, 4, y.shape, dtype=np.uint16)
        frames.append((y10, np.repeat(u.astype(np.uint16) << 2, 2, axis=0), np.repeat(v.astype(np.uint16) << 2, 2, axis=0)))   # 4:2:2: full-height chroma
    src = tmp_path / 'master.y4m'
    with open(src, 'wb') as fh:
        fh.write(f'YUV4MPEG2 W{w} H{h} F30:1 Ip A1:1 C422p10\n'.encode())
        for planes in frames:
            fh.write(b'FRAME\n')
            for pl in planes:
                fh.write(np.ascontiguousarray(pl, dtype='<u2').tobytes())
    if tagged_hdr:
        (tmp_path / 'master.y4m.json').write_text(json.dumps({'color_primaries': 'bt2020', 'color_transfer': 'smpte2084', 'color_space': 'bt2020nc'}))
    res = transcoder.convert_video(src, tmp_path, encoder='b200', device=0)
    assert res['status'] == 'SUCCESS' and res['hdr'] is tagged_hdr and res['method'] == 'B200', res
    rep, dec = _decode_mp4((tmp_path / 'master.mp4').read_bytes())
    assert rep['sps']['bit_depth'] == (10 if tagged_hdr else 8) and len(dec) == n and dec[0][0].shape == (h, w)
    if tagged_hdr:
        assert _psnr(dec[2][0], frames[2][0], 1023) > 26
        assert _psnr(dec[2][1], clip.frame(2)[1].astype(np.uint16) << 2, 1023) > 30      # chroma survived the 4:2:2 -> 4:2:0 reduction
    else:
        want = np.minimum((frames[2][0].astype(np.int32) + 2) >> 2, 255)
        assert _psnr(dec[2][0], want, 255) > 26 and int(dec[2][0].max()) <= 255
