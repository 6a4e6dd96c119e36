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
    assert transcoder.convert_video(src, tmp_path, encoder='b200', device=0, stop_event=ev)['status'] == 'CANCELLED'


def test_upscale_and_encode(tmp_path):
    """config 4 in miniature: geometry rule + fused scale -> P010 -> Main10"""
    assert upscale.target_geometry(1920, 1080) == (3840, 2160) and upscale.target_geometry(1280, 720) == (1920, 1080)
    w, h, n = 480, 270, 6
    clip = SynthClip(w, h, seed=8, noise=0.0)
    src = tmp_path / 'small.y4m'
    write_y4m(src, [clip.frame(i) for i in range(n)], w, h, (30, 1))
    res = upscale.process_video(src, tmp_path, target_height=540, device=0)
    assert res['status'] == 'SUCCESS' and (res['width'], res['height']) == (960, 540), res
    rep, dec = _decode_mp4((tmp_path / 'small.mp4').read_bytes())
    assert len(dec) == n and dec[0][0].shape == (540, 960) and rep['sps']['bit_depth'] == 10
    from oracle import pixel_ref
    want = pixel_ref.scale_plane(clip.frame(2)[0], 960, 540, 10)
    assert _psnr(dec[2][0], want, 1023) > 33
