"""Derivation layer vs golden vectors generated from the reference (tests/golden/make_derive_golden.py)."""
import gzip
import json
from pathlib import Path

import pytest

from hevc_b200 import derive
from hevc_b200.probe import VideoInfo

GOLD = json.loads(gzip.decompress((Path(__file__).parent / 'golden' / 'derive_golden.json.gz').read_bytes()))


def _info(c):
    hdr = c['hdr']
    return VideoInfo(c['w'], c['h'], c['fps'], 'bt2020' if hdr else 'bt709', 'smpte2084' if hdr else 'bt709',
                     'bt2020nc' if hdr else 'bt709', 'yuv420p', '', '', 2, hdr, 'eng', None, c['duration'])


def test_grid_matches_reference():
    assert len(GOLD['cases']) >= 1000
    for c in GOLD['cases']:
        info = _info(c)
        assert list(derive.calculate_apple_hevc_level(info)) == c['apple_level'], c
        assert list(derive.calculate_nvenc_hevc_level(info)) == c['nvenc_level'], c
        assert list(derive.calculate_dynamic_values(info, False, '')) == c['dynamic'], c
        assert ':'.join(derive.x265_option_list(info)) == c['x265'], c


def test_gop_alignment():
    for key, want in GOLD['gops'].items():
        fps, sec = (float(t) for t in key.split('/'))
        assert derive.compute_aligned_gop(fps, sec) == want, key


def test_hdr_metadata_strings():
    g = GOLD['hdrmeta']
    assert derive.build_hdr_metadata('', '', False) == g['default_x265']
    assert derive.build_hdr_metadata('', '', True) == g['default_nvenc']
    assert derive.build_hdr_metadata('G(1,2)B(3,4)R(5,6)WP(7,8)L(9,10)', '4000,1000', False) == g['custom_x265']


# SURVEY.md section 3.2 known-answer table
@pytest.mark.parametrize('w,h,fps,hdr,dur,level,dyn', [
    (1920, 1080, 30.0, False, 5.0, ('4', 'main'), (19, 20, 2940, 3528, 90)),
    (1280, 720, 30.0, False, 5.0, ('3.1', 'main'), (18, 19, 1176, 1411, 90)),
    (3840, 2160, 30.0, False, 5.0, ('5', 'main'), (20, 21, 11760, 14112, 60)),
    (3840, 2160, 60.0, True, 5.0, ('5.1', 'main'), (19, 20, 23520, 28224, 120)),
    (3840, 2160, 60.0, True, 60.0, ('5.1', 'main'), (21, 22, 23520, 28224, 120)),
    (7680, 4320, 30.0, True, 10.0, ('6', 'main'), (20, 21, 47040, 56448, 60)),
])
def test_known_answers(w, h, fps, hdr, dur, level, dyn):
    info = _info({'w': w, 'h': h, 'fps': fps, 'hdr': hdr, 'duration': dur})
    assert derive.calculate_apple_hevc_level(info) == level
    assert derive.calculate_dynamic_values(info, False, '') == dyn


def test_b200_params_4k60_hdr():
    p = derive.derive_b200_params(_info({'w': 3840, 'h': 2160, 'fps': 60.0, 'hdr': True, 'duration': 5.0}))
    assert (p.profile_idc, p.level_idc, p.tier, p.bit_depth) == (2, 153, 0, 10)
    assert (p.crf, p.vbv_maxrate_kbps, p.vbv_bufsize_kbit, p.keyint, p.min_keyint) == (19, 23520, 28224, 120, 60)
    assert (p.colour_primaries, p.transfer_characteristics, p.matrix_coeffs) == (9, 16, 9)
    assert p.master_display == (13250, 34500, 7500, 3000, 34000, 16000, 15635, 16450, 10000000, 50)
    assert (p.max_cll, p.max_fall, p.aud, p.repeat_headers, p.hrd, p.chroma_loc) == (1000, 400, 1, 1, 1, 0)


def test_b200_params_1080p_sdr():
    p = derive.derive_b200_params(_info({'w': 1920, 'h': 1080, 'fps': 30.0, 'hdr': False, 'duration': 5.0}))
    assert (p.profile_idc, p.level_idc, p.tier, p.bit_depth) == (1, 120, 0, 8)
    assert (p.crf, p.keyint, p.min_keyint, p.aud, p.repeat_headers, p.hrd) == (19, 90, 45, 0, 0, 0)


def test_ffmpeg_argv_branches_match_reference():
    """the libx265 / NVENC command-line branch kept for encoder='cpu'|'nvenc' (core/transcoder.py:357-495)"""
    from pathlib import Path
    from hevc_b200 import transcoder as T
    info = _info({'w': 3840, 'h': 2160, 'fps': 60.0, 'hdr': True, 'duration': 5.0})
    cmd = T.build_ffmpeg_command(Path('in.mp4'), Path('out/in.mp4'), T.build_ffmpeg_params(info, False, ''), 2, 'eng')
    assert cmd == GOLD['cmd_4k60_hdr']
    for c, flags in GOLD['audio'].items():
        assert T.get_audio_flags(int(c)) == flags
    for g in GOLD['nvenc']:
        hdr = g['hdr']
        info = VideoInfo(g['w'], g['h'], g['fps'], 'bt2020' if hdr else 'bt709', 'smpte2084' if hdr else 'bt709',
                         'bt2020nc' if hdr else 'bt709', 'yuv420p', '', '', 6, hdr, 'fra', None, 10.0)
        p = T.build_ffmpeg_params(info, True, 'nvidia b200')
        assert (p.vparams, p.hdr_metadata, p.pix_fmt, p.profile, p.level) == (g['vparams'], g['meta'], g['pix_fmt'], g['profile'], g['level'])
        assert [T.adjust_nvenc_params(p.vparams, a) for a in range(0, 6)] == g['retry']
        assert T.build_ffmpeg_command(Path('a b.mkv'), Path('o/a b.mp4'), p, 6, 'fra', T.adjust_nvenc_params(p.vparams, 2)) == g['cmd']
