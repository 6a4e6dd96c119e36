"""Generate tests/golden/derive_golden.json by importing the reference's own derivation layer.

Run in the build container only (needs /root/reference):  python tests/golden/make_derive_golden.py
The GPU box has no /root/reference; tests read the committed JSON."""
import itertools
import gzip
import json
import sys
from pathlib import Path

sys.path.insert(0, '/root/reference')
from core import transcoder as ref          # noqa: E402
from core import utils as ref_utils         # noqa: E402
from core.probe import VideoInfo            # noqa: E402

RES = [(640, 360), (854, 480), (1280, 720), (1920, 1080), (1080, 1920), (2560, 1440), (3840, 2160), (4096, 2160), (7680, 4320)]
FPS = [23.976, 24.0, 25.0, 29.97, 30.0, 50.0, 59.94, 60.0, 119.88, 120.0]
DUR = [None, 1.0, 5.0, 10.0, 60.0, 600.0]
cases = []
for (w, h), fps, hdr, dur in itertools.product(RES, FPS, (False, True), DUR):
    info = VideoInfo(w, h, fps, 'bt2020' if hdr else 'bt709', 'smpte2084' if hdr else 'bt709',
                     'bt2020nc' if hdr else 'bt709', 'yuv420p', '', '', 2, hdr, 'eng', None, dur)
    p = ref.build_ffmpeg_params(info, False, 'unknown')
    cases.append({
        'w': w, 'h': h, 'fps': fps, 'hdr': hdr, 'duration': dur,
        'apple_level': list(ref.calculate_apple_hevc_level(info)),
        'nvenc_level': list(ref.calculate_nvenc_hevc_level(info)),
        'dynamic': list(ref.calculate_dynamic_values(info, False, '')),
        'x265': p.vparams[1], 'pix_fmt': p.pix_fmt, 'profile': p.profile, 'level': p.level,
    })
gops = {f'{fps}/{sec}': ref.compute_aligned_gop(fps, sec) for fps in FPS + [12.5, 15.0, 1.0, 0.5, 240.0]
        for sec in (1.0, 2.0, 2.1, 2.5, 2.625, 3.0, 3.15, 8.0)}
hdrmeta = {
    'default_x265': ref_utils.build_hdr_metadata('', '', False),
    'default_nvenc': ref_utils.build_hdr_metadata('', '', True),
    'custom_x265': ref_utils.build_hdr_metadata('G(1,2)B(3,4)R(5,6)WP(7,8)L(9,10)', '4000,1000', False),
}
cmd_info = VideoInfo(3840, 2160, 60.0, 'bt2020', 'smpte2084', 'bt2020nc', 'yuv420p', '', '', 2, True, 'eng', None, 5.0)
cmd = ref.build_ffmpeg_command(Path('in.mp4'), Path('out/in.mp4'), ref.build_ffmpeg_params(cmd_info, False, ''), 2, 'eng')
nvenc = []
for (w, h), fps, hdr in itertools.product(RES, (24.0, 29.97, 60.0, 120.0), (False, True)):
    info = VideoInfo(w, h, fps, 'bt2020' if hdr else 'bt709', 'smpte2084' if hdr else 'bt709', 'bt2020nc' if hdr else 'bt709',
                     'yuv420p', '', '', 6, hdr, 'fra', None, 10.0)
    p = ref.build_ffmpeg_params(info, True, 'nvidia b200')
    nvenc.append({'w': w, 'h': h, 'fps': fps, 'hdr': hdr, 'vparams': p.vparams, 'meta': p.hdr_metadata, 'pix_fmt': p.pix_fmt,
                  'profile': p.profile, 'level': p.level,
                  'retry': [ref.adjust_nvenc_params(p.vparams, a) for a in range(0, 6)],
                  'cmd': ref.build_ffmpeg_command(Path('a b.mkv'), Path('o/a b.mp4'), p, 6, 'fra', ref.adjust_nvenc_params(p.vparams, 2))})
audio = {str(c): ref.get_audio_flags(c) for c in range(0, 10)}
out = Path(__file__).with_name("derive_golden.json.gz")
out.write_bytes(gzip.compress(json.dumps({'cases': cases, 'gops': gops, 'hdrmeta': hdrmeta, 'cmd_4k60_hdr': cmd, 'nvenc': nvenc, 'audio': audio}, separators=(",", ":")).encode(), mtime=0))
print(len(cases), 'cases ->', out)
